"""Model API mirror: ``DeepChopper.from_pretrained`` / ``from_checkpoint`` (deepchopper/models/dc_hg.py:70-163)
returning a module with ``forward(input_ids, input_quals) -> logits[B,L,2]`` and Lightning-style
``predict_step(batch, idx) -> (logits, batch["labels"])`` (deepchopper/models/basic_module.py:90-100,197-207).

The arithmetic runs in libdcb200 (tcgen05 GEMMs, Toeplitz / blocked-FFT long conv, fused head) -- bf16 operands,
fp32 accumulation, fp32 residual stream.  CUDA only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _native
from ._native import check, lib

BACKBONE_NAME = "hyenadna-small-32k-seqlen"   # dc_hg.py:128
MAX_TOKENS = 32768
ROW_TILE = 128


class Weights:
    """Device-resident weights (dcb200_weights) built from a reference state dict."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], ctx: _native.Context):
        names, ptrs, numels, keep = [], [], [], []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        n = len(names)
        arr_names = (C.c_char_p * n)(*names)
        arr_ptrs = (C.c_void_p * n)(*ptrs)
        arr_numel = (C.c_int64 * n)(*numels)
        self._h = C.c_void_p()
        self.ctx = ctx
        check(lib().dcb200_weights_create(ctx.handle, arr_names, arr_ptrs, arr_numel, n, C.byref(self._h)))
        del keep
        from . import ops
        self.op_handle = ops.register_weights(self)   # what torch.ops.dcb200.forward takes

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            from . import ops
            ops.unregister_weights(self.op_handle)
            lib().dcb200_weights_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    if os.path.isdir(path):
        for name in ("model.safetensors", "pytorch_model.bin", "model.ckpt"):
            if os.path.exists(os.path.join(path, name)):
                path = os.path.join(path, name)
                break
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "state_dict" in obj:      # Lightning .ckpt (dc_hg.py:90-117)
        obj = obj["state_dict"]
    return obj


class DeepChopperModel(torch.nn.Module):
    """Inference twin of ``TokenClassificationLit`` (basic_module.py:34)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device: int | str | torch.device = 0):
        super().__init__()
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise _native.Dcb200Error("deepchopper_b200 runs on CUDA (B200, sm_100a) only; there is no CPU fallback")
        self._device = dev
        self._ctx = _native.torch_context(dev)
        self._weights = Weights(state_dict, self._ctx)

    @property
    def device(self):
        return self._device

    def _ctx_now(self) -> _native.Context:
        return _native.torch_context(self._device)

    @torch.no_grad()
    def forward_tokens(self, tok: torch.Tensor, quals: torch.Tensor, want_logits: bool = True,
                       want_labels: bool = False):
        """tok uint8 [B,L], quals fp32 [B,L] on the device, L % 128 == 0.  Returns (logits|None, labels|None)."""
        B, L = tok.shape
        assert L % ROW_TILE == 0 and tok.dtype == torch.uint8 and quals.dtype == torch.float32
        assert tok.is_contiguous() and quals.is_contiguous()
        logits, labels = torch.ops.dcb200.forward(tok, quals, self._weights.op_handle, bool(want_logits), bool(want_labels))
        return (logits if want_logits else None), (labels if want_labels else None)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, input_quals: torch.Tensor) -> torch.Tensor:
        """basic_module.py:90-100.  Any [B, L]: rows are right-filled to a multiple of 128 (the model is
        causal, so the filler cannot influence the returned columns)."""
        B, L = input_ids.shape
        if L > MAX_TOKENS:
            raise ValueError(f"sequence length {L} exceeds {MAX_TOKENS}")
        Lrow = (L + ROW_TILE - 1) // ROW_TILE * ROW_TILE
        dev = self._device
        tok = torch.full((B, Lrow), 4, dtype=torch.uint8, device=dev)
        tok[:, :L] = input_ids.to(dev).to(torch.uint8)
        q = torch.zeros((B, Lrow), dtype=torch.float32, device=dev)
        q[:, :L] = input_quals.to(dev, torch.float32)
        logits, _ = self.forward_tokens(tok, q, True, False)
        return logits[:, :L, :] if Lrow != L else logits

    def predict_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0):
        """basic_module.py:197-207."""
        return self.forward(batch["input_ids"], batch["input_quals"]), batch["labels"]

    def eval(self):
        return self

    def close(self):
        """Release the device weights now (otherwise when the module is garbage-collected)."""
        self._weights.close()


class DeepChopper:
    """``deepchopper.DeepChopper`` (dc_hg.py:18-163): factory only."""

    @staticmethod
    def from_state_dict(state_dict, device=0) -> DeepChopperModel:
        return DeepChopperModel(state_dict, device)

    @staticmethod
    def from_checkpoint(checkpoint_path: str, device=0) -> DeepChopperModel:
        """dc_hg.py:70-117: Lightning ``.ckpt`` (``state_dict`` entry) or any torch/safetensors state dict."""
        return DeepChopperModel(_load_state_dict(str(checkpoint_path)), device)

    @staticmethod
    def from_pretrained(pretrained_model_name_or_path: str, device=0, **kwargs) -> DeepChopperModel:
        """dc_hg.py:119-163.  A local directory/file is loaded directly; a hub id (``yangliz5/deepchopper``)
        is fetched with huggingface_hub when a network/cache is available."""
        p = str(pretrained_model_name_or_path)
        if os.path.exists(p):
            return DeepChopperModel(_load_state_dict(p), device)
        try:
            from huggingface_hub import hf_hub_download
            path = hf_hub_download(p, "model.safetensors", **{k: v for k, v in kwargs.items() if k in ("revision", "cache_dir", "token")})
        except Exception as e:  # noqa: BLE001
            raise FileNotFoundError(f"cannot resolve '{p}' locally and the hub is unreachable: {e}") from e
        return DeepChopperModel(_load_state_dict(path), device)
