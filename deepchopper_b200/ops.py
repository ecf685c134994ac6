"""``torch.library`` custom ops of the model path (north star: "Python/PyTorch custom ops for the model path").

Three ops in the ``dcb200`` namespace, CUDA only (there is no CPU kernel: a CPU tensor raises torch's own
"no kernel for this backend" error -- no fallback), each a thin call through the C ABI of include/dcb200.h on the
CURRENT torch stream:

* ``torch.ops.dcb200.encode(blob, seq_off, qual_off, lens, Lpad, Lrow) -> (tok u8 [R,Lrow], qual f32 [R,Lrow])``
  replaces ``tokenize_and_align_labels_and_quals_ids`` + the collator (deepchopper/models/llm/tokenizer.py:34-93,145-178).
* ``torch.ops.dcb200.forward(tok, qual, weights, want_logits, want_labels) -> (logits f32 [B,L,2], labels u8 [B,L])``
  replaces ``TokenClassificationModule.forward`` (deepchopper/models/llm/hyena.py:29-41); ``weights`` is the integer
  handle of a registered :class:`deepchopper_b200.model.Weights` (an op argument must be a tensor or a scalar).
* ``torch.ops.dcb200.smooth_chop(labels, starts, lens, qual_lens, params[8]) -> (n_adapter, adapter_iv, n_keep,
  keep_iv, action)`` replaces the interval step of ``deepchopper-chop`` (src/bin/predict.rs:130-192,
  src/smooth/predict.rs:186-209); fp32 ``labels`` [N,2] are read as logits (argmax fused, src/smooth/predict.rs:275).

``DeepChopperModel.forward`` / ``forward_tokens``, ``encode.encode_batch_device`` and ``smooth.smooth_chop_device`` call
these ops, so the module is traceable as ordinary torch ops (Meta implementations give the output shapes).

The ops are defined with the low-level ``torch.library.Library`` API (schema + CUDA / Meta kernels): the
``torch.library.custom_op`` decorator costs 2.7 s on the first call of a process (it pulls in the fake-tensor / FX stack),
which is more than a 100k-read ``predict --chop`` run spends on the GPU.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Tuple

import torch

from . import _native
from ._native import ChopParams, check, lib

_WEIGHTS = weakref.WeakValueDictionary()   # handle -> model.Weights (owned by its DeepChopperModel)
_next_handle = [1]


def register_weights(w) -> int:
    """Make ``w`` (model.Weights) reachable from the integer handle the forward op takes."""
    h = _next_handle[0]
    _next_handle[0] += 1
    _WEIGHTS[h] = w
    return h


def unregister_weights(handle: int):
    _WEIGHTS.pop(int(handle), None)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_LIB = torch.library.Library("dcb200", "DEF")
_LIB.define("encode(Tensor blob, Tensor seq_off, Tensor qual_off, Tensor lens, int Lpad, int Lrow) -> (Tensor, Tensor)")
_LIB.define("encode_rows(Tensor blob, Tensor seq_off, Tensor qual_off, Tensor lens, Tensor lpad_rows, int Lpad, int Lrow) -> "
            "(Tensor, Tensor)")
_LIB.define("forward(Tensor tok, Tensor qual, int weights, bool want_logits, bool want_labels) -> (Tensor, Tensor)")
_LIB.define("smooth_chop(Tensor labels, Tensor starts, Tensor lens, Tensor qual_lens, int[] params) -> "
            "(Tensor, Tensor, Tensor, Tensor, Tensor)")


def encode(blob: torch.Tensor, seq_off: torch.Tensor, qual_off: torch.Tensor, lens: torch.Tensor, Lpad: int,
           Lrow: int) -> Tuple[torch.Tensor, torch.Tensor]:
    assert blob.dtype == torch.uint8 and seq_off.dtype == torch.int64 and qual_off.dtype == torch.int64
    assert lens.dtype == torch.int32 and blob.is_contiguous() and seq_off.is_contiguous() and qual_off.is_contiguous()
    R = int(lens.numel())
    tok = torch.empty((R, Lrow), dtype=torch.uint8, device=blob.device)
    qual = torch.empty((R, Lrow), dtype=torch.float32, device=blob.device)
    ctx = _native.torch_context(blob.device)
    check(lib().dcb200_encode_batch(ctx.handle, _p(blob), _p(seq_off), _p(qual_off), _p(lens), R, int(Lpad), int(Lrow),
                                    _p(tok), _p(qual)))
    return tok, qual


def _encode_meta(blob, seq_off, qual_off, lens, Lpad, Lrow):
    R = lens.numel()
    return blob.new_empty((R, Lrow), dtype=torch.uint8), blob.new_empty((R, Lrow), dtype=torch.float32)


def encode_rows(blob: torch.Tensor, seq_off: torch.Tensor, qual_off: torch.Tensor, lens: torch.Tensor,
                lpad_rows: torch.Tensor, Lpad: int, Lrow: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """dcb200_encode_batch_rows: every row left-padded to ITS OWN collated length (several reference batches, one launch)."""
    assert blob.dtype == torch.uint8 and seq_off.dtype == torch.int64 and qual_off.dtype == torch.int64
    assert lens.dtype == torch.int32 and lpad_rows.dtype == torch.int32 and lpad_rows.numel() == lens.numel()
    assert blob.is_contiguous() and seq_off.is_contiguous() and qual_off.is_contiguous() and lpad_rows.is_contiguous()
    R = int(lens.numel())
    tok = torch.empty((R, Lrow), dtype=torch.uint8, device=blob.device)
    qual = torch.empty((R, Lrow), dtype=torch.float32, device=blob.device)
    ctx = _native.torch_context(blob.device)
    check(lib().dcb200_encode_batch_rows(ctx.handle, _p(blob), _p(seq_off), _p(qual_off), _p(lens), _p(lpad_rows), R, int(Lpad),
                                         int(Lrow), _p(tok), _p(qual)))
    return tok, qual


def _encode_rows_meta(blob, seq_off, qual_off, lens, lpad_rows, Lpad, Lrow):
    R = lens.numel()
    return blob.new_empty((R, Lrow), dtype=torch.uint8), blob.new_empty((R, Lrow), dtype=torch.float32)


def forward(tok: torch.Tensor, qual: torch.Tensor, weights: int, want_logits: bool,
            want_labels: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    B, L = tok.shape
    assert tok.dtype == torch.uint8 and qual.dtype == torch.float32 and tok.is_contiguous() and qual.is_contiguous()
    w = _WEIGHTS.get(int(weights))
    if w is None:
        raise _native.Dcb200Error(f"dcb200::forward: unknown weights handle {weights}")
    logits = torch.empty((B, L, 2) if want_logits else (0, 0, 2), dtype=torch.float32, device=tok.device)
    labels = torch.empty((B, L) if want_labels else (0, 0), dtype=torch.uint8, device=tok.device)
    ctx = _native.torch_context(tok.device)
    check(lib().dcb200_forward(ctx.handle, w.handle, _p(tok), _p(qual), B, L, _p(logits) if want_logits else None,
                               _p(labels) if want_labels else None))
    return logits, labels


def _forward_meta(tok, qual, weights, want_logits, want_labels):
    B, L = tok.shape
    return (tok.new_empty((B, L, 2) if want_logits else (0, 0, 2), dtype=torch.float32),
            tok.new_empty((B, L) if want_labels else (0, 0), dtype=torch.uint8))


def smooth_chop(labels: torch.Tensor, starts: torch.Tensor, lens: torch.Tensor, qual_lens: torch.Tensor,
                params: List[int]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``qual_lens``: int32 [R], or an EMPTY tensor for "not given"; ``params``: the 8 fields of dcb200_chop_params."""
    p = ChopParams(*[int(x) for x in params])
    R = int(lens.numel())
    ap = int(p.approved_interval_number)
    dev = labels.device
    assert starts.dtype == torch.int64 and lens.dtype == torch.int32 and labels.is_contiguous()
    n_ad = torch.zeros(R, dtype=torch.int32, device=dev)
    ad = torch.zeros((R, ap, 2), dtype=torch.int32, device=dev)
    ad_buf = ad if ap else torch.zeros(2, dtype=torch.int32, device=dev)  # approved == 0: the C ABI wants a non-null pointer
    n_keep = torch.zeros(R, dtype=torch.int32, device=dev)
    keep = torch.zeros((R, ap + 1, 2), dtype=torch.int32, device=dev)
    act = torch.zeros(R, dtype=torch.uint8, device=dev)
    ql = _p(qual_lens) if qual_lens.numel() else None
    ctx = _native.torch_context(dev)
    if labels.dtype == torch.float32:
        check(lib().dcb200_smooth_chop_logits(ctx.handle, _p(labels), labels.numel() // 2, _p(starts), _p(lens), ql, R,
                                              C.byref(p), _p(n_ad), _p(ad_buf), _p(n_keep), _p(keep), _p(act)))
    else:
        assert labels.dtype in (torch.int8, torch.uint8)
        check(lib().dcb200_smooth_chop(ctx.handle, _p(labels), labels.numel(), _p(starts), _p(lens), ql, R, C.byref(p),
                                       _p(n_ad), _p(ad_buf), _p(n_keep), _p(keep), _p(act)))
    return n_ad, ad, n_keep, keep, act


def _smooth_chop_meta(labels, starts, lens, qual_lens, params):
    R = lens.numel()
    ap = int(params[2])
    i32 = dict(dtype=torch.int32)
    return (lens.new_empty(R, **i32), lens.new_empty((R, ap, 2), **i32),
            lens.new_empty(R, **i32), lens.new_empty((R, ap + 1, 2), **i32), lens.new_empty(R, dtype=torch.uint8))


for _name, _cuda, _meta in (("encode", encode, _encode_meta), ("encode_rows", encode_rows, _encode_rows_meta),
                            ("forward", forward, _forward_meta),
                            ("smooth_chop", smooth_chop, _smooth_chop_meta)):
    _LIB.impl(_name, _cuda, "CUDA")     # CUDA only: a CPU tensor finds no kernel (no fallback)
    _LIB.impl(_name, _meta, "Meta")


def params_list(p: ChopParams) -> List[int]:
    return [int(getattr(p, name)) for name, _ in ChopParams._fields_]
