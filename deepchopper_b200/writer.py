"""Prediction directory wire format between ``predict`` and ``chop``.

``CustomWriter.write_on_batch_end`` (deepchopper/models/callbacks.py:12-25) saves, per batch,
``{out}/{dataloader_idx}/{global_rank}_{batch_idx}.pt`` = ``torch.save`` of
``{"prediction": f32[B,L,2], "target": i64[B,L], "seq": i64[B,L], "qual": f32[B,L], "id": i64[B,256]}``,
which the Rust loader reads back (src/smooth/predict.rs:263-317: needs float / int64 storages)."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

IGNORE = -100


def batch_dict(logits: torch.Tensor, tok: torch.Tensor, qual: torch.Tensor, id_rows: np.ndarray, lens: np.ndarray,
               Lpad: int) -> Dict[str, torch.Tensor]:
    """Assemble the reference's per-batch dict from device outputs (columns >= Lpad are row filler)."""
    B = tok.shape[0]
    pred = logits[:, :Lpad, :].float().cpu().contiguous()
    seq = tok[:, :Lpad].to(torch.int64).cpu().contiguous()
    q = qual[:, :Lpad].float().cpu().contiguous()
    target = torch.full((B, Lpad), IGNORE, dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        target[b, Lpad - 1 - n:Lpad - 1] = 0          # labels = zeros + [-100] (tokenizer.py:160-166), left pad -100
    return {"prediction": pred, "target": target, "seq": seq, "qual": q,
            "id": torch.from_numpy(np.ascontiguousarray(id_rows, dtype=np.int64))}


def write_batch(output_dir: str, rank: int, batch_idx: int, d: Dict[str, torch.Tensor], dataloader_idx: int = 0) -> str:
    folder = os.path.join(output_dir, str(dataloader_idx))
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, f"{rank}_{batch_idx}.pt")
    torch.save(d, path)
    return path


COMPACT_SUFFIX = ".dcl.npz"


def write_batch_compact(output_dir: str, rank: int, batch_idx: int, labels, lens: np.ndarray, Lpad: int, ids,
                        truncated: np.ndarray, dataloader_idx: int = 0) -> str:
    """Compact sidecar: what ``chop`` needs of a batch and nothing else -- the per-base labels of every
    read (argmax already taken on the GPU), bit-packed, plus ids and the truncation flags: 1 bit per base instead of the
    28 bytes per token of the reference's ``.pt`` dict.  Only this package's ``chop`` reads it; the ``.pt`` layout stays
    the default.  ``labels`` is the u8 [B, >= Lpad] label matrix of the batch (torch tensor on any device, or numpy)."""
    lab = labels[:, :Lpad]
    lab = lab.cpu().numpy() if hasattr(lab, "cpu") else np.asarray(lab)
    lens = np.asarray(lens, dtype=np.int64)
    cols = np.arange(Lpad)[None, :]
    mask = (cols >= (Lpad - 1 - lens)[:, None]) & (cols < Lpad - 1)          # the read's span: left pads | read | SEP
    flat = (lab[mask] != 0)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idb = [i.encode("latin1") if isinstance(i, str) else bytes(i) for i in ids]
    id_off = np.concatenate([[0], np.cumsum([len(b) for b in idb])]).astype(np.int64)
    folder = os.path.join(output_dir, str(dataloader_idx))
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, f"{rank}_{batch_idx}{COMPACT_SUFFIX}")
    with open(path, "wb") as f:
        np.savez(f, bits=np.packbits(flat), offsets=offsets, id_bytes=np.frombuffer(b"".join(idb), dtype=np.uint8),
                 id_offsets=id_off, truncated=np.asarray(truncated, dtype=np.uint8))
    return path


def read_batch_compact(path: str) -> Dict[str, np.ndarray]:
    z = np.load(path)
    offsets = z["offsets"]
    labels = np.unpackbits(z["bits"])[: int(offsets[-1])].astype(np.int8)
    idb, ido = z["id_bytes"].tobytes(), z["id_offsets"]
    ids = [idb[ido[k]:ido[k + 1]].decode("latin1") for k in range(len(ido) - 1)]
    return {"compact": True, "labels": labels, "offsets": offsets, "ids": ids, "truncated": z["truncated"]}


def list_batches(path: str):
    """All ``.pt`` files (and compact sidecars) under a prediction directory (src/smooth/predict.rs:219-226 walks
    recursively)."""
    out = []
    for root, _, files in os.walk(path):
        for f in files:
            if f.endswith(".pt") or f.endswith(COMPACT_SUFFIX):
                out.append(os.path.join(root, f))
    return sorted(out)
