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


def list_batches(path: str):
    """All ``.pt`` files under a prediction directory (src/smooth/predict.rs:219-226 walks recursively)."""
    out = []
    for root, _, files in os.walk(path):
        for f in files:
            if f.endswith(".pt"):
                out.append(os.path.join(root, f))
    return sorted(out)
