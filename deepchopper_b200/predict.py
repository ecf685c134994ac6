"""Predict pipeline: length-bucketed batches -> GPU encode -> model -> smooth/chop coordinates.

Mirror of the reference's hot loop ``trainer.predict(...)`` (deepchopper/cli.py:66-152,
deepchopper/models/basic_module.py:197-207) followed by the interval step of ``deepchopper chop``
(src/bin/predict.rs:130-192), without the ``.pt`` file hop in between (the compatible writer lives in
``deepchopper_b200.writer``).

Batching: the reference collates FASTQ-order batches and LEFT-pads to the batch maximum
(tokenizer.py:34-93); pads are semantic (no attention mask).  Here reads are sorted by length
so a batch's left pads are few; each batch is still "left-pad to the batch maximum" exactly as the
reference would collate those same reads.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native
from ._native import ChopParams, check, lib
from .encode import MAX_TOKENS

ROW_TILE = 128


@dataclass
class Batch:
    rows: np.ndarray      # indices into the read set
    Lpad: int             # collated length = max(len)+1 (reference semantics)
    Lrow: int             # row stride, multiple of 128


def plan_batches(lens: np.ndarray, token_budget: int = 1024 * 1024, max_rows: int = 8192,
                 sort: bool = True, long_read_cap: int = 128 * 32768) -> List[Batch]:
    """Length-bucketed batches of ~token_budget padded tokens.  ``sort=False`` keeps FASTQ order
    (the reference's own batching when combined with a fixed ``max_rows``).  For reads so long that the budget holds
    fewer than 128 of them the budget stretches to one full 128-row tile (at most ``long_read_cap`` tokens, ~25 GB of
    activations): the long convolution's MMAs have the batch rows as their M dimension, and a 32-row batch of 32 kb
    reads would spend 3/4 of them on nothing."""
    lens = np.minimum(np.asarray(lens, dtype=np.int64), MAX_TOKENS - 1)
    order = np.argsort(lens, kind="stable") if sort else np.arange(lens.size)
    batches: List[Batch] = []
    i = 0
    n = order.size
    while i < n:
        j = i
        mx = 0
        while j < n and (j - i) < max_rows:
            m2 = max(mx, int(lens[order[j]]))
            lrow = (m2 + 1 + ROW_TILE - 1) // ROW_TILE * ROW_TILE
            budget = max(token_budget, min(ROW_TILE * lrow, long_read_cap)) if sort else token_budget
            if j > i and (j - i + 1) * lrow > budget:
                break
            mx = m2
            j += 1
        # The long-convolution kernel's work items are (channel, 128-row tile): 256 x rows/128 items over 148 SMs.
        # Full row tiles, and a multiple of 4 of them when the batch is large (256 x 4 k / 148 = 6.92 k waves: a 1 %
        # tail instead of up to 13 %); the token budget is soft (<= 1.3 x).
        if sort and j - i > ROW_TILE and j < n:
            rows = j - i
            big = 4 * ROW_TILE
            up = (rows // big + 1) * big
            if rows >= 3 * ROW_TILE and i + up <= n and up <= max_rows:
                lrow_up = (int(lens[order[i + up - 1]]) + 1 + ROW_TILE - 1) // ROW_TILE * ROW_TILE
                if up * lrow_up <= 1.3 * token_budget:
                    rows = up
            if rows != up:
                rows = rows // big * big if rows >= big else rows // ROW_TILE * ROW_TILE
            j = i + rows
            mx = int(lens[order[j - 1]])
        lpad = mx + 1
        batches.append(Batch(order[i:j].copy(), lpad, (lpad + ROW_TILE - 1) // ROW_TILE * ROW_TILE))
        i = j
    return batches


@dataclass
class Launch:
    """Several of the reference's FASTQ-order batches in one launch (dcb200_encode_batch_rows): every row keeps the pad
    count of ITS batch -- the pads are semantic -- while the kernels see full 128-row tiles instead of 12-16 rows."""
    rows: np.ndarray      # read indices, member after member
    lpad: np.ndarray      # int32 per row: collated length of the row's own batch
    Lpad: int             # max(lpad)
    Lrow: int             # row stride, multiple of 128
    members: list         # [(position of the batch in the input list, first row, one past the last row, Batch)]


def group_batches(batches: Sequence[Batch], token_budget: int = 1024 * 1024) -> List[Launch]:
    """Pack FASTQ-order batches into launches of about ``token_budget`` padded tokens, batches of similar padded length
    together (a row is right-filled to the longest member's stride).  Results per read are those of the reference's own
    batching: the model is causal and rows are independent, so what a row sees in its first ``lpad`` columns does not
    depend on its neighbours in the launch."""
    order = sorted(range(len(batches)), key=lambda i: (batches[i].Lrow, i))
    out: List[Launch] = []
    cur: list = []
    rows_cur = 0
    lrow_cur = 0

    def flush():
        nonlocal cur, rows_cur, lrow_cur
        if not cur:
            return
        rows = np.concatenate([batches[i].rows for i in cur])
        lpad = np.concatenate([np.full(batches[i].rows.size, batches[i].Lpad, np.int32) for i in cur])
        members, r0 = [], 0
        for i in cur:
            members.append((i, r0, r0 + batches[i].rows.size, batches[i]))
            r0 += batches[i].rows.size
        out.append(Launch(rows, lpad, int(lpad.max()), lrow_cur, members))
        cur, rows_cur, lrow_cur = [], 0, 0

    lrow_first = 0
    for i in order:
        b = batches[i]
        lrow = max(lrow_cur, b.Lrow)
        # a launch closes when the budget is reached, or -- once it holds a full 128-row tile -- when the next batch is
        # more than an eighth longer than its first one (every row is right-filled to the longest member's stride)
        if cur and ((rows_cur + b.rows.size) * lrow > token_budget or
                    (rows_cur >= ROW_TILE and 8 * lrow > 9 * lrow_first)):
            flush()
            lrow = b.Lrow
        if not cur:
            lrow_first = b.Lrow
        cur.append(i)
        rows_cur += b.rows.size
        lrow_cur = lrow
    flush()
    return out


def shard_batches(batches: Sequence[Batch], rank: int, world: int) -> List[Batch]:
    """Deal batches to ranks greedily by padded token count: no collective on the path."""
    if world <= 1:
        return list(batches)
    load = [0] * world
    mine: List[Batch] = []
    for b in sorted(batches, key=lambda b: -b.rows.size * b.Lrow):
        r = int(np.argmin(load))
        load[r] += b.rows.size * b.Lrow
        if r == rank:
            mine.append(b)
    return mine


class DevicePipeline:
    """Device-resident hot path on torch CUDA tensors (inputs already in HBM)."""

    def __init__(self, model, params: Optional[ChopParams] = None):
        self.model = model
        self.params = params or ChopParams.default()
        self.device = model.device

    def upload(self, blob: np.ndarray, seq_off: np.ndarray, qual_off: np.ndarray, lens: np.ndarray,
               batches: Sequence[Batch]):
        dev = self.device
        self.blob = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
        lens = np.minimum(np.asarray(lens, dtype=np.int64), MAX_TOKENS - 1)
        self.items = []
        for b in batches:
            so = torch.from_numpy(np.ascontiguousarray(seq_off[b.rows], dtype=np.int64)).to(dev)
            qo = torch.from_numpy(np.ascontiguousarray(qual_off[b.rows], dtype=np.int64)).to(dev)
            ln = torch.from_numpy(lens[b.rows].astype(np.int32)).to(dev)
            st = torch.from_numpy((np.arange(b.rows.size, dtype=np.int64) * b.Lrow + (b.Lpad - 1) - lens[b.rows])).to(dev)
            self.items.append((b, so, qo, ln, st))
        torch.cuda.synchronize(dev)

    def upload_items(self, items):
        """``items`` = [(Batch, buf u8, seq_off, qual_off, len i32)] with offsets relative to each batch's own byte
        buffer (what bench.py synthesises per batch): one device blob, offsets rebased."""
        dev = self.device
        total = int(sum(it[1].size for it in items))
        self.blob = torch.empty(max(1, total), dtype=torch.uint8, device=dev)
        self.items = []
        base = 0
        for b, buf, so, qo, ln in items:
            self.blob[base:base + buf.size].copy_(torch.from_numpy(buf), non_blocking=False)
            lens = ln.astype(np.int64)
            st = torch.from_numpy(np.arange(b.rows.size, dtype=np.int64) * b.Lrow + (b.Lpad - 1) - lens).to(dev)
            self.items.append((b, torch.from_numpy(so + base).to(dev), torch.from_numpy(qo + base).to(dev),
                               torch.from_numpy(ln.astype(np.int32)).to(dev), st))
            base += buf.size
        torch.cuda.synchronize(dev)

    @torch.no_grad()
    def run_batch(self, item, want_logits=False):
        from .encode import encode_batch_device
        from .smooth import smooth_chop_device
        b, so, qo, ln, st = item
        ctx = _native.torch_context(self.device)
        tok, qual = encode_batch_device(self.blob, so, qo, ln, b.Lpad, ctx, b.Lrow)
        logits, labels = self.model.forward_tokens(tok, qual, want_logits, True)
        res = smooth_chop_device(labels.view(-1), st, ln, self.params, None, ctx)
        return logits, labels, res

    def run_all(self):
        out = None
        for item in self.items:
            out = self.run_batch(item)
        return out


class HostPipeline:
    """End-to-end form on HOST buffers through dcb200_predict_batch_host: per batch, H2D of the FASTQ
    bytes + offsets, encode, model, smooth/chop, D2H of the coordinate tables."""

    def __init__(self, model, params: Optional[ChopParams] = None):
        self.model = model
        self.params = params or ChopParams.default()
        self.ctx = _native.torch_context(model.device)

    def pack(self, blob: np.ndarray, seq_off, qual_off, lens, batches: Sequence[Batch], pin: bool = True):
        """Gather each batch's seq/qual strings into one contiguous pinned host buffer (ingest work,
        outside the timed hot path)."""
        lens = np.minimum(np.asarray(lens, dtype=np.int64), MAX_TOKENS - 1)
        ap = int(self.params.approved_interval_number)
        self.items = []
        for b in batches:
            ln = lens[b.rows]
            tot = int(ln.sum())
            buf = torch.empty(2 * tot, dtype=torch.uint8)
            if pin:
                buf = buf.pin_memory()
            nb = buf.numpy()
            so = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
            qo = so + tot
            for k, r in enumerate(b.rows):
                n = int(ln[k])
                nb[so[k]:so[k] + n] = blob[seq_off[r]:seq_off[r] + n]
                nb[qo[k]:qo[k] + n] = blob[qual_off[r]:qual_off[r] + n]
            R = b.rows.size
            outs = dict(n_adapter=np.zeros(R, np.int32), adapter_iv=np.zeros((R, ap, 2), np.int32),
                        n_keep=np.zeros(R, np.int32), keep_iv=np.zeros((R, ap + 1, 2), np.int32),
                        action=np.zeros(R, np.uint8))
            self.items.append((b, buf, so, qo, ln.astype(np.int32), outs))

    def pack_items(self, items, pin: bool = True):
        """Adopt per-batch byte buffers that are already packed [sequences | quality strings] (see
        DevicePipeline.upload_items): one pinned copy per batch."""
        ap = int(self.params.approved_interval_number)
        self.items = []
        for b, nb, so, qo, ln in items:
            buf = torch.empty(nb.size, dtype=torch.uint8)
            if pin:
                buf = buf.pin_memory()
            buf.numpy()[:] = nb
            R = b.rows.size
            outs = dict(n_adapter=np.zeros(R, np.int32), adapter_iv=np.zeros((R, ap, 2), np.int32),
                        n_keep=np.zeros(R, np.int32), keep_iv=np.zeros((R, ap + 1, 2), np.int32),
                        action=np.zeros(R, np.uint8))
            self.items.append((b, buf, np.ascontiguousarray(so, np.int64), np.ascontiguousarray(qo, np.int64),
                               np.ascontiguousarray(ln, np.int32), outs))

    def bytes_per_pass(self):
        h2d = sum(it[1].numel() + it[2].nbytes + it[3].nbytes + it[4].nbytes + it[0].rows.size * 8 for it in self.items)
        d2h = sum(sum(v.nbytes for v in it[5].values()) for it in self.items)
        return h2d, d2h

    def run_batch(self, item, labels_out: Optional[np.ndarray] = None, logits_out: Optional[np.ndarray] = None):
        b, buf, so, qo, ln, o = item
        p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        if isinstance(b, Launch):   # several FASTQ-order batches, every row padded as in its own batch
            check(lib().dcb200_predict_batch_host_rows(
                self.ctx.handle, self.model._weights.handle, C.c_void_p(buf.data_ptr()), buf.numel(), p(so), p(qo), p(ln),
                p(b.lpad), None, int(b.rows.size), int(b.Lpad), C.byref(self.params),
                p(logits_out) if logits_out is not None else None, p(labels_out) if labels_out is not None else None,
                p(o["n_adapter"]), p(o["adapter_iv"]), p(o["n_keep"]), p(o["keep_iv"]), p(o["action"])))
            return o
        check(lib().dcb200_predict_batch_host(
            self.ctx.handle, self.model._weights.handle, C.c_void_p(buf.data_ptr()), buf.numel(), p(so), p(qo), p(ln),
            None, int(b.rows.size), int(b.Lpad), C.byref(self.params),
            p(logits_out) if logits_out is not None else None, p(labels_out) if labels_out is not None else None,
            p(o["n_adapter"]), p(o["adapter_iv"]), p(o["n_keep"]), p(o["keep_iv"]), p(o["action"])))
        return o

    def run_all(self):
        for item in self.items:
            self.run_batch(item)
