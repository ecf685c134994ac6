"""``predict`` + ``chop`` in one pass: FASTQ file -> chopped FASTQ (bgzip), no prediction files in between.

The reference runs two programs with a directory of ``.pt`` dicts as the wire format (deepchopper/cli.py:66-198 ->
src/bin/predict.rs:197-384; 28 bytes per token on disk, re-read and argmax-ed by ``deepchopper-chop``).  Here the labels
never leave the GPU: per batch  encode -> forward (labels only) -> smooth / intervals / chop coordinates  (the three
``torch.ops.dcb200`` ops), the per-read decisions come back once at the end, and the native writer assembles the records
in FASTQ order (dcb200_chop_write_bgzf).  The output is byte-identical to ``predict --compact`` followed by ``chop``
(tests/test_gpu_pipeline.py) -- which is itself checked against the oracle's restatement of ``deepchopper-chop``.
"""
from __future__ import annotations

import os
import time
from typing import Optional, Tuple

import numpy as np
import torch

from ._native import ChopParams
from .chop import write_chopped_fastq
from .encode import MAX_TOKENS, encode_batch_device, index_fastq, read_fastq_bytes
from .predict import group_batches, plan_batches
from .smooth import smooth_chop_device


def normalised_sequence_bytes(buf: np.ndarray) -> np.ndarray:
    """The sequence ``deepchopper-chop`` decodes from a prediction batch's tokens (src/smooth/predict.rs:301) is a function
    of the FASTQ sequence itself: A C G T (any case, U -> T) survive, every other byte reads back as N."""
    table = np.full(256, ord("N"), dtype=np.uint8)
    for a, b in zip(b"ACGTacgtUu", b"ACGTACGTTT"):
        table[a] = b
    return np.ascontiguousarray(table[np.asarray(buf, dtype=np.uint8)])


@torch.no_grad()
def predict_chop_fastq(fq: str, model, params: Optional[ChopParams] = None, output_prefix: Optional[str] = None,
                       token_budget: int = 1024 * 1024, batch_size: Optional[int] = None, threads: int = 0, level: int = 6,
                       max_sample: Optional[int] = None, verbose: bool = False, preloaded=None) -> Tuple[str, int, int]:
    """FASTQ -> ``{prefix|stem}.{n_pred}pd.{n_out}record.chop.fq.gz``.  ``batch_size`` = None: length-bucketed batches of
    ~``token_budget`` padded tokens; an integer: the reference's FASTQ-order batches of that many reads (left pads are
    semantic, so the two give different logits near ties -- like any change of batch size in the reference).  Those small
    batches are packed into launches of ~``token_budget`` tokens in which every row keeps the pad count of its own batch
    (``predict.group_batches``): same inputs per read as the reference's collation, full row tiles for the kernels.
    ``preloaded`` = (bytes, FastqIndex) when the caller already read and indexed the file (the CLI does that on a thread
    while CUDA and the weights come up).  Returns (output path, #predictions, #records written)."""
    t0 = time.time()
    params = params or ChopParams.default()
    dev = model.device
    if preloaded is not None:
        buf, ix = preloaded
    else:
        buf = read_fastq_bytes(fq)
        ix = index_fastq(buf)
    R = len(ix)
    n = R if max_sample is None else min(R, int(max_sample))
    lens = np.minimum(ix.seq_len[:n].astype(np.int64), MAX_TOKENS - 1)       # tokenizer.py:154-163 (truncation)
    if batch_size is None:
        batches = plan_batches(lens, token_budget=token_budget)
        # largest batch first: the library's grow-only workspaces are sized once (a regrowth frees and reallocates GBs)
        batches = sorted(batches, key=lambda b: -b.rows.size * b.Lrow)
        launches = None
    else:
        batches = plan_batches(lens, token_budget=1 << 62, max_rows=int(batch_size), sort=False)
        launches = sorted(group_batches(batches, token_budget), key=lambda g: -g.rows.size * g.Lrow)
    t_index = time.time()
    blob = torch.from_numpy(buf).to(dev)
    torch.cuda.synchronize(dev)
    t_up = time.time()
    seq_off = torch.from_numpy(np.ascontiguousarray(ix.seq_off[:n])).to(dev)
    qual_off = torch.from_numpy(np.ascontiguousarray(ix.qual_off[:n])).to(dev)
    lens_dev = torch.from_numpy(lens.astype(np.int32)).to(dev)
    qlen_dev = torch.from_numpy(np.ascontiguousarray(ix.qual_len[:n].astype(np.int32))).to(dev)
    outs = []
    t_first = None
    from . import ops  # noqa: F401  (registers torch.ops.dcb200.*)
    for b in (launches if launches is not None else batches):
        if verbose and len(outs) == 1:
            torch.cuda.synchronize(dev)      # (verbose only) first batch = one-time costs: kernel loading, Toeplitz tables
            t_first = time.time()
        rows = torch.from_numpy(b.rows).to(dev)
        ln = lens_dev[rows]
        if launches is None:
            tok, qual = encode_batch_device(blob, seq_off[rows], qual_off[rows], ln, b.Lpad, None, b.Lrow)
            lpad64 = b.Lpad
        else:
            lpad = torch.from_numpy(b.lpad).to(dev)
            tok, qual = torch.ops.dcb200.encode_rows(blob, seq_off[rows], qual_off[rows], ln, lpad, int(b.Lpad), int(b.Lrow))
            lpad64 = lpad.to(torch.int64)
        _, labels = model.forward_tokens(tok, qual, False, True)
        starts = torch.arange(b.rows.size, dtype=torch.int64, device=dev) * b.Lrow + (lpad64 - 1) - ln.to(torch.int64)
        # a read cut to the model's window has qual_len != predicted length -> passthrough (src/bin/predict.rs:160-164)
        outs.append(smooth_chop_device(labels.view(-1), starts, ln, params, qlen_dev[rows]))
    # host work that does not need the GPU's results runs while the batches above are still executing
    pseq = normalised_sequence_bytes(ix.buf)
    pseq_ptr = (np.uint64(pseq.ctypes.data) + ix.seq_off.astype(np.uint64)).astype(np.uint64)
    pseq_len = np.zeros(R, np.int32)
    pseq_len[:n] = lens
    approved = int(params.approved_interval_number)
    has_pred = np.zeros(R, np.uint8)
    action = np.zeros(R, np.uint8)
    n_ad_all = np.zeros(R, np.int32)
    n_keep_all = np.zeros(R, np.int32)
    ad_all = np.zeros((R, max(1, approved), 2), np.int32)
    keep_all = np.zeros((R, approved + 1, 2), np.int32)
    if batches:
        order = np.concatenate([b.rows for b in (launches if launches is not None else batches)])
        n_ad, ad, n_keep, keep, act = (torch.cat([o[i] for o in outs]).cpu().numpy() for i in range(5))
        has_pred[order] = 1
        action[order] = act
        n_ad_all[order] = n_ad
        n_keep_all[order] = n_keep
        if approved:
            ad_all[order, :approved] = ad
        keep_all[order] = keep
    t_gpu = time.time()
    if output_prefix:
        out_dir = os.path.dirname(output_prefix) or "."
        stem = output_prefix
    else:
        out_dir = os.getcwd()
        stem = os.path.splitext(os.path.basename(fq))[0]
    tmp = os.path.join(out_dir, f".deepchopper_temp_{os.getpid()}.fq.gz")
    n_out, n_text = write_chopped_fastq(tmp, ix, has_pred, pseq_ptr, pseq_len, action, n_ad_all, ad_all, n_keep_all, keep_all,
                                        threads=threads, level=level)
    out = f"{stem}.{n}pd.{n_out}record.chop.fq.gz"
    if not output_prefix and not os.path.isabs(out):
        out = os.path.join(os.getcwd(), out)
    os.replace(tmp, out)
    if verbose:
        import resource
        rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
        bases = int(lens.sum())
        t1 = time.time()
        print(f"predict+chop: {n} reads, {bases} bases, {len(batches)} batches{"" if launches is None else f" in {len(launches)} launches"} -> {n_out} records ({n_text} text bytes); "
              f"read+index {t_index - t0:.2f} s, upload {t_up - t_index:.2f} s, first batch {(t_first or t_gpu) - t_up:.2f} s, "
              f"other batches + results {t_gpu - (t_first or t_gpu):.2f} s, write {t1 - t_gpu:.2f} s, "
              f"{bases / (t1 - t0) / 1e6:.1f} M bases/s wall, peak RSS {rss:.0f} MB")
    return out, n, n_out
