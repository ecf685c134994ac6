"""``deepchopper chop`` driver: prediction batches + FASTQ -> chopped FASTQ (bgzip).

Host mirror of src/bin/predict.rs (CLI 19-78, process_chunk 130-192, main 197-384),
src/smooth/predict.rs:212-317 (``.pt`` loader) and src/output/split.rs:60-226 (record assembly).
The per-read arithmetic -- argmax, drop ignored positions, majority vote, intervals, chop coordinates --
runs on the GPU (dcb200_smooth_chop_logits); the host only slices strings and writes BGZF."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from ._native import ChopParams, FastqIndexC, check, lib
from .encode import FastqIndex, index_fastq, read_fastq_bytes
from .smooth import CHOP_TYPES, MIN_READ_LEN, _ID_TABLE, smooth_chop_device
from .writer import COMPACT_SUFFIX, IGNORE, list_batches, read_batch_compact


@dataclass
class ReadResult:
    seq: str                 # decoded from the prediction tensor (non-ACGT -> N), src/smooth/predict.rs:301
    n: int
    is_truncated: bool
    logits_ref: Tuple[int, int, int]   # (batch index, row, first column) to re-run with the FASTQ qual length


def write_chopped_fastq(path: str, ix: FastqIndex, has_pred: np.ndarray, pseq_ptr: np.ndarray, pseq_len: np.ndarray,
                        action: np.ndarray, n_adapter: np.ndarray, adapter_iv: np.ndarray, n_keep: np.ndarray,
                        keep_iv: np.ndarray, threads: int = 0, level: int = 6) -> Tuple[int, int]:
    """dcb200_chop_write_bgzf: record assembly + BGZF on host threads (src/bin/predict.rs:266-364,
    src/output/split.rs:60-226, src/output/writefq.rs).  All per-record arrays are in FASTQ order; ``pseq_ptr`` holds the
    address of each record's predicted sequence bytes (kept alive by the caller).  Returns (#records, #text bytes)."""
    R = len(ix)
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)  # noqa: E731
    buf = c(ix.buf, np.uint8)
    arrs = dict(name_off=c(ix.name_off, np.int64), name_len=c(ix.name_len, np.int32), head_len=c(ix.head_len, np.int32),
                seq_off=c(ix.seq_off, np.int64), seq_len=c(ix.seq_len, np.int32), qual_off=c(ix.qual_off, np.int64),
                qual_len=c(ix.qual_len, np.int32))
    cix = FastqIndexC(buf.ctypes.data, *[arrs[k].ctypes.data for k in
                                         ("name_off", "name_len", "head_len", "seq_off", "seq_len", "qual_off", "qual_len")])
    has_pred, action = c(has_pred, np.uint8), c(action, np.uint8)
    pseq_ptr, pseq_len = c(pseq_ptr, np.uint64), c(pseq_len, np.int32)
    n_adapter, n_keep = c(n_adapter, np.int32), c(n_keep, np.int32)
    adapter_iv, keep_iv = c(adapter_iv, np.int32), c(keep_iv, np.int32)
    assert adapter_iv.ndim == 3 and keep_iv.ndim == 3 and adapter_iv.shape[0] == R and keep_iv.shape[0] == R
    nrec, ntext = C.c_int64(0), C.c_int64(0)
    check(lib().dcb200_chop_write_bgzf(C.byref(cix), R, has_pred.ctypes.data, pseq_ptr.ctypes.data, pseq_len.ctypes.data,
                                       action.ctypes.data, n_adapter.ctypes.data, adapter_iv.ctypes.data,
                                       adapter_iv.shape[1], n_keep.ctypes.data, keep_iv.ctypes.data, keep_iv.shape[1],
                                       os.fsencode(path), int(threads), int(level), C.byref(nrec), C.byref(ntext)))
    return int(nrec.value), int(ntext.value)


def _read_spans(target: np.ndarray):
    """Per row: first kept column and number of kept columns (target != -100).  The collator makes the
    kept span contiguous (left pads, read, SEP)."""
    keep = target != IGNORE
    lens = keep.sum(axis=1).astype(np.int32)
    first = np.where(lens > 0, keep.argmax(axis=1), 0).astype(np.int64)
    L = target.shape[1]
    cols = np.arange(L)[None, :]
    contiguous = (keep == ((cols >= first[:, None]) & (cols < (first + lens)[:, None]))).all()
    if not contiguous:
        raise ValueError("prediction batch with non-contiguous kept positions is not supported")
    return first, lens


def load_prediction_batches(paths: Iterable[str], max_batches: Optional[int] = None):
    files: List[str] = []
    for p in paths:
        files += list_batches(p) if os.path.isdir(p) else [p]
    if max_batches is not None:
        files = files[:max_batches]
    batches = []
    for f in files:
        try:
            d = read_batch_compact(f) if f.endswith(COMPACT_SUFFIX) else torch.load(f, map_location="cpu")
        except Exception as e:  # noqa: BLE001  (src/smooth/predict.rs:246-256: report and skip)
            print(f"load pt {f} fail caused by Error: {e!r}")
            continue
        batches.append(d)
    return batches


def chop_fastq(predicts: List[str], fq: str, params: Optional[ChopParams] = None, output_prefix: Optional[str] = None,
               max_batch_size: Optional[int] = None, device: int = 0, batches=None, threads: int = 0,
               level: int = 6) -> Tuple[str, int, int]:
    """src/bin/predict.rs:197-384.  Returns (output path, #predictions, #records written)."""
    params = params or ChopParams.default()
    dev = torch.device("cuda", device)
    batches = batches if batches is not None else load_prediction_batches(predicts, max_batch_size)
    # ---- FASTQ index (id -> quality length) ------------------------------------------------------------
    buf = read_fastq_bytes(fq)
    ix = index_fastq(buf)
    fq_ids = [ix.name(r) for r in range(len(ix))]
    qlen_of = {rid: int(ix.qual_len[r]) for r, rid in enumerate(fq_ids)}
    # ---- predictions: GPU argmax + smooth + intervals + chop coordinates per batch --------------------
    R = len(ix)
    row_of = {rid: r for r, rid in enumerate(fq_ids)}
    approved = int(params.approved_interval_number)
    has_pred = np.zeros(R, np.uint8)
    action = np.zeros(R, np.uint8)
    n_ad_all = np.zeros(R, np.int32)
    n_keep_all = np.zeros(R, np.int32)
    ad_all = np.zeros((R, max(1, approved), 2), np.int32)
    keep_all = np.zeros((R, approved + 1, 2), np.int32)
    pseq_ptr = np.zeros(R, np.uint64)
    pseq_len = np.zeros(R, np.int32)
    keepalive = []
    n_pred_ids = set()
    pseq_fastq = None          # predicted sequence of compact batches = the normalised FASTQ sequence (ACGT, else N)
    for d in batches:
        if d.get("compact"):
            # compact sidecar: labels only; the sequence the reference decodes from the prediction tensor's tokens is a
            # function of the FASTQ sequence itself (tokenizer: A C G T -> 7..10, anything else -> N / UNK -> 'N')
            if pseq_fastq is None:
                table = np.full(256, ord("N"), dtype=np.uint8)
                for a, b in zip(b"ACGTacgtUu", b"ACGTACGTTT"):
                    table[a] = b
                pseq_fastq = np.ascontiguousarray(table[np.asarray(ix.buf, dtype=np.uint8)])
                keepalive.append(pseq_fastq)
            ids = d["ids"]
            offs = d["offsets"]
            lens = np.diff(offs).astype(np.int32)
            qual_lens = np.array([qlen_of.get(i, int(l)) for i, l in zip(ids, lens)], dtype=np.int32)
            n_ad, ad, n_keep, keep, act = smooth_chop_device(
                torch.from_numpy(d["labels"]).to(dev), torch.from_numpy(offs[:-1].astype(np.int64)).to(dev),
                torch.from_numpy(lens).to(dev), params, torch.from_numpy(qual_lens).to(dev), logits=False)
            n_ad, ad, n_keep, keep, act = (t.cpu().numpy() for t in (n_ad, ad, n_keep, keep, act))
            n_pred_ids.update(ids)
            for b, rid in enumerate(ids):
                r = row_of.get(rid)
                if r is None:
                    continue
                has_pred[r] = 1
                action[r] = act[b]
                n_ad_all[r] = n_ad[b]
                n_keep_all[r] = n_keep[b]
                ad_all[r, :ad.shape[1]] = ad[b]
                keep_all[r, :keep.shape[1]] = keep[b]
                pseq_ptr[r] = pseq_fastq.ctypes.data + int(ix.seq_off[r])
                pseq_len[r] = lens[b]
            continue
        pred = d["prediction"].float().contiguous()
        target = d["target"].to(torch.int64).numpy()
        seq = d["seq"].to(torch.int64).numpy()
        idarr = d["id"].to(torch.int64).numpy()
        B, L = target.shape
        first, lens = _read_spans(target)
        ids = []
        for b in range(B):
            n_id = int(idarr[b, 0])
            ids.append(bytes(idarr[b, 2:2 + n_id].astype(np.uint8)).decode("latin1"))
        qual_lens = np.array([qlen_of.get(i, int(l)) for i, l in zip(ids, lens)], dtype=np.int32)
        starts = (np.arange(B, dtype=np.int64) * L + first)
        n_ad, ad, n_keep, keep, act = smooth_chop_device(
            pred.to(dev), torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev), params,
            torch.from_numpy(qual_lens).to(dev), logits=True)
        n_ad, ad, n_keep, keep, act = (t.cpu().numpy() for t in (n_ad, ad, n_keep, keep, act))
        # sequence decoded from the prediction tensor (non-ACGT -> N), src/smooth/predict.rs:301
        letters = np.ascontiguousarray(_ID_TABLE[np.where((seq >= 0) & (seq < 256), seq, 0).astype(np.uint8)])
        keepalive.append(letters)
        base = letters.ctypes.data
        n_pred_ids.update(ids)
        for b, rid in enumerate(ids):          # a later batch overrides an earlier one, like the reference's HashMap
            r = row_of.get(rid)
            if r is None:
                continue
            has_pred[r] = 1
            action[r] = act[b]
            n_ad_all[r] = n_ad[b]
            n_keep_all[r] = n_keep[b]
            ad_all[r, :ad.shape[1]] = ad[b]
            keep_all[r, :keep.shape[1]] = keep[b]
            pseq_ptr[r] = base + int(b) * L + int(first[b])
            pseq_len[r] = lens[b]
    if len(row_of) != R:                        # duplicated FASTQ ids: every occurrence gets the id's prediction
        for r, rid in enumerate(fq_ids):
            src = row_of[rid]
            if src != r and has_pred[src]:
                has_pred[r], action[r], n_ad_all[r], n_keep_all[r] = 1, action[src], n_ad_all[src], n_keep_all[src]
                ad_all[r], keep_all[r], pseq_ptr[r], pseq_len[r] = ad_all[src], keep_all[src], pseq_ptr[src], pseq_len[src]
    # ---- FASTQ order: assemble records + BGZF on host threads (native) --------------------------------------
    if output_prefix:
        out_dir = os.path.dirname(output_prefix) or "."
        stem = output_prefix
    else:
        out_dir = os.getcwd()                                 # the reference names the output relative to the CWD
        stem = os.path.splitext(os.path.basename(fq))[0]      # Path::file_stem (src/bin/predict.rs:349)
    tmp = os.path.join(out_dir, f".deepchopper_temp_{os.getpid()}.fq.gz")
    n_out, _ = write_chopped_fastq(tmp, ix, has_pred, pseq_ptr, pseq_len, action, n_ad_all, ad_all, n_keep_all, keep_all,
                                   threads=threads, level=level)
    out = f"{stem}.{len(n_pred_ids)}pd.{n_out}record.chop.fq.gz"
    if not output_prefix:
        out = os.path.join(os.getcwd(), out) if not os.path.isabs(out) else out
    os.replace(tmp, out)
    return out, len(n_pred_ids), n_out


def params_from_cli(smooth_window=21, min_interval_size=13, approved_intervals=20, max_process_intervals=4,
                    min_read_length=20, output_chopped=False, chop_type="all") -> ChopParams:
    """Flag names of ``deepchopper chop`` (deepchopper/cli.py:155-198) -> dcb200_chop_params."""
    if chop_type not in CHOP_TYPES:
        raise ValueError("Invalid chop type")                 # src/output/split.rs:33-41
    return ChopParams.default(smooth_window_size=smooth_window, min_interval_size=min_interval_size,
                              approved_interval_number=approved_intervals, max_process_intervals=max_process_intervals,
                              min_read_length_after_chop=min_read_length, min_read_length=MIN_READ_LEN,
                              chop_type=CHOP_TYPES[chop_type], output_chopped_seqs=int(bool(output_chopped)))
