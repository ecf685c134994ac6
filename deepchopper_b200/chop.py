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


WRITE_APPEND, WRITE_NO_EOF = 1, 2     # dcb200_chop_write_bgzf_part flags


def write_chopped_fastq(path: str, ix: FastqIndex, has_pred: np.ndarray, pseq_ptr: np.ndarray, pseq_len: np.ndarray,
                        action: np.ndarray, n_adapter: np.ndarray, adapter_iv: np.ndarray, n_keep: np.ndarray,
                        keep_iv: np.ndarray, threads: int = 0, level: int = 6, flags: int = 0) -> Tuple[int, int]:
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
    check(lib().dcb200_chop_write_bgzf_part(C.byref(cix), R, has_pred.ctypes.data, pseq_ptr.ctypes.data,
                                            pseq_len.ctypes.data, action.ctypes.data, n_adapter.ctypes.data,
                                            adapter_iv.ctypes.data, adapter_iv.shape[1], n_keep.ctypes.data,
                                            keep_iv.ctypes.data, keep_iv.shape[1], os.fsencode(path), int(threads),
                                            int(level), int(flags), C.byref(nrec), C.byref(ntext)))
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


def list_prediction_files(paths: Iterable[str], max_batches: Optional[int] = None) -> List[str]:
    """``--pdt`` arguments -> prediction files: a directory is walked recursively (src/smooth/predict.rs:219-226) and
    truncated to ``max_batches`` files PER directory like ``load_predicts_from_batch_pts(.., max_predicts)``."""
    files: List[str] = []
    for p in paths:
        if os.path.isdir(p):
            fs = list_batches(p)
            files += fs[:max_batches] if max_batches is not None else fs
        else:
            files.append(p)
    return files


def iter_prediction_batches(paths: Iterable[str], max_batches: Optional[int] = None):
    """One decoded prediction batch at a time (``.pt`` dict or compact sidecar).  A generator: a batch's tensors
    (28 bytes per token in the reference's layout) are dropped before the next file is read; a file that fails to load
    is reported and skipped like src/smooth/predict.rs:246-256."""
    for f in list_prediction_files(paths, max_batches):
        try:
            d = read_batch_compact(f) if f.endswith(COMPACT_SUFFIX) else torch.load(f, map_location="cpu")
        except Exception as e:  # noqa: BLE001
            print(f"load pt {f} fail caused by Error: {e!r}")
            continue
        yield d


def load_prediction_batches(paths: Iterable[str], max_batches: Optional[int] = None):
    """Eager form of :func:`iter_prediction_batches` (small inputs / tests)."""
    return list(iter_prediction_batches(paths, max_batches))


def _decode_pt_batch(d) -> Tuple[np.ndarray, np.ndarray, List[str], np.ndarray, np.ndarray, np.ndarray]:
    """A ``.pt`` dict -> (row starts into the flat logits, kept lengths, ids, truncated flags, flat predicted-sequence
    letters, letter offsets).  Mirrors src/smooth/predict.rs:263-317 minus the argmax (fused into the GPU kernel)."""
    target = d["target"].to(torch.int64).numpy()
    idarr = d["id"].to(torch.int64).numpy()
    B, L = target.shape
    first, lens = _read_spans(target)
    n_id = np.clip(idarr[:, 0], 0, idarr.shape[1] - 2)
    idb = idarr[:, 2:].astype(np.uint8)
    ids = [idb[b, :n_id[b]].tobytes().decode("latin1") for b in range(B)]
    truncated = idarr[:, 1] != 0
    # sequence decoded from the prediction batch's tokens (non-ACGT -> N), src/smooth/predict.rs:301; only the reads'
    # own spans are kept (1 byte per base), the [B, L] int64 tensors die with `d`
    seq = d["seq"].numpy()
    cols = np.arange(L)[None, :]
    mask = (cols >= first[:, None]) & (cols < (first + lens)[:, None])
    toks = seq[mask]
    letters = np.ascontiguousarray(_ID_TABLE[np.where((toks >= 0) & (toks < 256), toks, 0).astype(np.uint8)])
    loff = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)])
    starts = np.arange(B, dtype=np.int64) * L + first
    return starts, lens, ids, truncated, letters, loff


def chop_fastq(predicts: List[str], fq: str, params: Optional[ChopParams] = None, output_prefix: Optional[str] = None,
               max_batch_size: Optional[int] = None, device: int = 0, batches=None, threads: int = 0,
               level: int = 6, suffix: str = "gz", verbose: bool = False,
               strict_ids: bool = False, chunk_bytes: Optional[int] = None) -> Tuple[str, int, int]:
    """src/bin/predict.rs:197-384.  Returns (output path, #predictions, #records written).

    Prediction files are consumed one at a time (per read only the chop decision, its <= 20 intervals and one byte per
    base of predicted sequence survive a batch); the FASTQ text is held whole, like the id -> Predict map the reference
    holds whole (src/bin/predict.rs:222-235)."""
    import time
    if chunk_bytes is None and batches is None and not strict_ids and os.path.getsize(fq) > (2 << 30):
        chunk_bytes = 256 << 20                  # large inputs are streamed (peak memory = one piece + the decision tables)
    if chunk_bytes:
        return chop_fastq_streaming(predicts, fq, params, output_prefix, max_batch_size, device, threads, level, chunk_bytes,
                                    suffix, verbose)
    t_start = time.time()
    params = params or ChopParams.default()
    dev = torch.device("cuda", device)
    it = iter(batches) if batches is not None else iter_prediction_batches(predicts, max_batch_size)
    # ---- FASTQ index (id -> row, quality length) --------------------------------------------------------
    buf = read_fastq_bytes(fq)
    ix = index_fastq(buf)
    R = len(ix)
    fq_ids = [ix.name(r) for r in range(R)]
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
    n_batches = 0
    t_index = time.time()
    t_load = t_dev = 0.0
    # ---- predictions: GPU argmax + smooth + intervals + chop coordinates, one batch at a time ---------------
    while True:
        tl = time.time()
        d = next(it, None)
        t_load += time.time() - tl
        if d is None:
            break
        n_batches += 1
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
            starts = offs[:-1].astype(np.int64)
            dev_in = torch.from_numpy(d["labels"]).to(dev)
            is_logits = False
            letters = None
        else:
            starts, lens, ids, _trunc, letters, loff = _decode_pt_batch(d)
            dev_in = d["prediction"].float().contiguous().to(dev)
            is_logits = True
            keepalive.append(letters)
        rows = np.fromiter((row_of.get(i, -1) for i in ids), dtype=np.int64, count=len(ids))
        if strict_ids and (rows < 0).any():
            raise KeyError(f"id not found: {ids[int(np.flatnonzero(rows < 0)[0])]}")      # src/cli.rs:95
        qual_lens = np.where(rows >= 0, ix.qual_len[np.maximum(rows, 0)], lens).astype(np.int32)
        td = time.time()
        n_ad, ad, n_keep, keep, act = smooth_chop_device(
            dev_in, torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev), params,
            torch.from_numpy(qual_lens).to(dev), logits=is_logits)
        n_ad, ad, n_keep, keep, act = (t.cpu().numpy() for t in (n_ad, ad, n_keep, keep, act))
        t_dev += time.time() - td
        del dev_in, d
        n_pred_ids.update(ids)
        sel = rows >= 0                        # a later batch overrides an earlier one, like the reference's HashMap
        r = rows[sel]
        has_pred[r] = 1
        action[r] = act[sel]
        n_ad_all[r] = n_ad[sel]
        n_keep_all[r] = n_keep[sel]
        if ad.shape[1]:
            ad_all[r, :ad.shape[1]] = ad[sel]
        keep_all[r, :keep.shape[1]] = keep[sel]
        if letters is None:
            pseq_ptr[r] = np.uint64(pseq_fastq.ctypes.data) + ix.seq_off[r].astype(np.uint64)
        else:
            pseq_ptr[r] = np.uint64(letters.ctypes.data) + loff[:-1][sel].astype(np.uint64)
        pseq_len[r] = lens[sel]
    if len(row_of) != R:                        # duplicated FASTQ ids: every occurrence gets the id's prediction
        src = np.fromiter((row_of[i] for i in fq_ids), dtype=np.int64, count=R)
        dup = (src != np.arange(R)) & (has_pred[src] != 0)
        for arr in (has_pred, action, n_ad_all, n_keep_all, ad_all, keep_all, pseq_ptr, pseq_len):
            arr[dup] = arr[src[dup]]
    t_pred = time.time()
    # ---- FASTQ order: assemble records + BGZF on host threads (native) --------------------------------------
    if output_prefix:
        out_dir = os.path.dirname(output_prefix) or "."
        stem = output_prefix
    else:
        out_dir = os.getcwd()                                 # the reference names the output relative to the CWD
        stem = os.path.splitext(os.path.basename(fq))[0]      # Path::file_stem (src/bin/predict.rs:349)
    tmp = os.path.join(out_dir, f".deepchopper_temp_{os.getpid()}.fq.gz")
    n_out, n_text = write_chopped_fastq(tmp, ix, has_pred, pseq_ptr, pseq_len, action, n_ad_all, ad_all, n_keep_all, keep_all,
                                        threads=threads, level=level)
    out = f"{stem}.{len(n_pred_ids)}pd.{n_out}record.chop.fq.{suffix}"
    if not output_prefix:
        out = os.path.join(os.getcwd(), out) if not os.path.isabs(out) else out
    os.replace(tmp, out)
    if verbose:                                               # src/bin/predict.rs:369-381 logs wall time and peak RSS
        import resource
        rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
        print(f"chop: {n_batches} prediction batches, {len(n_pred_ids)} predictions, {R} FASTQ records -> {n_out} records, "
              f"{n_text} text bytes; read+index FASTQ {t_index - t_start:.2f} s, load predictions {t_load:.2f} s, GPU smooth / "
              f"intervals {t_dev:.2f} s, host scatter {t_pred - t_index - t_load - t_dev:.2f} s, write {time.time() - t_pred:.2f} s, "
              f"peak RSS {rss:.0f} MB")
    return out, len(n_pred_ids), n_out


def iter_fastq_chunks(path: str, chunk_bytes: int = 256 << 20):
    """Whole-record pieces of a FASTQ file (plain or gzip / bgzip), each about ``chunk_bytes`` of text: the streaming
    reader of src/output/writefq.rs:174-193 / src/bin/predict.rs:275-316.  A piece ends after a line count that is a
    multiple of 4 (records are 4 lines, like the record index assumes)."""
    import gzip
    with open(path, "rb") as probe:
        magic = probe.read(2)
    f = gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")
    carry = b""
    lines_mod = 0            # lines of the current record already inside `carry`
    try:
        while True:
            data = f.read(chunk_bytes)
            if not data:
                break
            buf = np.frombuffer(carry + data, dtype=np.uint8)
            nl = np.flatnonzero(buf == 10)
            usable = (nl.size // 4) * 4
            if usable == 0:
                carry = buf.tobytes()
                continue
            cut = int(nl[usable - 1]) + 1
            yield buf[:cut]
            carry = buf[cut:].tobytes()
    finally:
        f.close()
    if carry.strip():
        yield np.frombuffer(carry, dtype=np.uint8)


def chop_fastq_streaming(predicts: List[str], fq: str, params: Optional[ChopParams] = None,
                         output_prefix: Optional[str] = None, max_batch_size: Optional[int] = None, device: int = 0,
                         threads: int = 0, level: int = 6, chunk_bytes: int = 256 << 20, suffix: str = "gz",
                         verbose: bool = False) -> Tuple[str, int, int]:
    """``chop`` with the FASTQ streamed in pieces like src/bin/predict.rs:275-316: the predictions are reduced to per-read
    decisions first (id -> row of compact tables, the reference's id -> Predict map), then every FASTQ piece is indexed,
    matched by id and appended to the output.  Peak memory = the decision tables + one piece, not the file.  The
    decompressed output equals :func:`chop_fastq`'s."""
    import time
    t_start = time.time()
    params = params or ChopParams.default()
    dev = torch.device("cuda", device)
    approved = int(params.approved_interval_number)
    row_of_id: Dict[str, int] = {}
    tabs = {k: [] for k in ("action", "n_ad", "ad", "n_keep", "keep", "plen", "poff")}
    letters_chunks: List[np.ndarray] = []          # predicted sequences of .pt batches (1 byte per base)
    letters_base = 0
    n_rows = 0
    for d in iter_prediction_batches(predicts, max_batch_size):
        if d.get("compact"):
            ids, offs = d["ids"], d["offsets"]
            lens = np.diff(offs).astype(np.int32)
            starts = offs[:-1].astype(np.int64)
            dev_in, is_logits = torch.from_numpy(d["labels"]).to(dev), False
            poff = np.full(len(ids), -1, np.int64)           # sequence comes from the FASTQ piece itself
        else:
            starts, lens, ids, _trunc, letters, loff = _decode_pt_batch(d)
            dev_in, is_logits = d["prediction"].float().contiguous().to(dev), True
            letters_chunks.append(letters)
            poff = letters_base + loff[:-1]
            letters_base += int(letters.size)
        # no FASTQ yet: the truncation gate (prediction length != quality length -> passthrough, src/bin/predict.rs:160-164)
        # is applied per piece below, exactly where the reference applies it
        n_ad, ad, n_keep, keep, act = (t.cpu().numpy() for t in smooth_chop_device(
            dev_in, torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev), params, None, logits=is_logits))
        del dev_in, d
        for k, rid in enumerate(ids):                          # a later batch overrides an earlier one (HashMap insert)
            row_of_id[rid] = n_rows + k
        n_rows += len(ids)
        for key, val in (("action", act), ("n_ad", n_ad), ("ad", ad), ("n_keep", n_keep), ("keep", keep), ("plen", lens),
                         ("poff", poff)):
            tabs[key].append(val)
    cat = lambda key, shape, dt: (np.concatenate(tabs[key]) if tabs[key] else np.zeros(shape, dt))  # noqa: E731
    action_p, n_ad_p, n_keep_p = cat("action", 0, np.uint8), cat("n_ad", 0, np.int32), cat("n_keep", 0, np.int32)
    ad_p, keep_p = cat("ad", (0, approved, 2), np.int32), cat("keep", (0, approved + 1, 2), np.int32)
    plen_p, poff_p = cat("plen", 0, np.int32), cat("poff", 0, np.int64)
    letters_all = np.concatenate(letters_chunks) if letters_chunks else np.zeros(1, np.uint8)
    del letters_chunks
    t_pred = time.time()
    if output_prefix:
        out_dir = os.path.dirname(output_prefix) or "."
        stem = output_prefix
    else:
        out_dir = os.getcwd()
        stem = os.path.splitext(os.path.basename(fq))[0]
    tmp = os.path.join(out_dir, f".deepchopper_temp_{os.getpid()}.fq.gz")
    open(tmp, "wb").close()
    table = np.full(256, ord("N"), dtype=np.uint8)
    for a, b in zip(b"ACGTacgtUu", b"ACGTACGTTT"):
        table[a] = b
    n_out = n_text = n_fq = n_pieces = 0
    for piece in iter_fastq_chunks(fq, chunk_bytes):
        ix = index_fastq(piece)
        R = len(ix)
        n_fq += R
        n_pieces += 1
        rows = np.fromiter((row_of_id.get(ix.name(r), -1) for r in range(R)), dtype=np.int64, count=R)
        has = rows >= 0
        pr = np.maximum(rows, 0)
        act = np.where(has, action_p[pr] if action_p.size else 0, 0).astype(np.uint8)
        plen = np.where(has, plen_p[pr] if plen_p.size else 0, 0).astype(np.int32)
        act[has & (plen != ix.qual_len)] = 0                       # truncated prediction -> passthrough
        n_ad = np.where(has, n_ad_p[pr] if n_ad_p.size else 0, 0).astype(np.int32)
        n_keep = np.where(has & (act != 0), n_keep_p[pr] if n_keep_p.size else 0, 0).astype(np.int32)
        ad = ad_p[pr] if ad_p.shape[0] else np.zeros((R, approved, 2), np.int32)
        keep = keep_p[pr] if keep_p.shape[0] else np.zeros((R, approved + 1, 2), np.int32)
        if ad.shape[1] == 0:
            ad = np.zeros((R, 1, 2), np.int32)
        pseq_piece = np.ascontiguousarray(table[piece])
        poff = poff_p[pr] if poff_p.size else np.full(R, -1, np.int64)
        ptr = np.where(poff >= 0, np.uint64(letters_all.ctypes.data) + np.maximum(poff, 0).astype(np.uint64),
                       np.uint64(pseq_piece.ctypes.data) + ix.seq_off.astype(np.uint64)).astype(np.uint64)
        nr, nt = write_chopped_fastq(tmp, ix, has.astype(np.uint8), ptr, plen, act, n_ad, ad, n_keep, keep, threads=threads,
                                     level=level, flags=WRITE_APPEND | WRITE_NO_EOF)
        n_out += nr
        n_text += nt
    # the BGZF end-of-file block
    z = np.zeros(0, np.uint8)
    end_ix = FastqIndex(z, z.astype(np.int64), z.astype(np.int32), z.astype(np.int32), z.astype(np.int64), z.astype(np.int32),
                        z.astype(np.int64), z.astype(np.int32))
    write_chopped_fastq(tmp, end_ix, z, z.astype(np.uint64), z.astype(np.int32), z, z.astype(np.int32),
                        np.zeros((0, 1, 2), np.int32), z.astype(np.int32), np.zeros((0, approved + 1, 2), np.int32),
                        threads=1, level=level, flags=WRITE_APPEND)
    out = f"{stem}.{len(row_of_id)}pd.{n_out}record.chop.fq.{suffix}"
    if not output_prefix and not os.path.isabs(out):
        out = os.path.join(os.getcwd(), out)
    os.replace(tmp, out)
    if verbose:
        import resource
        rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0
        print(f"chop (streaming): {len(row_of_id)} predictions, {n_fq} FASTQ records in {n_pieces} pieces -> {n_out} records, "
              f"{n_text} text bytes; predictions {t_pred - t_start:.2f} s, FASTQ pass {time.time() - t_pred:.2f} s, "
              f"peak RSS {rss:.0f} MB")
    return out, len(row_of_id), n_out


# ---- PyO3-named entry points (src/python.rs:879-958) ---------------------------------------------------------

def load_predicts_from_batch_pt(pt_path, ignore_label: int = IGNORE):
    """``deepchopper.load_predicts_from_batch_pt(pt_path, ignore_label)`` (src/smooth/predict.rs:263-317):
    ``{id: Predict}`` of one ``.pt`` batch -- argmax over the 2 classes (ties -> 0), positions whose target equals
    ``ignore_label`` dropped, tokens decoded to ACGTN, id and truncation flag from the ``id`` row."""
    from .smooth import Predict
    d = torch.load(os.fspath(pt_path), map_location="cpu")
    pred = d["prediction"]
    target = d["target"].to(torch.int64).numpy()
    seq = d["seq"].to(torch.int64).numpy()
    idarr = d["id"].to(torch.int64).numpy()
    lab = (pred[..., 1] > pred[..., 0]).to(torch.int8).numpy()        # argmax(2), first index wins a tie
    out = {}
    for b in range(target.shape[0]):
        keep = target[b] != ignore_label                              # summary_predict_generic (src/utils.rs:9-31)
        n_id = int(idarr[b, 0])
        rid = bytes(idarr[b, 2:2 + n_id].astype(np.uint8)).decode("latin1")
        toks = seq[b][keep]
        letters = _ID_TABLE[np.where((toks >= 0) & (toks < 256), toks, 0).astype(np.uint8)].tobytes().decode("ascii")
        out[rid] = Predict(lab[b][keep].tolist(), letters, rid, bool(idarr[b, 1] != 0), None)
    return out


def load_predicts_from_batch_pts(pt_path, ignore_label: int = IGNORE, max_predicts: Optional[int] = None):
    """``deepchopper.load_predicts_from_batch_pts(pt_path, ignore_label=-100, max_predicts=None)``
    (src/smooth/predict.rs:212-261): every ``.pt`` under ``pt_path`` (first ``max_predicts`` files), merged; a file that
    fails to load is reported and skipped."""
    files = [f for f in list_batches(os.fspath(pt_path)) if f.endswith(".pt")]
    if max_predicts is not None:
        files = files[:max_predicts]
    out = {}
    for f in files:
        try:
            out.update(load_predicts_from_batch_pt(f, ignore_label))
        except Exception as e:  # noqa: BLE001
            print(f"load pt {f} fail caused by Error: {e!r}")
    return out


def predict_cli(predicts, fq, smooth_window_size: int = 21, min_interval_size: int = 13,
                approved_interval_number: int = 20, max_process_intervals: int = 4,
                min_read_length_after_chop: int = 20, output_chopped_seqs: bool = False, chop_type: str = "all",
                threads: Optional[int] = 2, output_prefix: Optional[str] = None,
                max_batch_size: Optional[int] = None) -> None:
    """``deepchopper.predict_cli`` (src/python.rs:827-876 -> src/cli.rs:57-165): the older, non-streaming chop entry.
    Same arithmetic as ``deepchopper-chop``; the output is named ``{prefix|stem}.{n}pd.{m}record.chop.fq.bgz`` and a
    prediction whose id is not in the FASTQ is an error (``id not found``, src/cli.rs:95).  The reference emits the
    records in hash-map order; here they come out in FASTQ order."""
    p = params_from_cli(smooth_window_size, min_interval_size, approved_interval_number, max_process_intervals,
                        min_read_length_after_chop, output_chopped_seqs, chop_type)
    predicts = [os.fspath(x) for x in ([predicts] if isinstance(predicts, (str, os.PathLike)) else predicts)]
    chop_fastq(predicts, os.fspath(fq), p, output_prefix, max_batch_size, threads=threads or 0, suffix="bgz",
               strict_ids=True)


def params_from_cli(smooth_window=21, min_interval_size=13, approved_intervals=20, max_process_intervals=4,
                    min_read_length=20, output_chopped=False, chop_type="all") -> ChopParams:
    """Flag names of ``deepchopper chop`` (deepchopper/cli.py:155-198) -> dcb200_chop_params."""
    if chop_type not in CHOP_TYPES:
        raise ValueError("Invalid chop type")                 # src/output/split.rs:33-41
    return ChopParams.default(smooth_window_size=smooth_window, min_interval_size=min_interval_size,
                              approved_interval_number=approved_intervals, max_process_intervals=max_process_intervals,
                              min_read_length_after_chop=min_read_length, min_read_length=MIN_READ_LEN,
                              chop_type=CHOP_TYPES[chop_type], output_chopped_seqs=int(bool(output_chopped)))
