"""``deepchopper chop`` driver: prediction batches + FASTQ -> chopped FASTQ (bgzip).

Host mirror of src/bin/predict.rs (CLI 19-78, process_chunk 130-192, main 197-384),
src/smooth/predict.rs:212-317 (``.pt`` loader) and src/output/split.rs:60-226 (record assembly).
The per-read arithmetic -- argmax, drop ignored positions, majority vote, intervals, chop coordinates --
runs on the GPU (dcb200_smooth_chop_logits); the host only slices strings and writes BGZF."""
from __future__ import annotations

import os
import struct
import zlib
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from ._native import ChopParams
from .encode import FastqIndex, index_fastq, read_fastq_bytes
from .smooth import (ACTION_ADAPTERS, ACTION_CHOP_I, ACTION_CHOP_T, ACTION_PASSTHROUGH, ACTION_UNCHOPPED, CHOP_TYPES, MIN_READ_LEN,
                     id_list2seq, smooth_chop_device)
from .writer import IGNORE, list_batches


@dataclass
class ReadResult:
    seq: str                 # decoded from the prediction tensor (non-ACGT -> N), src/smooth/predict.rs:301
    n: int
    is_truncated: bool
    logits_ref: Tuple[int, int, int]   # (batch index, row, first column) to re-run with the FASTQ qual length


class _BgzfWriter:
    """Minimal BGZF (blocked gzip, htslib-compatible) writer: <=64 KiB blocks + the 28-byte EOF block."""
    EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")

    def __init__(self, path: str, level: int = 6):
        self.f = open(path, "wb")
        self.buf = bytearray()
        self.level = level

    def write(self, data: bytes):
        self.buf += data
        while len(self.buf) >= 0xff00:
            self._block(bytes(self.buf[:0xff00]))
            del self.buf[:0xff00]

    def _block(self, data: bytes):
        c = zlib.compressobj(self.level, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        bsize = len(comp) + 25
        self.f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize))
        self.f.write(comp)
        self.f.write(struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    def close(self):
        if self.buf:
            self._block(bytes(self.buf))
            self.buf.clear()
        self.f.write(self.EOF)
        self.f.close()


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
            d = torch.load(f, map_location="cpu")
        except Exception as e:  # noqa: BLE001  (src/smooth/predict.rs:246-256: report and skip)
            print(f"load pt {f} fail caused by Error: {e!r}")
            continue
        batches.append(d)
    return batches


def chop_fastq(predicts: List[str], fq: str, params: Optional[ChopParams] = None, output_prefix: Optional[str] = None,
               max_batch_size: Optional[int] = None, device: int = 0, batches=None) -> Tuple[str, int, int]:
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
    results: Dict[str, tuple] = {}
    for d in batches:
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
        for b in range(B):
            s = id_list2seq(seq[b, first[b]:first[b] + lens[b]])
            results[ids[b]] = (int(act[b]), ad[b, :n_ad[b]].tolist(), keep[b, :n_keep[b]].tolist(), s)
    # ---- stream the FASTQ in order, assemble records -------------------------------------------------------
    if output_prefix:
        out_dir = os.path.dirname(output_prefix) or "."
        stem = output_prefix
    else:
        out_dir = os.getcwd()                                 # the reference names the output relative to the CWD
        stem = os.path.splitext(os.path.basename(fq))[0]      # Path::file_stem (src/bin/predict.rs:349)
    tmp = os.path.join(out_dir, f".deepchopper_temp_{os.getpid()}.fq.gz")
    w = _BgzfWriter(tmp)
    n_out = 0
    for r, rid in enumerate(fq_ids):
        res = results.get(rid)
        if res is None:
            continue                                          # no prediction -> dropped (src/bin/predict.rs:141-144)
        act, adapters, kept, pseq = res
        qual = ix.qual(r).decode("latin1")
        if act == ACTION_PASSTHROUGH:
            w.write(f"@{ix.header(r)}\n{ix.seq(r).decode('latin1')}\n+\n{qual}\n".encode("latin1"))
            n_out += 1
        elif act == ACTION_UNCHOPPED:
            w.write(f"@{rid}\n{pseq}\n+\n{qual}\n".encode("latin1"))
            n_out += 1
        elif act == ACTION_ADAPTERS:
            for s, e in adapters:
                w.write(f"@{rid}|{s}:{e}\n{pseq[s:e]}\n+\n{qual[s:e]}\n".encode("latin1"))
                n_out += 1
        else:
            tag = "T" if act == ACTION_CHOP_T else "I"
            for s, e in kept:
                w.write(f"@{rid}|{s}:{e}|{tag}\n{pseq[s:e]}\n+\n{qual[s:e]}\n".encode("latin1"))
                n_out += 1
    w.close()
    out = f"{stem}.{len(results)}pd.{n_out}record.chop.fq.gz"
    if not output_prefix:
        out = os.path.join(os.getcwd(), out) if not os.path.isabs(out) else out
    os.replace(tmp, out)
    return out, len(results), n_out


def params_from_cli(smooth_window=21, min_interval_size=13, approved_intervals=20, max_process_intervals=4,
                    min_read_length=20, output_chopped=False, chop_type="all") -> ChopParams:
    """Flag names of ``deepchopper chop`` (deepchopper/cli.py:155-198) -> dcb200_chop_params."""
    if chop_type not in CHOP_TYPES:
        raise ValueError("Invalid chop type")                 # src/output/split.rs:33-41
    return ChopParams.default(smooth_window_size=smooth_window, min_interval_size=min_interval_size,
                              approved_interval_number=approved_intervals, max_process_intervals=max_process_intervals,
                              min_read_length_after_chop=min_read_length, min_read_length=MIN_READ_LEN,
                              chop_type=CHOP_TYPES[chop_type], output_chopped_seqs=int(bool(output_chopped)))
