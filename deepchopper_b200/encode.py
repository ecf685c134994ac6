"""FASTQ ingest + GPU encoder: host mirror of parse_fastq_file (deepchopper/data/only_fq.py:21-85),
encode_qual / normalize_seq (src/python.rs:25-35,272-275), tokenize_and_align_labels_and_quals_ids
(deepchopper/models/llm/tokenizer.py:145-178) and the LEFT-padding collator (tokenizer.py:34-93).

The host only finds record boundaries (newline index); the bytes -> token / normalised-quality
work runs in dcb200_encode_batch on the GPU."""
from __future__ import annotations

import ctypes as C
import gzip
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._native import check, lib

MAX_TOKENS = 32768          # tokenizer max_length passed by the reference (tokenizer.py:96-114)
MAX_ID_LENGTH = 256         # tokenizer.py:146
PAD, SEP = 4, 1


def encode_qual(qual: str, qual_offset: int = 33) -> List[int]:
    """deepchopper.encode_qual (src/python.rs:25-35)."""
    a = np.frombuffer(qual.encode("latin1"), dtype=np.uint8).astype(np.int64) - int(qual_offset)
    if a.size and a.min() < 0:
        raise OverflowError("quality character below the offset")   # Rust u8 subtraction panics
    return a.tolist()


_NORM = np.full(256, ord("N"), dtype=np.uint8)
for _c in b"ACGTN-":
    _NORM[_c] = _c
for _a, _b in zip(b"acgtn", b"ACGTN"):
    _NORM[_a] = _b
_NORM[ord("U")] = _NORM[ord("u")] = ord("T")
_NORM[ord(".")] = _NORM[ord("~")] = ord("-")


def normalize_seq(seq: str, iupac: bool = False) -> str:
    """deepchopper.normalize_seq (src/python.rs:272-275, needletail normalize, iupac=False)."""
    if iupac:
        raise NotImplementedError("iupac=True is not on the predict path")
    a = np.frombuffer(seq.encode("latin1"), dtype=np.uint8)
    a = a[(a != 32) & (a != 9) & (a != 13) & (a != 10)]
    return _NORM[a].tobytes().decode("ascii")


@dataclass
class FastqIndex:
    """Record boundaries inside one FASTQ byte buffer (4-line records)."""
    buf: np.ndarray            # uint8, the whole text
    name_off: np.ndarray       # int64 [R] offset of the byte after '@'
    name_len: np.ndarray       # int32 [R] length of the id (up to the first blank)
    head_len: np.ndarray       # int32 [R] length of the full header line after '@'
    seq_off: np.ndarray        # int64 [R]
    seq_len: np.ndarray        # int32 [R]
    qual_off: np.ndarray       # int64 [R]
    qual_len: np.ndarray       # int32 [R]

    def __len__(self):
        return int(self.seq_off.size)

    def name(self, r: int) -> str:
        o = int(self.name_off[r])
        return self.buf[o:o + int(self.name_len[r])].tobytes().decode("ascii", "replace")

    def header(self, r: int) -> str:
        o = int(self.name_off[r])
        return self.buf[o:o + int(self.head_len[r])].tobytes().decode("ascii", "replace")

    def seq(self, r: int) -> bytes:
        o = int(self.seq_off[r])
        return self.buf[o:o + int(self.seq_len[r])].tobytes()

    def qual(self, r: int) -> bytes:
        o = int(self.qual_off[r])
        return self.buf[o:o + int(self.qual_len[r])].tobytes()


def read_fastq_bytes(path: str, threads: int = 0) -> np.ndarray:
    """Plain / gzip / bgzip FASTQ -> uint8 array (compression sniffed like src/output/writefq.rs:84-135).  Native:
    dcb200_read_file_inflate inflates BGZF blocks on host threads (a plain gzip stream is inherently one thread)."""
    import ctypes as C
    import os
    ptr, n, kind = C.c_void_p(), C.c_int64(0), C.c_int32(0)
    check(lib().dcb200_read_file_inflate(os.fsencode(path), int(threads), C.byref(ptr), C.byref(n), C.byref(kind)))
    try:
        if n.value == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()
    finally:
        lib().dcb200_free(ptr)


def index_fastq(buf: np.ndarray, threads: int = 0) -> FastqIndex:
    """Record index of a FASTQ text, built natively on host threads (dcb200_index_fastq).  Validates like
    only_fq.py:38-85: '@' headers, '+' separators, equal and non-zero sequence / quality lengths (ValueError)."""
    from ._native import Dcb200Error, FastqIndexArrays
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    arr = FastqIndexArrays()
    rc = lib().dcb200_index_fastq(C.c_void_p(buf.ctypes.data) if buf.size else None, int(buf.size), int(threads), C.byref(arr))
    if rc != 0:
        msg = lib().dcb200_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise Dcb200Error(f"libdcb200 error {rc}: {msg}")
    R = int(arr.n_records)

    def take(ptr, dtype):
        if R == 0:
            return np.zeros(0, dtype=dtype)
        ctype = C.c_int64 if dtype == np.int64 else C.c_int32
        try:
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(R,)).copy()
        finally:
            lib().dcb200_free(ptr)

    return FastqIndex(buf, take(arr.name_off, np.int64), take(arr.name_len, np.int32), take(arr.head_len, np.int32),
                      take(arr.seq_off, np.int64), take(arr.seq_len, np.int32), take(arr.qual_off, np.int64),
                      take(arr.qual_len, np.int32))


def id_rows(index: FastqIndex, rows: Sequence[int], truncated: np.ndarray) -> np.ndarray:
    """``id`` feature rows [len, truncated, ascii..., 0 pad] x 256 (tokenizer.py:169-175)."""
    out = np.zeros((len(rows), MAX_ID_LENGTH), dtype=np.int64)
    for k, r in enumerate(rows):
        n = int(index.name_len[r])
        o = int(index.name_off[r])
        row = np.concatenate([[n, int(truncated[k])], index.buf[o:o + n].astype(np.int64)])[:MAX_ID_LENGTH]
        out[k, : row.size] = row
    return out


def encode_batch_device(bytes_dev, seq_off_dev, qual_off_dev, len_dev, Lpad: int, ctx=None, Lrow: int | None = None):
    """dcb200_encode_batch on torch CUDA tensors -> (tok uint8 [R,Lpad], qual float32 [R,Lpad])."""
    import torch
    from . import ops  # noqa: F401  (registers torch.ops.dcb200.*)
    Lrow = int(Lrow or (Lpad + 3) // 4 * 4)
    return torch.ops.dcb200.encode(bytes_dev, seq_off_dev, qual_off_dev, len_dev, int(Lpad), Lrow)


def encode_records(recs: Sequence[Tuple[str, str, str]], Lpad: int, device="cuda"):
    """Convenience: [(id, seq, qual)] -> device tensors through the GPU encoder."""
    import torch
    blob = bytearray()
    so, qo, ln = [], [], []
    for _, s, q in recs:
        if len(s) != len(q):
            raise ValueError("sequence and quality lengths differ")
        n = min(len(s), MAX_TOKENS - 1)
        so.append(len(blob))
        blob += s.encode("latin1")
        qo.append(len(blob))
        blob += q.encode("latin1")
        ln.append(n)
    if max(ln) + 1 > Lpad:
        raise ValueError("Lpad too small")
    dev = torch.device(device)
    b = torch.frombuffer(bytes(blob), dtype=torch.uint8).to(dev)
    return encode_batch_device(b, torch.tensor(so, dtype=torch.int64, device=dev),
                               torch.tensor(qo, dtype=torch.int64, device=dev),
                               torch.tensor(ln, dtype=torch.int32, device=dev), Lpad)
