"""Host mirror of the reference's PyO3 smoothing surface (src/python.rs:666-708,815-818 and the
``Predict`` pyclass, src/smooth/predict.rs:33-209), computing on the GPU through the C ABI.

Per-read list-in/list-out functions keep the reference's names and argument meaning; the batched
forms (``smooth_chop_host`` / ``smooth_chop_device``) are what the predict/chop pipeline uses.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._native import ChopParams, Context, check, default_context, lib

ACTION_PASSTHROUGH, ACTION_CHOP_T, ACTION_CHOP_I, ACTION_ADAPTERS, ACTION_UNCHOPPED = 0, 1, 2, 3, 4
CHOP_TYPES = {"terminal": 0, "internal": 1, "all": 2}

# src/default.rs
QUAL_OFFSET = 33
MIN_READ_LEN = 150
MIN_CHOPED_SEQ_LEN = 20
IGNORE_LABEL = -100


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


@dataclass
class ChopResult:
    n_adapter: np.ndarray   # [R] int32
    adapter_iv: np.ndarray  # [R, approved, 2] int32
    n_keep: np.ndarray      # [R] int32
    keep_iv: np.ndarray     # [R, approved+1, 2] int32
    action: np.ndarray      # [R] uint8

    def adapters(self, r: int) -> List[Tuple[int, int]]:
        return [tuple(map(int, x)) for x in self.adapter_iv[r, : self.n_adapter[r]]]

    def kept(self, r: int) -> List[Tuple[int, int]]:
        return [tuple(map(int, x)) for x in self.keep_iv[r, : self.n_keep[r]]]


def smooth_chop_host(labels: np.ndarray, starts: np.ndarray, lens: np.ndarray, params: Optional[ChopParams] = None,
                     qual_lens: Optional[np.ndarray] = None, ctx: Optional[Context] = None) -> ChopResult:
    """dcb200_smooth_chop_host: R reads of int8 labels in one host buffer -> chop coordinates."""
    ctx = ctx or default_context()
    p = params or ChopParams.default()
    labels = np.ascontiguousarray(labels, dtype=np.int8)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    if qual_lens is not None:
        qual_lens = np.ascontiguousarray(qual_lens, dtype=np.int32)
    R = int(lens.size)
    if R and (starts.min() < 0 or int((starts + lens).max()) > labels.size):
        raise ValueError("read range outside the label buffer")
    ap = int(p.approved_interval_number)
    res = ChopResult(np.zeros(R, np.int32), np.zeros((R, ap, 2), np.int32), np.zeros(R, np.int32),
                     np.zeros((R, ap + 1, 2), np.int32), np.zeros(R, np.uint8))
    check(lib().dcb200_smooth_chop_host(ctx.handle, _ptr(labels), labels.size, _ptr(starts), _ptr(lens), _ptr(qual_lens),
                                        R, C.byref(p), _ptr(res.n_adapter), _ptr(res.adapter_iv), _ptr(res.n_keep),
                                        _ptr(res.keep_iv), _ptr(res.action)))
    return res


def smooth_chop_device(labels, starts, lens, params: Optional[ChopParams] = None, qual_lens=None,
                       ctx: Optional[Context] = None, logits: bool = False):
    """Device-resident form on torch CUDA tensors (labels int8/uint8 [N] or fp32 logits [N,2] when
    ``logits=True``; starts int64 [R]; lens int32 [R]).  Returns torch tensors on the same device.
    Asynchronous on ``ctx``'s stream."""
    import torch
    from . import ops
    p = params or ChopParams.default()
    if logits:
        assert labels.dtype == torch.float32 and labels.is_contiguous()
    else:
        assert labels.dtype in (torch.int8, torch.uint8) and labels.is_contiguous()
    ql = qual_lens if qual_lens is not None else torch.empty(0, dtype=torch.int32, device=labels.device)
    return torch.ops.dcb200.smooth_chop(labels, starts, lens, ql, ops.params_list(p))


def majority_voting_host(labels: np.ndarray, starts: np.ndarray, lens: np.ndarray, window: int,
                         ctx: Optional[Context] = None) -> np.ndarray:
    ctx = ctx or default_context()
    labels = np.ascontiguousarray(labels, dtype=np.int8)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    out = labels.copy()
    check(lib().dcb200_majority_voting_host(ctx.handle, _ptr(labels), labels.size, _ptr(starts), _ptr(lens),
                                            int(lens.size), int(window), _ptr(out)))
    return out


# ---- PyO3-named single-read functions ----------------------------------------------------------

def majority_voting(labels: Sequence[int], window_size: int) -> List[int]:
    """deepchopper.majority_voting (src/python.rs:815-818 -> src/smooth/utils.rs:48-97)."""
    a = np.asarray(labels, dtype=np.int8)
    if a.size == 0:
        return []
    return majority_voting_host(a, np.array([0]), np.array([a.size]), window_size).tolist()


def _regions(labels, window, min_interval, approved) -> List[Tuple[int, int]]:
    a = np.asarray(labels, dtype=np.int8)
    if a.size == 0:
        return []
    p = ChopParams.default(smooth_window_size=window, min_interval_size=min_interval,
                           approved_interval_number=approved, min_read_length=0)
    res = smooth_chop_host(a, np.array([0]), np.array([a.size]), p)
    return res.adapters(0)


def get_label_region(labels: Sequence[int]) -> List[Tuple[int, int]]:
    """deepchopper.get_label_region (src/python.rs:666-672 -> src/utils.rs:671-695)."""
    n = len(labels)
    return _regions(labels, 1, 0, n // 2 + 1)


def smooth_label_region(labels: Sequence[int], smooth_window_size: int, min_interval_size: int,
                        approved_interval_number: int) -> List[Tuple[int, int]]:
    """deepchopper.smooth_label_region (src/python.rs:674-690 -> src/utils.rs:699-721)."""
    return _regions(labels, smooth_window_size, min_interval_size, approved_interval_number)


def generate_unmaped_intervals(intervals: Sequence[Tuple[int, int]], total_length: int) -> List[Tuple[int, int]]:
    """Host bookkeeping twin of src/output/split.rs:260-292 for caller-supplied intervals (the GPU
    path gets the same coordinates as ``keep_iv`` from dcb200_smooth_chop)."""
    if not intervals:
        return [(0, total_length)]
    out, cur = [], 0
    for s, e in intervals:
        if cur < s:
            out.append((cur, s))
        cur = e
    if cur < total_length - 1:
        out.append((cur, total_length - 1))
    return out


def remove_intervals_and_keep_left(seq: str, intervals: Sequence[Tuple[int, int]]):
    """deepchopper.remove_intervals_and_keep_left (src/python.rs:693-708 -> src/output/split.rs:295-320)."""
    ivs = sorted((tuple(i) for i in intervals), key=lambda r: r[0])
    selected = generate_unmaped_intervals(ivs, len(seq))
    for s, _ in selected:
        if s >= len(seq):
            raise ValueError(f"InvalidInterval {s}")
    return [seq[s:e] for s, e in selected], selected


def summary_predict(predictions, labels, ignore_label: int = IGNORE_LABEL):
    """deepchopper.summary_predict (src/python.rs:516-523 -> src/utils.rs:33-55): drop ignored positions."""
    outp, outl = [], []
    for p, l in zip(predictions, labels):
        p = np.asarray(p)
        l = np.asarray(l)
        keep = l != ignore_label
        outp.append(p[keep].tolist())
        outl.append(l[keep].tolist())
    return outp, outl


_ID_TABLE = np.full(256, ord("N"), dtype=np.uint8)
for _k, _v in {7: "A", 8: "C", 9: "G", 10: "T", 11: "N"}.items():
    _ID_TABLE[_k] = ord(_v)


def id_list2seq(ids: Sequence[int]) -> str:
    """deepchopper.id_list2seq (src/python.rs:810-813 -> src/smooth/utils.rs:34-46)."""
    a = np.asarray(ids).astype(np.int64)
    a = np.where((a >= 0) & (a < 256), a, 0).astype(np.uint8)
    return _ID_TABLE[a].tobytes().decode("ascii")


class Predict:
    """``deepchopper.Predict`` (src/smooth/predict.rs:33-209)."""

    def __init__(self, prediction, seq: str, id: str, is_truncated: bool, qual: Optional[str] = None):  # noqa: A002
        self.prediction = list(prediction) if not isinstance(prediction, np.ndarray) else prediction
        self.seq = seq
        self.id = id
        self.is_truncated = bool(is_truncated)
        self.qual = qual

    def __repr__(self):
        return (f"Predict(prediction: {list(self.prediction)}, seq: {self.seq}, id: {self.id}, "
                f"is_truncated: {self.is_truncated}, qual: {self.qual})")

    def prediction_region(self):
        return get_label_region(self.prediction)

    def smooth_prediction(self, window_size: int):
        return _regions(self.prediction, window_size, 0, len(self.prediction) // 2 + 1)

    def smooth_label(self, window_size: int):
        return majority_voting(self.prediction, window_size)

    def smooth_and_select_intervals(self, smooth_window_size: int, min_interval_size: int,
                                    approved_interval_number: int):
        return smooth_label_region(self.prediction, smooth_window_size, min_interval_size, approved_interval_number)

    def seq_len(self) -> int:
        return len(self.seq)

    def qual_array(self):
        return [ord(c) - QUAL_OFFSET for c in self.qual] if self.qual else []
