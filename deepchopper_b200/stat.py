"""``deepchopper.StatResult`` and ``py_collect_statistics_for_predicts_parallel`` (src/smooth/stat.rs:16-308; PyO3
registration src/python.rs:934,954; used by scripts/eval_with_bam.py:620 and src/smooth/strategy.rs:314).

The smoothing / interval selection of every read runs in ONE batched GPU call (dcb200_smooth_chop_host: the same
kernel as the chop path); the raw label runs (``Predict.prediction_region``) are a vectorised scan of the concatenated
labels; what is left on the host is list building.  The rayon reduce of the reference merges per-read results in input
order, so every list below is in input order too."""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from ._native import ChopParams
from .smooth import MIN_READ_LEN, smooth_chop_host

FLANK_SIZE_COUNT_PLOYA = 5   # src/smooth/stat.rs:16


@dataclass
class StatResult:
    """src/smooth/stat.rs:19-41 (same field names; methods 74-130, merge 186-203, JSON state 67-72,150-176)."""
    predicts_with_chop: List[str] = field(default_factory=list)
    smooth_predicts_with_chop: List[str] = field(default_factory=list)
    smooth_internal_predicts: List[str] = field(default_factory=list)
    smooth_intervals: Dict[str, List[Tuple[int, int]]] = field(default_factory=dict)
    original_intervals: Dict[str, List[Tuple[int, int]]] = field(default_factory=dict)
    total_truncated: int = 0
    smooth_only_one: List[str] = field(default_factory=list)
    smooth_only_one_with_ploya: List[str] = field(default_factory=list)
    total_predicts: int = 0
    smooth_intervals_relative_pos: List[float] = field(default_factory=list)

    @classmethod
    def from_json(cls, json_path: str) -> "StatResult":
        with open(json_path) as f:
            d = json.loads("".join(line.rstrip("\n") for line in f))
        d["smooth_intervals"] = {k: [tuple(r) for r in v] for k, v in d["smooth_intervals"].items()}
        d["original_intervals"] = {k: [tuple(r) for r in v] for k, v in d["original_intervals"].items()}
        return cls(**d)

    def to_json(self) -> str:
        return json.dumps(asdict(self))

    def selected_predict_by_intervals(self, interval_number: int) -> List[str]:
        return [i for i in self.smooth_predicts_with_chop if len(self.smooth_intervals[i]) >= interval_number]

    def length_predicts_with_chop(self) -> List[int]:
        return [e - s for i in self.predicts_with_chop for s, e in self.original_intervals[i]]

    def number_predicts_with_chop(self) -> List[int]:
        return [len(self.original_intervals[i]) for i in self.predicts_with_chop]

    def lenghth_smooth_predicts_with_chop(self) -> List[int]:   # (sic, the reference's spelling)
        return [e - s for i in self.smooth_predicts_with_chop for s, e in self.smooth_intervals[i]]

    def number_smooth_predicts_with_chop(self) -> List[int]:
        return [len(self.smooth_intervals[i]) for i in self.smooth_predicts_with_chop]

    def merge(self, other: "StatResult") -> None:
        self.predicts_with_chop.extend(other.predicts_with_chop)
        self.smooth_predicts_with_chop.extend(other.smooth_predicts_with_chop)
        self.smooth_internal_predicts.extend(other.smooth_internal_predicts)
        self.smooth_intervals.update(other.smooth_intervals)
        self.original_intervals.update(other.original_intervals)
        self.total_truncated += other.total_truncated
        self.smooth_only_one.extend(other.smooth_only_one)
        self.smooth_only_one_with_ploya.extend(other.smooth_only_one_with_ploya)
        self.total_predicts += other.total_predicts
        self.smooth_intervals_relative_pos.extend(other.smooth_intervals_relative_pos)

    def __repr__(self) -> str:
        return (f"StatResult(total_predicts: {self.total_predicts},  predicts_with_chop: {len(self.predicts_with_chop)}, "
                f"smooth_predicts_with_chop: {len(self.smooth_predicts_with_chop)}, smooth_internal_predicts: "
                f"{len(self.smooth_internal_predicts)}, total_truncated: {self.total_truncated}, smooth_only_one: "
                f"{len(self.smooth_only_one)}, smooth_ploya_only_one: {len(self.smooth_only_one_with_ploya)})")

    def __getstate__(self):
        return self.to_json().encode()

    def __setstate__(self, state):
        d = json.loads(state.decode())
        d["smooth_intervals"] = {k: [tuple(r) for r in v] for k, v in d["smooth_intervals"].items()}
        d["original_intervals"] = {k: [tuple(r) for r in v] for k, v in d["original_intervals"].items()}
        self.__dict__.update(d)


def label_regions_batch(labels: np.ndarray, starts: np.ndarray, lens: np.ndarray) -> List[List[Tuple[int, int]]]:
    """get_label_region (src/utils.rs:671-695) of every read of a concatenated int8 label array (numpy per read), including
    the ``start == 0`` sentinel quirk: a run that begins at a read's index 0 loses its first base (and vanishes if it
    was one base long)."""
    R = len(lens)
    out: List[List[Tuple[int, int]]] = [[] for _ in range(R)]
    if R == 0:
        return out
    lab = np.asarray(labels)
    for r in range(R):
        s, n = int(starts[r]), int(lens[r])
        a = (lab[s:s + n] == 1).astype(np.int8)
        if n == 0 or not a.any():
            continue
        dd = np.diff(np.r_[np.int8(0), a, np.int8(0)])
        b = np.flatnonzero(dd == 1)
        e = np.flatnonzero(dd == -1)
        if b.size and b[0] == 0:
            b = b.copy()
            b[0] = 1
        keep = b < e
        out[r] = list(zip(b[keep].tolist(), e[keep].tolist()))
    return out


def collect_statistics_for_predicts(predicts: Sequence, smooth_window_size: int, min_interval_size: int,
                                    approved_interval_number: int, internal_threshold: float,
                                    ploya_threshold: int) -> StatResult:
    """deepchopper.py_collect_statistics_for_predicts_parallel (src/smooth/stat.rs:205-308)."""
    res = StatResult()
    sel = [p for p in predicts if len(p.seq) >= MIN_READ_LEN]
    if not sel:
        return res
    lens = np.array([len(p.prediction) for p in sel], dtype=np.int32)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    labels = np.concatenate([np.asarray(p.prediction, dtype=np.int8) for p in sel]) if int(lens.sum()) else np.zeros(0, np.int8)
    params = ChopParams.default(smooth_window_size=smooth_window_size, min_interval_size=min_interval_size,
                                approved_interval_number=approved_interval_number, min_read_length=0)
    sm = smooth_chop_host(labels, starts, lens, params)         # one GPU call for every read
    raw = label_regions_batch(labels, starts, lens)
    thr = np.float32(internal_threshold)
    for k, p in enumerate(sel):
        res.total_predicts += 1
        if p.is_truncated:
            res.total_truncated += 1
        if raw[k]:
            res.predicts_with_chop.append(p.id)
            res.original_intervals[p.id] = raw[k]
        smooth = sm.adapters(k)
        if smooth:
            res.smooth_predicts_with_chop.append(p.id)
            res.smooth_intervals[p.id] = list(smooth)
            if len(smooth) == 1:
                res.smooth_only_one.append(p.id)
                s0 = smooth[0][0]
                if p.seq[max(0, s0 - FLANK_SIZE_COUNT_PLOYA):s0].count("A") >= ploya_threshold:
                    res.smooth_only_one_with_ploya.append(p.id)
            for (_, e) in smooth:
                rel = np.float32(e) / np.float32(len(p.seq))
                res.smooth_intervals_relative_pos.append(float(rel))
                if rel < thr:
                    res.smooth_internal_predicts.append(p.id)
    return res


py_collect_statistics_for_predicts_parallel = collect_statistics_for_predicts
