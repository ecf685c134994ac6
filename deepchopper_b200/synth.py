"""Seeded synthetic inputs shared by tests and bench.py."""
from __future__ import annotations

import numpy as np


def read_lengths(rng, n, median=1000, sigma=0.6, lo=200, hi=8192):
    return np.clip(np.round(rng.lognormal(np.log(median), sigma, n)), lo, hi).astype(np.int64)


def planted_labels(rng, lens):
    """Config 5: background 0 with flip noise p=0.02, 0-3 planted adapter runs (len 30-120, interior
    noise 0.05), 60 % with a terminal run.  Returns (labels int8 concatenated, starts int64, lens int32)."""
    lens = np.asarray(lens, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    total = int(lens.sum())
    lab = (rng.random(total) < 0.02).astype(np.int8)
    nrun = rng.integers(0, 4, lens.size)
    for r in range(lens.size):
        n = int(lens[r])
        for _ in range(int(nrun[r])):
            ln = int(rng.integers(30, 121))
            if n <= ln + 2:
                continue
            s = int(rng.integers(0, n - ln))
            seg = (rng.random(ln) >= 0.05).astype(np.int8)
            lab[starts[r] + s: starts[r] + s + ln] = seg
        if rng.random() < 0.6 and n > 80:
            ln = int(rng.integers(30, 81))
            lab[starts[r] + n - ln: starts[r] + n] = (rng.random(ln) >= 0.05).astype(np.int8)
    return lab, starts, lens.astype(np.int32)


def planted_labels_fast(rng, lens):
    """Vectorised variant for millions of reads (same distributional recipe, different stream)."""
    lens = np.asarray(lens, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    total = int(lens.sum())
    lab = (rng.random(total) < 0.02).astype(np.int8)
    R = lens.size
    for k in range(4):
        on = rng.random(R) < (0.6 if k == 3 else 0.5)
        ln = rng.integers(30, 121 if k < 3 else 81, R)
        ok = on & (lens > ln + 2)
        s = np.where(k == 3, lens - ln, (rng.random(R) * np.maximum(lens - ln, 1)).astype(np.int64))
        idx = np.nonzero(ok)[0]
        if idx.size == 0:
            continue
        reps = ln[idx]
        base = np.repeat(starts[idx] + s[idx], reps)
        within = np.arange(int(reps.sum())) - np.repeat(np.cumsum(reps) - reps, reps)
        pos = base + within
        lab[pos] = (rng.random(pos.size) >= 0.05).astype(np.int8)
    return lab, starts, lens.astype(np.int32)


def fastq_reads(rng, n, lo=500, hi=2000, lengths=None, n_frac=0.001):
    """Config 1/2 style reads: iid ACGT with 0.1 % N, Phred ~ clipped N(20,8) in [1,50]."""
    lengths = rng.integers(lo, hi + 1, n) if lengths is None else np.asarray(lengths)
    recs = []
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i, ln in enumerate(lengths):
        ln = int(ln)
        s = alphabet[rng.integers(0, 4, ln)].copy()
        s[rng.random(ln) < n_frac] = ord("N")
        q = np.clip(np.round(rng.normal(20, 8, ln)), 1, 50).astype(np.uint8) + 33
        recs.append((f"read_{i:07d}", s.tobytes().decode(), q.tobytes().decode()))
    return recs


def fastq_text(recs) -> bytes:
    return "".join(f"@{rid}\n{s}\n+\n{q}\n" for rid, s, q in recs).encode()
