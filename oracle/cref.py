"""ctypes loader for the C oracle (oracle/smooth_ref.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdcref.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "smooth_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libdcref.so"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
        except subprocess.CalledProcessError:
            # toolchain without libgomp: single-threaded oracle (cpu_baseline then reports cores=1 for this part)
            os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


class CRef:
    def __init__(self, lib):
        self.lib = lib
        i8p, i32p, i64p, u8p, f32p = (C.POINTER(t) for t in (C.c_int8, C.c_int32, C.c_int64, C.c_uint8, C.c_float))
        lib.dcref_majority_voting.argtypes = [i8p, C.c_int64, C.c_int, i8p]
        lib.dcref_majority_voting.restype = None
        lib.dcref_get_label_region.argtypes = [i8p, C.c_int64, i64p, C.c_int64]
        lib.dcref_get_label_region.restype = C.c_int64
        lib.dcref_smooth_chop.argtypes = [i8p, i64p, i32p, i32p, C.c_int64] + [C.c_int] * 8 + [i32p, i32p, i32p, i32p, u8p, C.c_int]
        lib.dcref_smooth_chop.restype = C.c_int
        lib.dcref_encode_read.argtypes = [u8p, u8p, C.c_int32, C.c_int32, u8p, f32p]
        lib.dcref_encode_read.restype = None

    @staticmethod
    def _p(a, t):
        return a.ctypes.data_as(C.POINTER(t))

    def majority_voting(self, labels, window):
        a = np.ascontiguousarray(labels, dtype=np.int8)
        out = np.empty_like(a)
        self.lib.dcref_majority_voting(self._p(a, C.c_int8), a.size, int(window), self._p(out, C.c_int8))
        return out

    def get_label_region(self, labels):
        a = np.ascontiguousarray(labels, dtype=np.int8)
        out = np.empty((max(1, a.size), 2), dtype=np.int64)
        n = self.lib.dcref_get_label_region(self._p(a, C.c_int8), a.size, self._p(out, C.c_int64), out.shape[0])
        return [tuple(map(int, r)) for r in out[:n]]

    def smooth_chop(self, labels, starts, lens, qual_lens=None, window=21, min_interval=13, approved=20,
                    max_process=4, min_after_chop=20, min_read_len=150, chop_type=2, ocq=0, threads=0):
        labels = np.ascontiguousarray(labels, dtype=np.int8)
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        R = lens.size
        n_ad = np.zeros(R, np.int32)
        ad = np.zeros((R, approved, 2), np.int32)
        n_keep = np.zeros(R, np.int32)
        keep = np.zeros((R, approved + 1, 2), np.int32)
        action = np.zeros(R, np.uint8)
        ql = None
        if qual_lens is not None:
            qual_lens = np.ascontiguousarray(qual_lens, dtype=np.int32)
            ql = self._p(qual_lens, C.c_int32)
        self.lib.dcref_smooth_chop(self._p(labels, C.c_int8), self._p(starts, C.c_int64), self._p(lens, C.c_int32), ql,
                                   R, window, min_interval, approved, max_process, min_after_chop, min_read_len,
                                   chop_type, ocq, self._p(n_ad, C.c_int32), self._p(ad, C.c_int32),
                                   self._p(n_keep, C.c_int32), self._p(keep, C.c_int32), self._p(action, C.c_uint8),
                                   threads)
        return {"n_adapter": n_ad, "adapter_iv": ad, "n_keep": n_keep, "keep_iv": keep, "action": action}

    def encode_read(self, seq: bytes, qual: bytes, Lpad: int):
        s = np.frombuffer(seq, dtype=np.uint8)
        q = np.frombuffer(qual, dtype=np.uint8)
        tok = np.empty(Lpad, np.uint8)
        qq = np.empty(Lpad, np.float32)
        self.lib.dcref_encode_read(self._p(s, C.c_uint8), self._p(q, C.c_uint8), s.size, Lpad,
                                   self._p(tok, C.c_uint8), self._p(qq, C.c_float))
        return tok, qq


_cached = None


def load() -> CRef:
    global _cached
    if _cached is None:
        _cached = CRef(C.CDLL(build()))
    return _cached
