"""ctypes binding of libdcb200.so (include/dcb200.h).  There is no CPU fallback: importing this module
never touches the GPU, but every compute call fails loudly if the library or a B200 is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCB200_LIB") or os.path.join(_HERE, "lib", "libdcb200.so")   # (override: kernel experiments)

EXPORTS = [
    "dcb200_last_error", "dcb200_version", "dcb200_chop_params_default", "dcb200_ctx_create", "dcb200_ctx_destroy",
    "dcb200_ctx_sync", "dcb200_ctx_stream", "dcb200_ctx_launch_count", "dcb200_encode_batch", "dcb200_encode_batch_rows", "dcb200_weights_create",
    "dcb200_weights_destroy", "dcb200_forward", "dcb200_smooth_chop", "dcb200_smooth_chop_logits",
    "dcb200_majority_voting", "dcb200_smooth_chop_host", "dcb200_majority_voting_host", "dcb200_predict_batch_host", "dcb200_predict_batch_host_rows",
    "dcb200_forward_debug", "dcb200_ctx_read_workspace", "dcb200_ctx_profile", "dcb200_ctx_profile_read",
    "dcb200_kernel_kind_name", "dcb200_chop_write_bgzf",
    "dcb200_read_file_inflate", "dcb200_free", "dcb200_ctx_set_option", "dcb200_ctx_get_option",
    "dcb200_index_fastq", "dcb200_chop_write_bgzf_part",
]


class ChopParams(C.Structure):
    """dcb200_chop_params (clap defaults of deepchopper-chop, src/bin/predict.rs:19-78)."""
    _fields_ = [("smooth_window_size", C.c_int32), ("min_interval_size", C.c_int32),
                ("approved_interval_number", C.c_int32), ("max_process_intervals", C.c_int32),
                ("min_read_length_after_chop", C.c_int32), ("min_read_length", C.c_int32),
                ("chop_type", C.c_int32), ("output_chopped_seqs", C.c_int32)]

    @classmethod
    def default(cls, **kw) -> "ChopParams":
        p = cls(21, 13, 20, 4, 20, 150, 2, 0)
        for k, v in kw.items():
            setattr(p, k, int(v))
        return p


class FastqIndexC(C.Structure):
    """dcb200_fastq_index."""
    _fields_ = [("fastq", C.c_void_p), ("name_off", C.c_void_p), ("name_len", C.c_void_p), ("head_len", C.c_void_p),
                ("seq_off", C.c_void_p), ("seq_len", C.c_void_p), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p)]


class FastqIndexArrays(C.Structure):
    """dcb200_fastq_index_arrays (malloc'ed by dcb200_index_fastq)."""
    _fields_ = [("n_records", C.c_int64), ("name_off", C.c_void_p), ("name_len", C.c_void_p), ("head_len", C.c_void_p),
                ("seq_off", C.c_void_p), ("seq_len", C.c_void_p), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p)]


class Dcb200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Dcb200Error(f"{LIB_PATH} is missing: run `python -m deepchopper_b200.build` (nvcc, sm_100a). "
                          "deepchopper_b200 has no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    l.dcb200_last_error.restype = C.c_char_p
    l.dcb200_version.restype = C.c_int
    l.dcb200_chop_params_default.argtypes = [C.POINTER(ChopParams)]
    l.dcb200_chop_params_default.restype = None
    l.dcb200_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    l.dcb200_ctx_destroy.argtypes = [vp]
    l.dcb200_ctx_sync.argtypes = [vp]
    l.dcb200_ctx_stream.argtypes = [vp]
    l.dcb200_ctx_stream.restype = vp
    l.dcb200_ctx_launch_count.argtypes = [vp]
    l.dcb200_ctx_launch_count.restype = i64
    l.dcb200_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
    l.dcb200_ctx_get_option.argtypes = [vp, C.c_char_p]
    l.dcb200_ctx_get_option.restype = i64
    l.dcb200_encode_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    l.dcb200_encode_batch_rows.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    l.dcb200_weights_create.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i64), i32, C.POINTER(vp)]
    l.dcb200_weights_destroy.argtypes = [vp]
    l.dcb200_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    l.dcb200_forward_debug.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, i32]
    l.dcb200_ctx_read_workspace.argtypes = [vp, C.c_char_p, vp, i64]
    l.dcb200_ctx_profile.argtypes = [vp, C.c_int]
    l.dcb200_ctx_profile_read.argtypes = [vp, vp, vp, i32, i32]
    l.dcb200_kernel_kind_name.argtypes = [i32]
    l.dcb200_kernel_kind_name.restype = C.c_char_p
    pp = C.POINTER(ChopParams)
    l.dcb200_read_file_inflate.argtypes = [C.c_char_p, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32)]
    l.dcb200_index_fastq.argtypes = [vp, i64, i32, C.POINTER(FastqIndexArrays)]
    l.dcb200_free.argtypes = [vp]
    l.dcb200_free.restype = None
    l.dcb200_chop_write_bgzf.argtypes = [C.POINTER(FastqIndexC), i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, C.c_char_p,
                                         i32, i32, vp, vp]
    l.dcb200_chop_write_bgzf_part.argtypes = [C.POINTER(FastqIndexC), i64, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32,
                                              C.c_char_p, i32, i32, i32, vp, vp]
    l.dcb200_smooth_chop.argtypes = [vp, vp, i64, vp, vp, vp, i64, pp, vp, vp, vp, vp, vp]
    l.dcb200_smooth_chop_logits.argtypes = [vp, vp, i64, vp, vp, vp, i64, pp, vp, vp, vp, vp, vp]
    l.dcb200_majority_voting.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
    l.dcb200_smooth_chop_host.argtypes = [vp, vp, i64, vp, vp, vp, i64, pp, vp, vp, vp, vp, vp]
    l.dcb200_majority_voting_host.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
    l.dcb200_predict_batch_host.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, pp, vp, vp, vp, vp, vp, vp, vp]
    l.dcb200_predict_batch_host_rows.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, i32, i32, pp, vp, vp, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        if name not in ("dcb200_last_error", "dcb200_chop_params_default", "dcb200_ctx_stream", "dcb200_ctx_launch_count",
                        "dcb200_kernel_kind_name", "dcb200_ctx_get_option", "dcb200_free"):
            getattr(l, name).restype = C.c_int
    _lib = l
    return l


def check(rc: int):
    if rc != 0:
        msg = lib().dcb200_last_error().decode("utf-8", "replace")
        raise Dcb200Error(f"libdcb200 error {rc}: {msg}")


class Context:
    """One device + one stream (dcb200_ctx).  stream=None -> library-owned stream; pass
    ``torch.cuda.current_stream().cuda_stream`` to share torch's."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        check(lib().dcb200_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(lib().dcb200_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        return int(lib().dcb200_ctx_stream(self._h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().dcb200_ctx_launch_count(self._h))

    def set_option(self, name: str, value: int):
        """dcb200_ctx_set_option, e.g. ``fft_min_len`` (which long-convolution kernel a batch length takes)."""
        check(lib().dcb200_ctx_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return int(lib().dcb200_ctx_get_option(self._h, name.encode()))

    def profile(self, enable: bool = True):
        check(lib().dcb200_ctx_profile(self._h, 1 if enable else 0))

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel kind: (total ms, launches)} measured with CUDA events on this ctx's stream."""
        n = 16
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        check(lib().dcb200_ctx_profile_read(self._h, ms, cnt, n, 1 if reset else 0))
        out = {}
        for i in range(n):
            name = lib().dcb200_kernel_kind_name(i)
            if name and cnt[i]:
                out[name.decode()] = (float(ms[i]), int(cnt[i]))
        return out

    def close(self):
        if self._h:
            lib().dcb200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


_torch_ctx: dict = {}


def torch_context(device=None) -> Context:
    """A Context bound to torch's CURRENT stream on ``device`` (so library kernels are ordered with
    torch's own allocations/copies).  Cached per (device, stream)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    s = torch.cuda.current_stream(idx).cuda_stream or 1   # 0x1 == cudaStreamLegacy
    key = (idx, s)
    if key not in _torch_ctx:
        _torch_ctx[key] = Context(idx, s)
    return _torch_ctx[key]
