"""deepchopper_b200 -- B200-native (sm_100a) implementation of DeepChopper's predict + smooth/chop hot path.

Host-side mirror of the reference's Python/PyO3 surface for that path; all compute goes through the
C ABI in include/dcb200.h (libdcb200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
from ._native import ChopParams, Context, Dcb200Error, default_context  # noqa: F401


def __getattr__(name):
    """Lazy re-exports of the reference's Python surface for this path (PyO3 module ``deepchopper.deepchopper``,
    src/python.rs:879-958, and ``deepchopper.DeepChopper``, deepchopper/models/dc_hg.py): importing the package stays
    free of torch."""
    if name in ("majority_voting", "get_label_region", "smooth_label_region", "remove_intervals_and_keep_left",
                "summary_predict", "id_list2seq", "Predict"):
        from . import smooth
        return getattr(smooth, name)
    if name in ("encode_qual", "normalize_seq"):
        from . import encode
        return getattr(encode, name)
    if name in ("load_predicts_from_batch_pt", "load_predicts_from_batch_pts", "predict_cli"):
        from . import chop
        return getattr(chop, name)
    if name in ("StatResult", "collect_statistics_for_predicts", "py_collect_statistics_for_predicts_parallel"):
        from . import stat
        return getattr(stat, name)
    if name == "DeepChopper":
        from .model import DeepChopper
        return DeepChopper
    raise AttributeError(name)
