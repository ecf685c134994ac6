"""deepchopper_b200 -- B200-native (sm_100a) implementation of DeepChopper's predict + smooth/chop hot path.

Host-side mirror of the reference's Python/PyO3 surface for that path; all compute goes through the
C ABI in include/dcb200.h (libdcb200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401
from ._native import ChopParams, Context, Dcb200Error, default_context  # noqa: F401
