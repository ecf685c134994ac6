"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/dcb200.h declares, and fails loudly (no CPU fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from deepchopper_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "dcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    from deepchopper_b200 import _native
    names = _declared()
    assert len(names) >= 18
    lib = C.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dcb200.h but not exported"
    assert sorted(_native.EXPORTS) == names


def test_params_default_and_version(built_lib):
    from deepchopper_b200 import _native
    l = _native.lib()
    assert l.dcb200_version() >= 100
    p = _native.ChopParams()
    l.dcb200_chop_params_default(C.byref(p))
    # clap defaults, src/bin/predict.rs:31-62 + MIN_READ_LEN src/default.rs:5
    assert (p.smooth_window_size, p.min_interval_size, p.approved_interval_number, p.max_process_intervals,
            p.min_read_length_after_chop, p.min_read_length, p.chop_type, p.output_chopped_seqs) == (21, 13, 20, 4, 20, 150, 2, 0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    from deepchopper_b200 import Dcb200Error, smooth
    with pytest.raises(Dcb200Error, match="no CPU fallback|no CUDA device"):
        smooth.majority_voting([1, 0, 1], 3)


def test_host_bookkeeping_kats():
    # pure host helpers keep the reference's semantics (src/output/split.rs:326-353, src/utils.rs:742-751)
    from deepchopper_b200 import smooth
    seq = "abcdefghijklmnopqrstuvwxyz"
    assert smooth.remove_intervals_and_keep_left(seq, [(1, 5), (10, 15), (20, 25)])[0] == ["a", "fghij", "pqrst"]
    assert smooth.remove_intervals_and_keep_left(seq, [(5, 10), (15, 20)])[0] == ["abcde", "klmno", "uvwxy"]
    assert smooth.remove_intervals_and_keep_left(seq, [])[0] == [seq]
    assert smooth.generate_unmaped_intervals([(8100, 8123)], 32768) == [(0, 8100), (8123, 32767)]
    assert smooth.summary_predict([[0, 0, 1], [1, 1, 1]], [[0, -100, 1], [-100, 1, -100]], -100) == ([[0, 1], [1]], [[0, 1], [1]])
    assert smooth.id_list2seq([0, 1, 6, 7, 8, 9, 10, 11]) == "NNNACGTN"
