"""CPU check of the blocked FFT long convolution (deepchopper_b200/csrc/lconv_core.cuh): tests/native/lconv_check.cu
compiles the kernel's own thread-level functions for the HOST, emulates one CTA thread by thread and compares with a
direct O(L^2) causal convolution in double at L = 128 ... 32768 (1 to 4 blocks, partial last blocks)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_lconv_host_emulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "lconv_check")
    subprocess.check_call([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "--extended-lambda", "--expt-relaxed-constexpr", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "lconv_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")
