"""BASELINE configs[0] with each side using its own labels (see tests/test_gpu_parity_configs0.py): the fraction of reads
whose smoothed intervals / chop decisions equal the fp32 oracle's.   python tools/parity_configs0.py [reads]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity_configs0 import compare_configs0  # noqa: E402

print(json.dumps(compare_configs0(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)))
