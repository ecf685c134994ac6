"""Dump a kernel timeline trace for one big batch: DCB200_TRACE=block|inproj|toeplitz python tools/trace_kernel.py
(the tracer is compiled in; the env var is read once when the ctx is created)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["DCB200_TRACE"] = os.environ.get("DCB200_TRACE", "block")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepchopper_b200._native import check, lib  # noqa: E402
from deepchopper_b200.init_weights import random_state_dict  # noqa: E402
from deepchopper_b200.model import DeepChopper  # noqa: E402

B, L = 512, 1024
model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
tok = torch.randint(7, 11, (B, L), dtype=torch.uint8, device="cuda")
q = torch.rand(B, L, device="cuda")
for _ in range(2):
    model.forward_tokens(tok, q, False, True)
torch.cuda.synchronize()
n = 3 * 4096 * 2
buf = torch.empty(n, dtype=torch.int64)
check(lib().dcb200_ctx_read_workspace(model._ctx_now().handle, b"trace", C.c_void_p(buf.data_ptr()), n * 8))
a = buf.numpy().reshape(3, 4096, 2)
t0 = min(int(a[r, 0, 1]) for r in range(3) if a[r, 0, 1] > 0)
maxn = int(sys.argv[1]) if len(sys.argv) > 1 else 260
for r, name in enumerate(["MMA", "EPI", "PROD"]):
    print("==", name)
    prev = None
    out = []
    for i in range(maxn):
        tag, t = int(a[r, i, 0]), int(a[r, i, 1])
        if t == 0:
            break
        dt = t - t0
        out.append(f"{tag}@{dt}" + (f"(+{t - prev})" if prev is not None else ""))
        prev = t
    for i in range(0, len(out), 8):
        print("  ".join(out[i:i + 8]))
