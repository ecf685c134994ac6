"""Timeline trace of the Toeplitz long-convolution kernel (DCB200_TRACE=toeplitz): stage cadence of the MMA warp.
    python tools/trace_toeplitz.py B L"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["DCB200_TRACE"] = "toeplitz"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepchopper_b200._native import check, lib  # noqa: E402
from deepchopper_b200.init_weights import random_state_dict  # noqa: E402
from deepchopper_b200.model import DeepChopper  # noqa: E402

B, L = int(sys.argv[1]), int(sys.argv[2])
model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
tok = torch.randint(7, 11, (B, L), dtype=torch.uint8, device="cuda")
q = torch.rand(B, L, device="cuda")
model.forward_tokens(tok, q, False, True)
torch.cuda.synchronize()
n = 3 * 4096 * 2
buf = torch.empty(n, dtype=torch.int64)
check(lib().dcb200_ctx_read_workspace(model._ctx_now().handle, b"trace", C.c_void_p(buf.data_ptr()), n * 8))
a = buf.numpy().reshape(3, 4096, 2)
for r, name in enumerate(["MMA", "EPI", "PROD"]):
    ev = [(int(a[r, i, 0]), int(a[r, i, 1])) for i in range(4096) if a[r, i, 1] > 0]
    if not ev:
        continue
    t0 = ev[0][1]
    print("==", name, len(ev), "events over", ev[-1][1] - t0, "cycles")
    if name == "MMA":
        st = [c for t, c in ev if t == 120]
        d = np.diff(st)
        print("  stages", len(st), "cadence mean %.0f median %.0f p90 %.0f max %d" % (d.mean(), np.median(d), np.percentile(d, 90), d.max()))
        for k in range(0, len(d), 200):
            print("   stages %4d-%4d: mean cadence %.0f" % (k, min(len(d), k + 200), d[k:k + 200].mean()))
    if len(sys.argv) > 3:
        print(" ".join(f"{t}:{c - t0}" for t, c in ev[:int(sys.argv[3])]))
