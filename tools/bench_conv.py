"""Long convolution: tensor-core Toeplitz kernel vs blocked shared-memory FFT kernel, per-launch CUDA-event time inside
a full forward at several read lengths (B rows of L tokens).  Prints ns per token-layer for both and the crossover.
    python tools/bench_conv.py [iters]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepchopper_b200 import _native  # noqa: E402
from deepchopper_b200.init_weights import random_state_dict  # noqa: E402
from deepchopper_b200.model import DeepChopper  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shapes = [(512, 2048), (256, 4096), (256, 5120), (128, 6144), (128, 7168), (128, 8192), (128, 10240), (128, 12288),
          (128, 16384), (128, 16512), (128, 20480), (128, 24576), (128, 32768)]
if len(sys.argv) > 2:
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[2:]]
model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
ctx = _native.torch_context(torch.device("cuda", 0))
rows = []
for B, L in shapes:
    tok = torch.randint(7, 11, (B, L), dtype=torch.uint8, device="cuda")
    q = torch.rand(B, L, device="cuda")
    res = {"B": B, "L": L}
    for kind, opt in (("toeplitz", 1 << 30), ("fft", 0)):
        ctx.set_option("fft_min_len", opt)
        model.forward_tokens(tok, q, False, True)
        torch.cuda.synchronize()
        ctx.profile(True)
        ctx.profile_read(reset=True)
        for _ in range(iters):
            model.forward_tokens(tok, q, False, True)
        torch.cuda.synchronize()
        prof = ctx.profile_read(reset=True)
        ctx.profile(False)
        ms, cnt = prof["toeplitz_conv" if kind == "toeplitz" else "fft_conv"]
        res[kind + "_us"] = ms / cnt * 1e3
        res[kind + "_ns_per_token_layer"] = ms / cnt * 1e6 / (B * L)
        res["layer_other_ns"] = sum(prof[k][0] / prof[k][1] for k in ("in_proj", "block")) * 1e6 / (B * L)
    rows.append(res)
    print(json.dumps(res), flush=True)
    del tok, q
    torch.cuda.empty_cache()
