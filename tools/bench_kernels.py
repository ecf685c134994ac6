"""Per-kernel CUDA-event times of the forward pass at one fixed shape (default B=512, L=1024), averaged over iterations.
    python tools/bench_kernels.py [B L iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepchopper_b200 import _native  # noqa: E402
from deepchopper_b200.init_weights import random_state_dict  # noqa: E402
from deepchopper_b200.model import DeepChopper  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
tok = torch.randint(7, 11, (B, L), dtype=torch.uint8, device="cuda")
q = torch.rand(B, L, device="cuda")
ctx = _native.torch_context(torch.device("cuda", 0))
for _ in range(3):
    model.forward_tokens(tok, q, False, True)
torch.cuda.synchronize()
ctx.profile(True)
ctx.profile_read(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    model.forward_tokens(tok, q, False, True)
e1.record()
torch.cuda.synchronize()
prof = ctx.profile_read(reset=True)
tot = e0.elapsed_time(e1) / iters
print(f"B={B} L={L} tokens={B * L}: {tot:.3f} ms per forward, {B * L / tot / 1e3:.1f} M tokens/s")
for name, (ms, cnt) in prof.items():
    if cnt:
        print(f"  {name:14s} {ms / cnt * 1e3:9.1f} us per launch  x{cnt // iters:2d}  {ms / iters:8.3f} ms per forward")
