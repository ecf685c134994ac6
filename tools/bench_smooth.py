"""BASELINE config 5 (post-processing only): GPU smooth / interval / chop-coordinate pass over precomputed per-base
labels resident in HBM.  Prints reads/s, bases/s and achieved GB/s (1 B per base) and checks a sample against the
C oracle.    python tools/bench_smooth.py [--reads 2000000]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepchopper_b200 import _native  # noqa: E402
from deepchopper_b200._native import ChopParams  # noqa: E402
from deepchopper_b200.smooth import smooth_chop_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=2_000_000)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--check", type=int, default=20000)
args = ap.parse_args()

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(20261018)
R = args.reads
# log-normal lengths, background flip noise p = 0.02, 60 % of reads with a terminal adapter run of 30-120
# bases with interior noise p = 0.05, up to 2 more internal runs
lens = torch.clamp(torch.round(torch.exp(torch.randn(R, generator=g, device=dev) * 0.6 + np.log(1000.0))), 200, 8192).to(torch.int32)
starts = torch.zeros(R, dtype=torch.int64, device=dev)
starts[1:] = torch.cumsum(lens.to(torch.int64), 0)[:-1]
N = int(lens.sum().item())
rid = torch.repeat_interleave(torch.arange(R, device=dev, dtype=torch.int32), lens.to(torch.int64))
pos = torch.arange(N, device=dev, dtype=torch.int64) - starts[rid.long()]
ln = lens[rid.long()].to(torch.int64)
labels = (torch.rand(N, generator=g, device=dev) < 0.02)
term = torch.rand(R, generator=g, device=dev) < 0.6
tlen = torch.randint(30, 121, (R,), generator=g, device=dev)
in_term = term[rid.long()] & (pos >= ln - tlen[rid.long()])
for k in range(2):  # internal runs
    has = torch.rand(R, generator=g, device=dev) < 0.3
    rl = torch.randint(30, 121, (R,), generator=g, device=dev)
    st = (torch.rand(R, generator=g, device=dev) * (lens - 150).clamp(min=1)).to(torch.int64)
    in_term |= has[rid.long()] & (pos >= st[rid.long()]) & (pos < st[rid.long()] + rl[rid.long()])
keep1 = torch.rand(N, generator=g, device=dev) >= 0.05
labels = torch.where(in_term, keep1, labels).to(torch.int8)
del rid, pos, ln, in_term, keep1
torch.cuda.synchronize()

ctx = _native.torch_context(dev)
params = ChopParams.default()
out = smooth_chop_device(labels, starts, lens, params, None, ctx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    out = smooth_chop_device(labels, starts, lens, params, None, ctx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters

# parity on a sample against the C oracle (bit-exact)
from oracle import cref  # noqa: E402  (checker only)
c = cref.load()
k = min(args.check, R)
nb = int((starts[k - 1] + lens[k - 1]).item())
chk = c.smooth_chop(labels[:nb].cpu().numpy(), starts[:k].cpu().numpy(), lens[:k].cpu().numpy())
names = ["n_adapter", "adapter_iv", "n_keep", "keep_iv", "action"]
for nm, t in zip(names, out):
    a = t[:k].cpu().numpy()
    b = chk[nm]
    if nm.endswith("_iv"):
        cnt = chk["n_adapter" if nm == "adapter_iv" else "n_keep"]
        for r in range(k):
            assert np.array_equal(a[r, :cnt[r]], b[r, :cnt[r]]), (nm, r)
    else:
        assert np.array_equal(a, b), nm
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
print(json.dumps({"workload": f"configs[4] scaled: {R} reads of precomputed int8 labels resident in HBM ({N} bases)",
                  "ms_per_pass": ms, "reads_per_sec": R / (ms / 1e3), "bases_per_sec": N / (ms / 1e3),
                  "achieved_gbs": N / (ms / 1e3) / 1e9, "frac_of_hbm_peak": N / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                  "parity_checked_reads": k, "adapters_found": int(out[0].sum().item())}))
