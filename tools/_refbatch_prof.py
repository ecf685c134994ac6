import sys, json, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from deepchopper_b200 import synth, _native
from deepchopper_b200.predict import Batch
from deepchopper_b200.model import DeepChopper
from deepchopper_b200.init_weights import random_state_dict
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
model = DeepChopper.from_state_dict(random_state_dict(0), device=dev)
rng = np.random.default_rng(0)
n_reads = 1536
lens = synth.read_lengths(rng, n_reads, hi=8000)
batches = []
for i in range(0, n_reads, 16):
    rows = np.arange(i, min(i + 16, n_reads))
    lpad = int(lens[rows].max()) + 1
    batches.append(Batch(rows, lpad, (lpad + 127) // 128 * 128))
index_of = {id(b): i for i, b in enumerate(batches)}
items = bench.make_items(lens, batches, 5, index_of)
ms, prof, launches, pipe = bench.run_pass(model, items, 3, 1, dev, profile=True)
print("profiled ms/batch", ms / 3 / len(batches), "launches", launches)
print({k: (round(v[0],2), v[1]) for k, v in prof.items()})
ms2, _, _, pipe2 = bench.run_pass(model, items, 3, 1, dev, profile=False)
print("unprofiled ms/batch", ms2 / 3 / len(batches))
# host-side cost: time the python loop alone without sync
t0 = time.perf_counter(); pipe2.run_all(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("python issue time per batch ms", (t1 - t0) / len(batches) * 1e3, "total", (t2 - t0) / len(batches) * 1e3)
print("Lrow hist", np.percentile([b.Lrow for b in batches], [10, 50, 90, 100]))
