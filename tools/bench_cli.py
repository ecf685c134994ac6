"""Wall-clock of the file-to-file routes on a synthetic FASTQ (subprocesses, start-up included):
  A  `predict --chop`                      one pass, no prediction files (deepchopper_b200/fused.py)
  B  `predict --compact` then `chop`       bit-packed label sidecars
  C  `predict` then `chop`                 the reference's .pt dicts (28 bytes per token)
  D  `predict --chop -b 16`                one pass in the REFERENCE's batching (FASTQ order, batch 16, batches packed into
                                           launches with every row padded as in its own batch); not compared with A-C:
                                           left pads are semantic, so its logits are those of other batches
and whether the outputs of A-C are byte-identical.    python tools/bench_cli.py [reads] [routes, e.g. AB]"""
import glob
import gzip
import hashlib
import os
import resource
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepchopper_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
routes = sys.argv[2] if len(sys.argv) > 2 else "ABC"
rng = np.random.default_rng(3)
lens = synth.read_lengths(rng, n)
d = tempfile.mkdtemp()
fq = os.path.join(d, "reads.fq")
with open(fq, "wb") as f:
    for i in range(0, n, 5000):
        f.write(synth.fastq_text(synth.fastq_reads(rng, min(5000, n - i), lengths=lens[i:i + 5000])).replace(
            b"@read_", b"@read_%d_" % (i // 5000)))
bases = int(lens.sum())
print(f"{n} reads, {bases / 1e6:.1f} M bases, FASTQ {os.path.getsize(fq) / 1e6:.1f} MB")
env = dict(os.environ, PYTHONPATH=ROOT)
py = [sys.executable, "-m", "deepchopper_b200.cli"]
common = ["--random-init", "--bucket", "--token-budget", str(1024 * 1024), "-v"]
digest = {}


def run(cmd):
    t0 = time.time()
    subprocess.check_call(cmd, env=env, cwd=d)
    return time.time() - t0


def out_digest(prefix):
    f = glob.glob(os.path.join(d, prefix + ".*.chop.fq.gz"))[0]
    return os.path.basename(f).split(".", 1)[1], hashlib.md5(gzip.open(f, "rb").read()).hexdigest()


if "A" in routes:
    t = run(py + ["predict", fq, "--chop", "--chop-output", os.path.join(d, "A"), "-t", "16"] + common)
    digest["A"] = out_digest("A")
    print(f"A predict --chop: {t:.2f} s wall = {bases / t / 1e6:.2f} M bases/s (start-up included)")
if "B" in routes:
    t1 = run(py + ["predict", fq, "-o", os.path.join(d, "predB"), "--compact"] + common)
    sz = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(d, "predB")) for f in fs)
    t2 = run(py + ["chop", os.path.join(d, "predB", "0"), fq, "-t", "16", "-o", os.path.join(d, "B"), "-v"])
    digest["B"] = out_digest("B")
    print(f"B predict --compact {t1:.2f} s ({sz / 1e6:.0f} MB of sidecars) + chop {t2:.2f} s = {bases / (t1 + t2) / 1e6:.2f} M bases/s")
if "C" in routes:
    t1 = run(py + ["predict", fq, "-o", os.path.join(d, "predC")] + common)
    sz = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(d, "predC")) for f in fs)
    t2 = run(py + ["chop", os.path.join(d, "predC", "0"), fq, "-t", "16", "-o", os.path.join(d, "C"), "-v"])
    digest["C"] = out_digest("C")
    print(f"C predict {t1:.2f} s ({sz / 1e6:.0f} MB of .pt) + chop {t2:.2f} s = {bases / (t1 + t2) / 1e6:.2f} M bases/s")
if "D" in routes:
    t = run(py + ["predict", fq, "--chop", "--chop-output", os.path.join(d, "D"), "-t", "16", "--random-init", "-b", "16",
                  "--token-budget", str(1024 * 1024), "-v"])
    print(f"D predict --chop -b 16 (reference batching): {t:.2f} s wall = {bases / t / 1e6:.2f} M bases/s (start-up included)")
print("outputs:", digest, "identical:", len(set(digest.values())) == 1)
print(f"peak RSS of the children: {resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1024:.0f} MB")
