"""Wall-clock of the two CLI steps on a synthetic FASTQ: `predict` (FASTQ -> predictions/0/*.pt) and `chop`
(predictions + FASTQ -> chopped .fq.gz), with the stages of each timed.    python tools/bench_cli.py [reads]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepchopper_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
extra = sys.argv[2:]
rng = np.random.default_rng(3)
recs = synth.fastq_reads(rng, n, lengths=synth.read_lengths(rng, n))
d = tempfile.mkdtemp()
fq = os.path.join(d, "reads.fq")
open(fq, "wb").write(synth.fastq_text(recs))
bases = sum(len(s) for _, s, _ in recs)
print(f"{n} reads, {bases / 1e6:.1f} M bases, FASTQ {os.path.getsize(fq) / 1e6:.1f} MB")
env = dict(os.environ, PYTHONPATH=ROOT)
t0 = time.time()
subprocess.check_call([sys.executable, "-m", "deepchopper_b200.cli", "predict", fq, "-o", os.path.join(d, "pred"), "--random-init",
                       "--bucket", "-v"] + extra, env=env, cwd=d)
t1 = time.time()
sz = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(d, "pred")) for f in fs)
print(f"predict: {t1 - t0:.2f} s wall ({bases / (t1 - t0) / 1e6:.2f} M bases/s incl. start-up), predictions {sz / 1e6:.0f} MB")
subprocess.check_call([sys.executable, "-m", "deepchopper_b200.cli", "chop", os.path.join(d, "pred", "0"), fq, "-t", "16",
                       "-o", os.path.join(d, "out")], env=env, cwd=d)
t2 = time.time()
print(f"chop: {t2 - t1:.2f} s wall ({bases / (t2 - t1) / 1e6:.2f} M bases/s incl. start-up)")
