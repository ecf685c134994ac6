import gzip, os, subprocess, sys, tempfile, glob
import numpy as np
sys.path.insert(0, '.')
from deepchopper_b200 import synth
rng = np.random.default_rng(3)
n = 6000
recs = synth.fastq_reads(rng, n, lengths=synth.read_lengths(rng, n))
d = tempfile.mkdtemp(); fq = os.path.join(d, "reads.fq"); open(fq, "wb").write(synth.fastq_text(recs))
env = dict(os.environ, PYTHONPATH=os.getcwd())
outs = []
for name, extra in (("pt", []), ("compact", ["--compact"])):
    subprocess.check_call([sys.executable, "-m", "deepchopper_b200.cli", "predict", fq, "-o", os.path.join(d, name), "--random-init", "--bucket"] + extra, env=env, cwd=d)
    subprocess.check_call([sys.executable, "-m", "deepchopper_b200.cli", "chop", os.path.join(d, name, "0"), fq, "-t", "8", "-o", os.path.join(d, "out_" + name)], env=env, cwd=d)
    f = glob.glob(os.path.join(d, "out_" + name + ".*.chop.fq.gz"))[0]
    outs.append((os.path.basename(f), gzip.open(f, "rb").read()))
print(outs[0][0], outs[1][0], "identical:", outs[0][1] == outs[1][1], len(outs[0][1]))
