"""Host-side chop output throughput (dcb200_chop_write_bgzf) on synthetic reads: text MB/s and records/s for several
thread counts, next to the single-threaded Python/zlib writer this code replaced.  No GPU needed.
    python tools/bench_chop_writer.py [reads] [threads ...]"""
import os
import struct
import sys
import tempfile
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepchopper_b200 import synth  # noqa: E402
from deepchopper_b200.chop import write_chopped_fastq  # noqa: E402
from deepchopper_b200.encode import index_fastq  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
thread_list = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 0]
rng = np.random.default_rng(20261018)
lens = synth.read_lengths(rng, n_reads)
recs = synth.fastq_reads(rng, n_reads, lengths=lens)
buf = np.frombuffer(synth.fastq_text(recs), dtype=np.uint8)
ix = index_fastq(buf)
R = len(ix)
# 60 % of the reads lose a terminal adapter (one kept piece), 5 % an internal one (two pieces), the rest pass through
u = rng.random(R)
act = np.where(u < 0.60, 1, np.where(u < 0.65, 2, 0)).astype(np.uint8)
n_keep = np.where(act == 1, 1, np.where(act == 2, 2, 0)).astype(np.int32)
keep = np.zeros((R, 21, 2), np.int32)
L = ix.seq_len.astype(np.int64)
cut = (L * 0.9).astype(np.int32)
keep[:, 0, 1] = np.where(act == 2, L // 2 - 40, cut)
keep[:, 1, 0] = L // 2 + 40
keep[:, 1, 1] = cut
ad = np.zeros((R, 20, 2), np.int32)
n_ad = np.zeros(R, np.int32)
has = np.ones(R, np.uint8)
ptr = (buf.ctypes.data + ix.seq_off).astype(np.uint64)     # predicted sequence == FASTQ sequence here
plen = ix.seq_len.astype(np.int32)
tmp = tempfile.mkdtemp()
print(f"{R} reads, {buf.size / 1e6:.1f} MB of FASTQ text, host cores {os.cpu_count()}")
for level, t in [(lv, t) for lv in (6, 0) for t in thread_list]:
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        nrec, ntext = write_chopped_fastq(os.path.join(tmp, "o.fq.gz"), ix, has, ptr, plen, act, n_ad, ad, n_keep, keep,
                                          threads=t, level=level)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    sz = os.path.getsize(os.path.join(tmp, "o.fq.gz"))
    print(f"native level={level} threads={t or os.cpu_count():3d}: {best * 1e3:8.1f} ms  {ntext / best / 1e6:8.1f} MB/s text  "
          f"{nrec / best / 1e3:8.1f} k records/s  ({nrec} records, {sz / 1e6:.1f} MB bgzf)")


def python_writer(path, sample):
    """the sequential writer this code replaced: Python string assembly + zlib per 0xff00-byte block"""
    f = open(path, "wb")
    pend = bytearray()

    def block(data):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(data) + c.flush()
        f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25))
        f.write(comp)
        f.write(struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))
    n = 0
    for r in range(sample):
        rid, qual, seq = ix.name(r), ix.qual(r).decode("latin1"), ix.seq(r).decode("latin1")
        if act[r] == 0:
            pend += f"@{ix.header(r)}\n{seq}\n+\n{qual}\n".encode("latin1")
        else:
            tag = "T" if act[r] == 1 else "I"
            for s, e in keep[r, :n_keep[r]]:
                pend += f"@{rid}|{s}:{e}|{tag}\n{seq[s:e]}\n+\n{qual[s:e]}\n".encode("latin1")
        n += 1
        while len(pend) >= 0xff00:
            block(bytes(pend[:0xff00]))
            del pend[:0xff00]
    if pend:
        block(bytes(pend))
    f.close()
    return n


sample = min(R, 20_000)
t0 = time.perf_counter()
python_writer(os.path.join(tmp, "p.fq.gz"), sample)
dt = time.perf_counter() - t0
text = float((ix.seq_len[:sample].astype(np.int64) * 2).sum())
print(f"python/zlib 1 thread ({sample} reads): {dt * 1e3:8.1f} ms  ~{text * 0.93 / dt / 1e6:6.1f} MB/s text  {sample / dt / 1e3:6.1f} k reads/s")

# ---- reading the BGZF file back (dcb200_read_file_inflate) next to Python's gzip module -------------------
import gzip  # noqa: E402

from deepchopper_b200.encode import read_fastq_bytes  # noqa: E402

path = os.path.join(tmp, "o.fq.gz")
for t in (1, 2, 0):
    t0 = time.perf_counter()
    got = read_fastq_bytes(path, threads=t)
    dt = time.perf_counter() - t0
    print(f"native inflate threads={t or os.cpu_count():3d}: {dt * 1e3:8.1f} ms  {got.size / dt / 1e6:8.1f} MB/s text")
t0 = time.perf_counter()
ref = gzip.open(path, "rb").read()
dt = time.perf_counter() - t0
print(f"python gzip 1 thread       : {dt * 1e3:8.1f} ms  {len(ref) / dt / 1e6:8.1f} MB/s text   identical: {ref == got.tobytes()}")
