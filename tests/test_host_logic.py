"""CPU tests of the host-side logic: batch planning / sharding (incl. a 2-process gloo run), the
prediction-directory wire format, FASTQ indexing."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

from deepchopper_b200 import encode, synth, writer
from deepchopper_b200.predict import plan_batches, shard_batches
from oracle import smooth_ref as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_batches_covers_every_read_once():
    rng = np.random.default_rng(0)
    lens = synth.read_lengths(rng, 5000)
    bs = plan_batches(lens, token_budget=256 * 1024)
    seen = np.concatenate([b.rows for b in bs])
    assert np.array_equal(np.sort(seen), np.arange(lens.size))
    for b in bs:
        assert b.Lpad == lens[b.rows].max() + 1 and b.Lrow % 128 == 0 and b.Lrow >= b.Lpad
        # the token budget is soft: row counts are rounded to full 128-row tiles (x4 when large), at most 1.3 x budget;
        # reads too long for 128 of them to fit get exactly one full row tile (the conv kernel's M dimension)
        assert b.rows.size <= 128 or b.rows.size * b.Lrow <= 1.3 * 256 * 1024
    strict = plan_batches(lens, token_budget=256 * 1024, long_read_cap=0)
    assert all(b.rows.size == 1 or b.rows.size * b.Lrow <= 1.3 * 256 * 1024 for b in strict)
    long_lens = rng.integers(16384, 32767, 300)
    lb = plan_batches(long_lens)
    assert np.array_equal(np.sort(np.concatenate([b.rows for b in lb])), np.arange(300))
    assert all(b.rows.size == 128 for b in lb[:-1]) and max(b.rows.size * b.Lrow for b in lb) <= 128 * 32768
    # reference batching: FASTQ order, fixed rows
    fb = plan_batches(lens, token_budget=1 << 62, max_rows=12, sort=False)
    assert all(np.array_equal(b.rows, np.arange(i * 12, min(lens.size, (i + 1) * 12))) for i, b in enumerate(fb))


def test_group_batches_keeps_each_batch_collation():
    """predict.group_batches: FASTQ-order batches packed into launches; every row keeps its own batch's collated length,
    every batch lands in exactly one launch, rows stay in batch order, the token budget holds (single-batch launches aside)."""
    from deepchopper_b200.predict import Batch, group_batches
    rng = np.random.default_rng(4)
    lens = np.clip(np.round(rng.lognormal(np.log(1000), 0.6, 3000)), 200, 32767).astype(np.int64)
    batches = []
    for i in range(0, lens.size, 12):
        rows = np.arange(i, min(i + 12, lens.size))
        lp = int(lens[rows].max()) + 1
        batches.append(Batch(rows, lp, (lp + 127) // 128 * 128))
    for budget in (1 << 20, 1 << 16, 1):
        launches = group_batches(batches, budget)
        seen = sorted(pos for g in launches for pos, _, _, _ in g.members)
        assert seen == list(range(len(batches)))
        for g in launches:
            assert g.Lrow % 128 == 0 and g.Lpad == int(g.lpad.max()) and g.Lpad <= g.Lrow
            assert g.rows.size * g.Lrow <= budget or len(g.members) == 1
            for pos, r0, r1, b in g.members:
                assert b is batches[pos] and np.array_equal(g.rows[r0:r1], b.rows)
                assert (g.lpad[r0:r1] == b.Lpad).all() and b.Lrow <= g.Lrow
        if budget == 1:
            assert len(launches) == len(batches)
    padded = sum(g.rows.size * g.Lrow for g in group_batches(batches, 1 << 20))
    assert padded <= 1.2 * sum(b.rows.size * b.Lrow for b in batches)      # similar lengths share a launch


def test_shard_batches_partition():
    rng = np.random.default_rng(1)
    bs = plan_batches(synth.read_lengths(rng, 20000))
    for world in (1, 2, 4, 8):
        parts = [shard_batches(bs, r, world) for r in range(world)]
        ids = sorted(id(b) for p in parts for b in p)
        assert ids == sorted(id(b) for b in bs)
        loads = [sum(b.rows.size * b.Lrow for b in p) for p in parts]
        assert max(loads) <= 1.15 * (sum(loads) / world) + 512 * 1024


def test_gloo_world2_sharding_agrees():
    """N>1 path on CPU: two gloo ranks plan the same batches, take disjoint shards, and the union of
    their read counts (all-reduced) is the whole read set -- no data-path collective needed."""
    script = r'''
import os, sys
sys.path.insert(0, os.environ["DCB_ROOT"])
import numpy as np, torch, torch.distributed as dist
from deepchopper_b200 import synth
from deepchopper_b200.predict import plan_batches, shard_batches
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lens = synth.read_lengths(np.random.default_rng(3), 4000)
mine = shard_batches(plan_batches(lens), rank, world)
t = torch.tensor([sum(b.rows.size for b in mine), int(sum(lens[b.rows].sum() for b in mine))], dtype=torch.int64)
dist.all_reduce(t)
assert int(t[0]) == lens.size and int(t[1]) == int(lens.sum()), t
rows = torch.zeros(lens.size, dtype=torch.int64)
for b in mine: rows[torch.from_numpy(b.rows)] += 1
dist.all_reduce(rows)
assert bool((rows == 1).all())
dist.destroy_process_group()
print("ok", rank)
'''
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "w.py")
        open(path, "w").write(script)
        env = dict(os.environ, DCB_ROOT=ROOT)
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29731", path],
                             env=env, capture_output=True, text=True, timeout=240)
        assert out.returncode == 0, out.stderr[-2000:]
        assert out.stdout.count("ok") == 2


def test_prediction_wire_format_roundtrip(tmp_path):
    """writer.batch_dict/write_batch produce what the reference loader consumes (src/smooth/predict.rs:263-317)."""
    rng = np.random.default_rng(2)
    recs = synth.fastq_reads(rng, 5, 30, 90)
    buf = np.frombuffer(synth.fastq_text(recs), dtype=np.uint8)
    ix = encode.index_fastq(buf)
    lens = ix.seq_len.astype(np.int64)
    Lpad = int(lens.max()) + 1
    Lrow = (Lpad + 127) // 128 * 128
    tok = torch.full((5, Lrow), 4, dtype=torch.uint8)
    qual = torch.zeros((5, Lrow))
    logits = torch.randn(5, Lrow, 2)
    table = {ord("A"): 7, ord("C"): 8, ord("G"): 9, ord("T"): 10, ord("N"): 11}
    for b in range(5):
        n = int(lens[b])
        tok[b, Lpad - 1 - n:Lpad - 1] = torch.tensor([table[c] for c in ix.seq(b)], dtype=torch.uint8)
        tok[b, Lpad - 1] = 1
    d = writer.batch_dict(logits, tok, qual, encode.id_rows(ix, np.arange(5), np.zeros(5, bool)), lens, Lpad)
    path = writer.write_batch(str(tmp_path), 0, 3, d)
    assert path.endswith(os.path.join("0", "0_3.pt"))
    back = torch.load(path)
    assert back["prediction"].dtype == torch.float32 and back["target"].dtype == torch.int64
    assert back["seq"].dtype == torch.int64 and back["id"].dtype == torch.int64 and back["id"].shape == (5, 256)
    preds = S.load_predicts_from_batch(back["prediction"].numpy(), back["target"].numpy(), back["seq"].numpy(),
                                       back["id"].numpy())
    assert sorted(preds) == sorted(r[0] for r in recs)
    for rid, s, _ in recs:
        assert preds[rid].seq == s and len(preds[rid].prediction) == len(s)


def test_index_fastq_validates():
    import pytest
    with pytest.raises(ValueError):
        encode.index_fastq(np.frombuffer(b"@a\nACGT\n+\nIII\n", dtype=np.uint8))
    with pytest.raises(ValueError):
        encode.index_fastq(np.frombuffer(b"a\nACGT\n+\nIIII\n", dtype=np.uint8))
    ix = encode.index_fastq(np.frombuffer(b"@a desc x\nACGT\n+\nIIII\n@b\nAC\n+\nII", dtype=np.uint8))
    assert [ix.name(0), ix.header(0), ix.name(1)] == ["a", "a desc x", "b"]
    assert ix.seq(1) == b"AC" and ix.qual(1) == b"II"
