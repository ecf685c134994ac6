"""bench.py's JSON line (driver contract), checked on the arm that runs without a GPU: `--impl reference` times the CPU
restatement on a tiny sample and must print exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-reads", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bases_per_sec_predict_smooth" and d["unit"] == "bases/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--ref-reads", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]


def test_bench_global_job_is_rank_independent():
    """bench.py --gpus N: ONE global job.  Lengths and batch plan are functions of the seed only; a batch's bytes are a
    function of (seed, global batch index), so the read set does not depend on how many ranks share it; the shards of 1,
    2 and 8 ranks partition the same batches and the token imbalance of the greedy deal stays small."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from deepchopper_b200.predict import plan_batches, shard_batches
    lens = bench.synth_lengths(20000, 5, "configs1")
    assert np.array_equal(lens, bench.synth_lengths(20000, 5, "configs1"))
    batches = plan_batches(lens, token_budget=256 * 1024)
    index_of = {id(b): i for i, b in enumerate(batches)}
    all_tokens = sum(b.rows.size * b.Lrow for b in batches)
    for world in (1, 2, 8):
        shards = [shard_batches(batches, r, world) for r in range(world)]
        assert sorted(index_of[id(b)] for s in shards for b in s) == list(range(len(batches)))
        tok = [sum(b.rows.size * b.Lrow for b in s) for s in shards]
        assert sum(tok) == all_tokens and max(tok) / (sum(tok) / world) < 1.1
    b = batches[3]
    items_a = bench.make_items(lens, [b], 5, index_of)
    items_b = bench.make_items(lens, [batches[0], b], 5, index_of)
    assert np.array_equal(items_a[0][1], items_b[1][1]) and np.array_equal(items_a[0][4], lens[b.rows].astype(np.int32))
    buf, so, qo, ln = items_a[0][1:]
    assert buf.size == 2 * int(ln.sum()) and set(np.unique(buf[: int(ln.sum())])) <= set(b"ACGTN")
    assert buf[int(ln.sum()):].min() >= 34 and buf[int(ln.sum()):].max() <= 83 and qo[0] == int(ln.sum())


def test_takes_fft_mirror_of_the_cost_model():
    """bench.takes_fft mirrors model.cu:use_fft_conv (the GPU test checks that the kernels actually launched agree with
    it): full 128-row tiles take the FFT from ~6.7 k tokens, not while a second block is mostly empty, and always above
    ~9.9 k; a 16-row batch from ~3.4 k tokens; an explicit fft_min_len is a plain threshold; nothing beyond 32768."""
    import bench
    d = 6144
    assert not bench.takes_fft(6656, d) and bench.takes_fft(6784, d) and bench.takes_fft(8192, d)
    assert not bench.takes_fft(8320, d) and not bench.takes_fft(9856, d) and bench.takes_fft(9984, d) and bench.takes_fft(32768, d)
    assert not bench.takes_fft(3200, d, 16) and bench.takes_fft(3456, d, 16)
    assert not bench.takes_fft(6144, d, 64) and bench.takes_fft(6784, d, 200)
    assert bench.takes_fft(128, 0) and not bench.takes_fft(32768, 1 << 30) and not bench.takes_fft(32896, 0)
