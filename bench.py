#!/usr/bin/env python
"""bench.py -- DeepChopper predict+smooth hot path on B200 (BASELINE.json metric: bases/s, reads/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

A step is one pass of the hot path (GPU encode -> HyenaDNA-small-32k classifier -> smooth/interval/
chop coordinates) over the whole per-GPU workload: BASELINE config[1], 100k synthetic reads with
log-normal lengths (median 1 kb, clipped to [200, 8000]) in length-bucketed batches.  `value` is
measured with the inputs resident in HBM; `e2e` goes through the C-ABI call on pinned HOST buffers
(dcb200_predict_batch_host: H2D + compute + D2H inside the timed region).  Multi-GPU: reads are
independent, every rank owns its own 100k-read shard (weak scaling), no collective on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bases_per_sec_predict_smooth"
UNIT = "bases/s"

# algorithmic work per padded token (SURVEY §8d / DESIGN.md), used for the roofline figures
FLOPS_PER_TOKEN = {"block": 2 * 256 * 256 + 2 * 2 * 256 * 1024,  # out_proj + LN2 + fc1 + GELU + fc2 + residual + LN, one kernel
                   "mlp": 2 * 2 * 256 * 1024,              # (DCB200_BLOCK=split) fused fc1 + GELU + fc2 + residual + LN
                   "in_proj": 2 * 256 * 768,               # fused with the short conv + first gate (Toeplitz path)
                   "out_proj": 2 * 256 * 256,
                   "head1": 2 * 256 * 1024, "head2": 2 * (1024 * 1024 + 2 * 1024)}
BYTES_PER_TOKEN = {"hyena_conv": 768 * 2 + 256 * 2,        # FFT fallback: read z (3 channels) + write y, bf16
                   "toeplitz_conv": 3 * 256 * 2,             # read vv + gate, write y (the Toeplitz table stays in L2)
                   "embed_ln": 1 + 256 * 4 + 256 * 2,      # token in, fp32 residual + bf16 LN out
                   "encode": 2 + 1 + 4,                     # seq+qual chars in, token + fp32 quality out
                   "smooth_chop": 1}                        # int8 label in (coordinates out are O(reads))


# kernels launched once per Hyena layer (4x per batch): their per-token work counts once per layer
PER_LAYER = {"in_proj", "out_proj", "mlp", "block", "toeplitz_conv", "hyena_conv"}
N_LAYERS = 4
# one Hyena layer as built: dense FLOPs and HBM bytes per token
LAYER_FLOPS = 2 * 256 * 768 + 2 * 256 * 256 + 2 * 2 * 256 * 1024
LAYER_BYTES = (512 + 1024) + 1536 + (512 + 1024 + 1024 + 512)  # front end, long conv, block tail


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback"}


def synth_workload(n_reads: int, seed: int, workload: str = "configs1"):
    """Synthetic dRNA-like reads as one byte blob [all seq strings | all quality strings]."""
    from deepchopper_b200 import synth
    rng = np.random.default_rng(seed)
    if workload == "stress":   # BASELINE configs[3]: 16-32 kb reads (not the headline; `--workload stress`)
        lens = rng.integers(16384, 32767, n_reads).astype(np.int64)
    else:
        lens = synth.read_lengths(rng, n_reads, hi=8000)
    total = int(lens.sum())
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, total, dtype=np.uint8)]
    seq[rng.random(total, dtype=np.float32) < 0.001] = ord("N")
    q = np.clip(np.rint(rng.standard_normal(total, dtype=np.float32) * 8 + 20), 1, 50).astype(np.uint8) + 33
    blob = np.concatenate([seq, q])
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return blob, off, off + total, lens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world <= 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, world: int, device) -> float:
    if world <= 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---- CPU reference (oracle) legs --------------------------------------------------------------------

def cpu_reference_pass(n_reads: int, seed: int, state_dict=None, threads: int | None = None, model=None):
    """The reference's own predict+smooth path restated on the host (oracle/): FASTQ order, batch 16,
    left-pad to the batch max (tokenizer.py:34-93), fp32 model on all host cores, then the smoothing /
    interval pass.  Returns (bases, reads, seconds, model)."""
    from oracle import hyena_ref as H
    from oracle import cref
    from deepchopper_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    if model is None:
        model = H.make_reference_model(0)
        if state_dict is not None:
            model.load_state_dict(state_dict)
    rng = np.random.default_rng(seed)
    lens = synth.read_lengths(rng, n_reads, hi=8000)
    recs = synth.fastq_reads(rng, n_reads, lengths=lens)
    c = cref.load()
    t0 = time.perf_counter()
    bases = 0
    with torch.no_grad():
        for i in range(0, n_reads, 16):
            feats = [H.tokenize_read(rid, s, q) for rid, s, q in recs[i:i + 16]]
            batch = H.collate(feats)
            logits = model(batch["input_ids"], batch["input_quals"])
            lab = (logits[..., 1] > logits[..., 0]).to(torch.int8).numpy()
            L = lab.shape[1]
            ln = np.array([len(f["input_ids"]) - 1 for f in feats], dtype=np.int32)
            starts = (np.arange(len(feats)) * L + (L - 1) - ln).astype(np.int64)
            c.smooth_chop(lab.reshape(-1), starts, ln, threads=threads)
            bases += int(ln.sum())
    return bases, n_reads, time.perf_counter() - t0, model


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_reads
    model = None
    times = []
    bases = 0
    for step in range(args.warmup + args.steps):
        b, r, dt, model = cpu_reference_pass(sample, args.seed + step, threads=threads, model=model)
        if step >= args.warmup:
            times.append(dt)
            bases += b
    total = sum(times)
    value = bases / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic reads, log-normal length (median 1 kb, max 8 kb)",
                   "batching": "FASTQ order, batch 16, left-pad to batch max (reference collator)",
                   "note": "reference = CPU restatement (oracle/): the Rust+Lightning reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} reads per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reads_per_sec": sample * args.steps / total,
    }
    print(json.dumps(line), flush=True)


# ---- our arm -----------------------------------------------------------------------------------------

def run_ours(args, rank, local, world):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; deepchopper_b200 has no CPU fallback")
    from deepchopper_b200 import _native
    from deepchopper_b200.init_weights import random_state_dict
    from deepchopper_b200.model import DeepChopper
    from deepchopper_b200.predict import DevicePipeline, HostPipeline, plan_batches

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    sd = random_state_dict(0)
    model = DeepChopper.from_state_dict(sd, device=dev)
    blob, seq_off, qual_off, lens = synth_workload(args.reads, args.seed + 1000 * rank, args.workload)
    batches = plan_batches(lens, token_budget=args.token_budget)
    bases = int(lens.sum())
    padded_tokens = int(sum(b.rows.size * b.Lrow for b in batches))

    pipe = DevicePipeline(model)
    pipe.upload(blob, seq_off, qual_off, lens, batches)
    ctx = _native.torch_context(dev)

    # warm-up (also builds the per-FFT-size filter spectra)
    for _ in range(args.warmup):
        pipe.run_all()
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    ctx.profile(True)
    ctx.profile_read(reset=True)
    launches0 = ctx.launches
    barrier(world)
    torch.cuda.synchronize(dev)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        pipe.run_all()
    ev1.record()
    torch.cuda.synchronize(dev)
    barrier(world)
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)

    ms_total = max_over_ranks(ms_total, world, dev)
    bases_all = sum_over_ranks(bases, world, dev)
    reads_all = sum_over_ranks(args.reads, world, dev)
    tokens_all = sum_over_ranks(padded_tokens, world, dev)
    value = bases_all * args.steps / (ms_total / 1e3)

    # ---- e2e through the C ABI on pinned host buffers ------------------------------------------------
    hp = HostPipeline(model)
    hp.pack(blob, seq_off, qual_off, lens, batches)
    h2d, d2h = hp.bytes_per_pass()
    hp.run_all()  # warm
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier(world)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hp.run_all()
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world, dev)
    e2e_value = bases_all * e2e_steps / e2e_s
    h2d = int(sum_over_ranks(h2d, world, dev))  # whole-job bytes per step, like `value`
    d2h = int(sum_over_ranks(d2h, world, dev))

    # ---- per-kernel roofline (CUDA events on the launching stream, timed region above) ---------------
    kernels = {}
    tok_steps = padded_tokens * args.steps
    base_steps = bases * args.steps
    for name, (ms, cnt) in prof.items():
        ent = {"ms_total": ms, "launches": cnt, "share": ms / ms_total if ms_total else None}
        mult = N_LAYERS if name in PER_LAYER else 1
        if name in FLOPS_PER_TOKEN:
            ach = FLOPS_PER_TOKEN[name] * mult * tok_steps / (ms / 1e3) / 1e12
            ent.update(bound="tensor", achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"])
        elif name in BYTES_PER_TOKEN:
            units = base_steps if name in ("encode", "smooth_chop") else tok_steps
            ach = BYTES_PER_TOKEN[name] * mult * units / (ms / 1e3) / 1e9
            ent.update(bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"])
        kernels[name] = ent
    # one whole Hyena layer (in_proj+conv front end, long conv, out_proj, MLP) against both rooflines
    layer_ms = sum(kernels[k]["ms_total"] for k in PER_LAYER if k in kernels)
    layer = None
    if layer_ms > 0:
        tl = tok_steps * N_LAYERS
        tf = LAYER_FLOPS * tl / (layer_ms / 1e3) / 1e12
        gb = LAYER_BYTES * tl / (layer_ms / 1e3) / 1e9
        layer = {"ns_per_token_layer": layer_ms * 1e6 / tl, "dense_tflops": tf, "frac_tensor": tf / peaks["tflops"],
                 "hbm_gbs": gb, "frac_hbm": gb / peaks["hbm_gbs"],
                 "note": "dense GEMM FLOPs only (the long convolution's Toeplitz MMAs are extra work, not counted)"}
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    roofline = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01c_traffic.json")
    if dom and os.path.exists(tpath):
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture, scaled
        # from the profiled batch to this run's average launch (bytes per token x tokens per launch)
        ent = json.load(open(tpath))["kernels"].get(dom)
        if ent:
            launches_per_step = kernels[dom]["launches"] / max(1, args.steps)
            tokens_per_launch = padded_tokens * (N_LAYERS if dom in PER_LAYER else 1) / max(1.0, launches_per_step)
            traffic = ent["dram_bytes_per_token"] * tokens_per_launch
    if dom:
        k = kernels[dom]
        roofline = {"kernel": dom, "bound": k.get("bound"), "achieved": k.get("achieved"), "peak": k.get("peak"),
                    "unit": k.get("unit"), "frac": k.get("frac"), "traffic": traffic, "peak_source": peaks["source"],
                    "avg_launch_ms": k["ms_total"] / max(1, k["launches"])}

    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"configs[1]: {args.reads} synthetic reads per GPU, log-normal length (median 1 kb, "
                                "clipped to [200, 8000]), random-init HyenaDNA-small-32k + DeepChopper head"
                                if args.workload == "configs1" else
                                f"configs[3] long-read stress: {args.reads} synthetic reads per GPU, length U[16384, 32766], "
                                "random-init HyenaDNA-small-32k + DeepChopper head"),
                   "batching": f"length-bucketed, <= {args.token_budget} padded tokens per batch, left-pad to batch max",
                   "batches_per_step": len(batches), "l2": "inputs_larger_than_l2", "parallelism": f"read-sharded x{world}"},
        "reads_per_sec": reads_all * args.steps / (ms_total / 1e3),
        "padded_tokens_per_sec": tokens_all * args.steps / (ms_total / 1e3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "dcb200_predict_batch_host (pinned host buffers)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "hyena_layer": layer,
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        b, r, dt, _ = cpu_reference_pass(args.cpu_reads, args.seed, state_dict=sd)
        line["cpu_baseline"] = {"value": b / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{r} reads of the same length distribution, FASTQ order, batch 16, "
                                          f"fp32 oracle on all host cores ({dt:.1f} s)",
                                "reads_per_sec": r / dt}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--token-budget", type=int, default=1024 * 1024)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-reads", type=int, default=192)
    ap.add_argument("--ref-reads", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "stress"],
                    help="configs1 = BASELINE configs[1] (the headline); stress = configs[3], 16-32 kb reads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        # rank 0 alone times the CPU reference; other ranks exit 0 without work (no process group needed)
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, local, world = dist_setup(args.gpus)
    try:
        if False:
            pass
        else:
            run_ours(args, rank, local, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
