#!/usr/bin/env python
"""bench.py -- DeepChopper predict+smooth hot path on B200 (BASELINE.json metric: bases/s, reads/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

A step is one pass of the hot path (GPU encode -> HyenaDNA-small-32k classifier -> smooth/interval/chop coordinates)
over this rank's part of ONE global synthetic read set:
  N = 1   BASELINE configs[1]: 100k reads, log-normal lengths (median 1 kb, clipped to [200, 8000]);
  N > 1   BASELINE configs[2]: 2M reads of the same distribution (a fixed job: strong scaling).  Every rank derives the
          same read lengths from the seed, plans the length-bucketed batches once (`plan_batches`) and takes its share with
          `shard_batches(rank, world)` -- the product's own sharding, no collective on the data path; the line reports the
          per-rank token imbalance and the slowest / fastest rank.
`value` is measured with the inputs resident in HBM; `e2e` goes through the C-ABI call on pinned HOST buffers
(dcb200_predict_batch_host: H2D + compute + D2H inside the timed region; FASTQ parsing / indexing and result files are
outside it -- tools/bench_cli.py times the file-to-file path).  At N = 1 the line also carries `extra_configs`
(configs[3] long-read stress sample, configs[4] smooth-only over 10M reads, the product in the reference's own batching:
FASTQ order, batch 16), `gpu_eager_baseline` (the fp32/TF32
PyTorch restatement of the reference on the same GPU) and `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bases_per_sec_predict_smooth"
UNIT = "bases/s"

# algorithmic work per padded token (SURVEY §8d / DESIGN.md), used for the roofline figures
FLOPS_PER_TOKEN = {"block": 2 * 256 * 256 + 2 * 2 * 256 * 1024,  # out_proj + LN2 + fc1 + GELU + fc2 + residual + LN, one kernel
                   "in_proj": 2 * 256 * 768,               # fused with the short conv + first gate
                   "head1": 2 * 256 * 1024, "head2": 2 * (1024 * 1024 + 2 * 1024)}
BYTES_PER_TOKEN = {"fft_conv": 3 * 256 * 2,                 # read vv + gate, write y (filter spectra + block scratch stay in L2)
                   "toeplitz_conv": 3 * 256 * 2,             # read vv + gate, write y (the Toeplitz table stays in L2)
                   "embed_ln": 1 + 256 * 4 + 256 * 2,      # token in, fp32 residual + bf16 LN out
                   "encode": 2 + 1 + 4,                     # seq+qual chars in, token + fp32 quality out
                   "smooth_chop": 1}                        # int8 label in (coordinates out are O(reads))
# kernels launched once per Hyena layer (4x per batch): their per-token work counts once per layer
PER_LAYER = {"in_proj", "block", "toeplitz_conv", "fft_conv"}
N_LAYERS = 4
# one Hyena layer as built: dense FLOPs and HBM bytes per token
LAYER_FLOPS = 2 * 256 * 768 + 2 * 256 * 256 + 2 * 2 * 256 * 1024
LAYER_BYTES = (512 + 1024) + 1536 + (512 + 1024 + 1024 + 512)  # front end, long conv, block tail


def toeplitz_flops_per_token(L: float) -> float:
    """MMA work of the Toeplitz long convolution per token and layer: 256 channels x 2 x 128 x (L/128 + 1) / 2."""
    return 256 * 2.0 * 128 * (L / 128 + 1) / 2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback"}


# ---- one global synthetic read set ---------------------------------------------------------------------

def synth_lengths(n_reads: int, seed: int, workload: str) -> np.ndarray:
    """Read lengths of the whole job: a function of (seed, workload) only, so every rank derives the same set."""
    from deepchopper_b200 import synth
    rng = np.random.default_rng(seed)
    if workload == "stress":   # BASELINE configs[3]: 16-32 kb reads
        return rng.integers(16384, 32767, n_reads).astype(np.int64)
    return synth.read_lengths(rng, n_reads, hi=8000)


_QUAL_TABLE = None


def synth_batch(lens: np.ndarray, seed: int, batch_index: int):
    """The bytes of one batch, a function of (seed, global batch index): iid ACGT with 0.1 % N and Phred ~ clipped
    N(20, 8) as ASCII+33, packed [all sequences | all quality strings].  Returns (buf u8, seq_off, qual_off, len i32)."""
    global _QUAL_TABLE
    if _QUAL_TABLE is None:
        g = np.random.default_rng(7)
        _QUAL_TABLE = (np.clip(np.rint(g.standard_normal(1 << 16) * 8 + 20), 1, 50) + 33).astype(np.uint8)
    rng = np.random.default_rng([seed, batch_index])
    tot = int(lens.sum())
    buf = np.empty(2 * tot, dtype=np.uint8)
    buf[:tot] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, tot, dtype=np.uint8)]
    n_n = rng.binomial(tot, 0.001)
    buf[rng.integers(0, tot, n_n)] = ord("N")
    buf[tot:] = _QUAL_TABLE[rng.integers(0, 1 << 16, tot, dtype=np.uint16)]
    so = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    return buf, so, so + tot, lens.astype(np.int32)


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


def gather_over_ranks(x: float, world: int, device) -> list:
    if world <= 1:
        return [float(x)]
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


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


def gpu_eager_baseline(sd, seed: int, dev, n_reads: int = 1536):
    """the reference's own GPU route restated -- stock PyTorch eager ops (cuBLAS with TF32 matmuls as
    `import deepchopper` sets them; cuFFT in fp32) on the same B200, (i) in the reference's batching (FASTQ
    order, batch 16) and (ii) in this repo's length-bucketed batches.  Context only: it is the oracle module moved to
    the GPU, not the product and not the arm the driver divides by."""
    from oracle import hyena_ref as H
    from deepchopper_b200 import synth
    from deepchopper_b200.predict import plan_batches
    old = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    model = H.make_reference_model(0)
    model.load_state_dict(sd)
    model = model.to(dev)
    rng = np.random.default_rng(seed)
    lens = synth.read_lengths(rng, n_reads, hi=8000)
    out = {"reads": int(n_reads), "precision": "TF32 matmul (torch 'high', SURVEY T13), fp32 cuFFT, fp32 activations"}

    def batch_tensors(rows, L):
        ids = torch.full((len(rows), L), 4, dtype=torch.int64)
        q = torch.zeros((len(rows), L), dtype=torch.float32)
        for k, r in enumerate(rows):
            n = int(lens[r])
            ids[k, L - 1 - n:L - 1] = torch.from_numpy(rng.integers(7, 11, n))
            ids[k, L - 1] = 1
            q[k, L - 1 - n:L - 1] = 0.03
        return ids.to(dev), q.to(dev)

    def timed(batches):
        ts = [batch_tensors(rows, L) for rows, L in batches]
        with torch.no_grad():
            for ids, q in ts[:2]:
                model(ids, q)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for ids, q in ts:
                lg = model(ids, q)
                (lg[..., 1] > lg[..., 0]).to(torch.int8)
            e1.record()
            torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / 1e3

    try:
        ref_batches = [(list(range(i, min(i + 16, n_reads))), int(lens[i:i + 16].max()) + 1) for i in range(0, n_reads, 16)]
        dt = timed(ref_batches)
        out["reference_batching"] = {"value": float(lens.sum() / dt), "unit": UNIT, "batching": "FASTQ order, batch 16",
                                     "seconds": dt}
        # fp32 activations: z [T,768] + FFT buffers ~ 20 KB per token -> 256k-token batches
        bk = plan_batches(lens, token_budget=256 * 1024)
        dt = timed([(list(b.rows), b.Lpad) for b in bk])
        out["bucketed"] = {"value": float(lens.sum() / dt), "unit": UNIT, "seconds": dt,
                           "batching": "length-bucketed, <= 262144 padded tokens per batch (fp32 activations)"}
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:200]
    finally:
        torch.set_float32_matmul_precision(old)
        del model
        torch.cuda.empty_cache()
    return out


# ---- our arm -----------------------------------------------------------------------------------------

def takes_fft(L: int, fft_min_len: int, rows: int = 128) -> bool:
    """Mirror of model.cu:use_fft_conv (which long-convolution kernel a batch of `rows` rows of padded length L takes)."""
    if L > 32768:
        return False
    if fft_min_len != 6144:
        return L >= fft_min_len
    nb = (L + 8191) // 8192
    rem = rows % 128
    tiles = rows // 128 + (1.0 if rem > 64 else (0.5 if rem > 0 else 0.0))
    return rows * nb * 8192.0 * (1.57, 1.76, 1.95, 2.06)[nb - 1] < tiles * 128.0 * 0.29e-3 * L * L


def conv_work(batches, fft_min_len):
    """Which long-convolution kernel each batch takes (ctx option fft_min_len) -> per-step padded tokens of each kernel,
    the Toeplitz kernel's MMA FLOPs and the FFT kernel's fp32 FLOPs (5 N log2 N per complex transform of N = 8192
    points, two transforms per 8192-token block, per channel and layer)."""
    w = {"fft_tokens": 0, "toeplitz_tokens": 0, "toeplitz_flops": 0.0, "fft_flops": 0.0}
    for b in batches:
        t = b.rows.size * b.Lrow
        if takes_fft(b.Lrow, fft_min_len, int(b.rows.size)):
            w["fft_tokens"] += t
            blocks = (b.Lrow + 8191) // 8192
            w["fft_flops"] += b.rows.size * 256 * blocks * 2 * 5.0 * 8192 * 13
        else:
            w["toeplitz_tokens"] += t
            w["toeplitz_flops"] += t * toeplitz_flops_per_token(b.Lrow)
    return w


def kernel_table(prof, ms_total, tok_steps, base_steps, peaks, conv=None, steps=1):
    kernels = {}
    for name, (ms, cnt) in prof.items():
        ent = {"ms_total": ms, "launches": cnt, "share": ms / ms_total if ms_total else None}
        mult = N_LAYERS if name in PER_LAYER else 1
        if name in ("fft_conv", "toeplitz_conv") and conv is not None:
            # a batch takes ONE of the two kernels: each is charged the tokens it actually processed
            tk = conv["fft_tokens" if name == "fft_conv" else "toeplitz_tokens"] * steps * N_LAYERS
            ach = BYTES_PER_TOKEN[name] * tk / (ms / 1e3) / 1e9
            ent.update(bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"],
                       ns_per_token_layer=ms * 1e6 / max(1, tk))
            if name == "toeplitz_conv":
                # the bound of THIS run's launches: above ~2k tokens the kernel's own MMA work (64 KFLOP x (L/128+1)/2 per
                # token) outweighs its 1536 B of HBM traffic
                tf = conv["toeplitz_flops"] * steps * N_LAYERS / (ms / 1e3) / 1e12
                if tf / peaks["tflops"] > ent["frac"]:
                    ent.update(bound="tensor", achieved=tf, peak=peaks["tflops"], unit="TFLOP/s", frac=tf / peaks["tflops"],
                               note="bound chosen per run: Toeplitz MMA work of this run's batch lengths")
            else:
                ent["fp32_tflops"] = conv["fft_flops"] * steps * N_LAYERS / (ms / 1e3) / 1e12
                ent["note"] = ("fp32 FFT on CUDA cores in shared memory: instruction-issue bound, neither HBM nor tensor "
                               "(ncu: profiles/r02_summary.md); fp32_tflops = 5 N log2 N per transform")
        elif name in FLOPS_PER_TOKEN:
            ach = FLOPS_PER_TOKEN[name] * mult * tok_steps / (ms / 1e3) / 1e12
            ent.update(bound="tensor", achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"])
        elif name in BYTES_PER_TOKEN:
            units = base_steps if name in ("encode", "smooth_chop") else tok_steps
            ach = BYTES_PER_TOKEN[name] * mult * units / (ms / 1e3) / 1e9
            ent.update(bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"])
        kernels[name] = ent
    return kernels


def layer_summary(kernels, tok_steps, peaks):
    layer_ms = sum(kernels[k]["ms_total"] for k in PER_LAYER if k in kernels)
    if layer_ms <= 0:
        return None
    tl = tok_steps * N_LAYERS
    tf = LAYER_FLOPS * tl / (layer_ms / 1e3) / 1e12
    gb = LAYER_BYTES * tl / (layer_ms / 1e3) / 1e9
    return {"ns_per_token_layer": layer_ms * 1e6 / tl, "dense_tflops": tf, "frac_tensor": tf / peaks["tflops"],
            "hbm_gbs": gb, "frac_hbm": gb / peaks["hbm_gbs"],
            "note": "dense GEMM FLOPs only (the long convolution's own arithmetic is extra work, not counted)"}


def make_items(lens, batches, seed, index_of):
    from deepchopper_b200.encode import MAX_TOKENS
    lens = np.minimum(lens, MAX_TOKENS - 1)
    return [(b,) + synth_batch(lens[b.rows], seed, index_of[id(b)]) for b in batches]


def run_pass(model, items, steps, warmup, dev, world=1, profile=True):
    """Upload `items`, run `warmup` + `steps` passes; returns (ms_total, prof, launches, pipe)."""
    from deepchopper_b200 import _native
    from deepchopper_b200.predict import DevicePipeline
    pipe = DevicePipeline(model)
    pipe.upload_items(items)
    ctx = _native.torch_context(dev)
    for _ in range(warmup):
        pipe.run_all()
    torch.cuda.synchronize(dev)
    if profile:
        ctx.profile(True)
        ctx.profile_read(reset=True)
    launches0 = ctx.launches
    barrier(world)
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        pipe.run_all()
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    prof = ctx.profile_read(reset=True) if profile else {}
    if profile:
        ctx.profile(False)
    return ms, prof, ctx.launches - launches0, pipe


def extra_stress(model, args, dev, peaks):
    """BASELINE configs[3] (long-read stress): a bounded sample of its read-length distribution through the same path."""
    from deepchopper_b200.predict import plan_batches
    n = args.stress_reads
    lens = synth_lengths(n, args.seed + 3, "stress")
    batches = plan_batches(lens, token_budget=args.token_budget)
    index_of = {id(b): i for i, b in enumerate(batches)}
    items = make_items(lens, batches, args.seed + 3, index_of)
    steps = 2
    ms, prof, launches, pipe = run_pass(model, items, steps, 1, dev)
    bases = int(lens.sum())
    tokens = int(sum(b.rows.size * b.Lrow for b in batches))
    from deepchopper_b200 import _native
    conv = conv_work(batches, _native.torch_context(dev).get_option("fft_min_len"))
    kernels = kernel_table(prof, ms, tokens * steps, bases * steps, peaks, conv, steps)
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"])
    del pipe
    torch.cuda.empty_cache()
    return {"workload": f"configs[3] long-read stress: a sample of {n} synthetic reads, length U[16384, 32766] "
                        f"(the config names 50k reads)", "value": bases * steps / (ms / 1e3), "unit": UNIT,
            "reads_per_sec": n * steps / (ms / 1e3), "padded_tokens_per_sec": tokens * steps / (ms / 1e3),
            "ms_per_step": ms / steps, "steps": steps, "batches_per_step": len(batches), "gpu_launches": int(launches),
            "roofline": {"kernel": dom, **{k: kernels[dom].get(k) for k in ("bound", "achieved", "peak", "unit", "frac")},
                         "traffic": None, "avg_launch_ms": kernels[dom]["ms_total"] / max(1, kernels[dom]["launches"])},
            "hyena_layer": layer_summary(kernels, tokens * steps, peaks),
            "conv_share": sum(kernels[k]["share"] for k in ("fft_conv", "toeplitz_conv") if k in kernels),
            "kernels": {k: {kk: v.get(kk) for kk in ("ms_total", "launches", "share", "bound", "frac", "ns_per_token_layer",
                                                     "fp32_tflops") if v.get(kk) is not None} for k, v in kernels.items()}}


def extra_reference_batching(model, args, dev, n_reads: int = 1536):
    """The product path in the REFERENCE's batching (configs[0] style: FASTQ order, batch 16, left-pad to the batch
    maximum; only_fq.py:198-202 + tokenizer.py:34-93) on the read set the GPU eager baseline uses: what a user gets from
    `predict` without --bucket.  `value`: the product's route -- those batches packed into launches in which every row
    keeps its own batch's pad count (predict.group_batches, dcb200_encode_batch_rows); `one_batch_per_launch`: the same
    batches launched one by one (device-resident, and through dcb200_predict_batch_host on pinned host buffers)."""
    from deepchopper_b200 import ops, synth  # noqa: F401
    from deepchopper_b200.predict import Batch, HostPipeline, group_batches
    from deepchopper_b200.smooth import smooth_chop_device
    rng = np.random.default_rng(args.seed)
    lens = synth.read_lengths(rng, n_reads, hi=8000)
    batches = []
    for i in range(0, n_reads, 16):
        rows = np.arange(i, min(i + 16, n_reads))
        lpad = int(lens[rows].max()) + 1
        batches.append(Batch(rows, lpad, (lpad + 127) // 128 * 128))
    index_of = {id(b): i for i, b in enumerate(batches)}
    items = make_items(lens, batches, args.seed + 5, index_of)
    steps = 3
    ms, _, launches, pipe = run_pass(model, items, steps, 1, dev, profile=False)
    del pipe
    hp = HostPipeline(model)
    hp.pack_items(items)
    hp.run_all()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        hp.run_all()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    del hp
    # grouped launches: one device blob [all sequences | all quality strings] in read order, per-launch index tensors
    tot = int(lens.sum())
    blob = np.empty(2 * tot, dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    for b, buf, so, qo, ln in items:
        t = int(ln.sum())
        for k, r in enumerate(b.rows):
            blob[off[r]:off[r] + ln[k]] = buf[so[k]:so[k] + ln[k]]
            blob[tot + off[r]:tot + off[r] + ln[k]] = buf[qo[k]:qo[k] + ln[k]]
    d_blob = torch.from_numpy(blob).to(dev)
    groups = group_batches(batches, args.token_budget)
    prepared = []
    for g in groups:
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)  # noqa: E731
        ln = t(lens[g.rows], np.int32)
        lp = t(g.lpad, np.int32)
        st = torch.arange(g.rows.size, dtype=torch.int64, device=dev) * g.Lrow + (lp.to(torch.int64) - 1) - ln.to(torch.int64)
        prepared.append((g, t(off[g.rows], np.int64), t(off[g.rows] + tot, np.int64), ln, lp, st))

    def run_groups():
        for g, so, qo, ln, lp, st in prepared:
            tok, qual = torch.ops.dcb200.encode_rows(d_blob, so, qo, ln, lp, int(g.Lpad), int(g.Lrow))
            _, labels = model.forward_tokens(tok, qual, False, True)
            smooth_chop_device(labels.view(-1), st, ln)

    # the same launches through the C ABI on pinned host buffers (dcb200_predict_batch_host_rows)
    hg = HostPipeline(model)
    g_items = []
    for g in groups:
        ln = lens[g.rows].astype(np.int32)
        so = (np.cumsum(ln) - ln).astype(np.int64)
        buf = np.concatenate([blob[off[r]:off[r] + lens[r]] for r in g.rows] + [blob[tot + off[r]:tot + off[r] + lens[r]] for r in g.rows])
        g_items.append((g, buf, so, so + int(ln.sum()), ln))
    hg.pack_items(g_items)
    hg.run_all()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        hg.run_all()
    torch.cuda.synchronize(dev)
    e2e_g = time.perf_counter() - t0
    del hg
    run_groups()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run_groups()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_g = e0.elapsed_time(e1)
    torch.cuda.empty_cache()
    bases = int(lens.sum())
    tokens = int(sum(b.rows.size * b.Lrow for b in batches))
    tokens_g = int(sum(g.rows.size * g.Lrow for g in groups))
    return {"workload": f"{n_reads} synthetic reads (the GPU eager baseline's read set), FASTQ order, batch 16, left-pad "
                        "to the batch maximum: the reference's own batching",
            "value": bases * steps / (ms_g / 1e3), "unit": UNIT, "launches_per_step": len(groups),
            "batches_per_step": len(batches), "steps": steps, "ms_per_step": ms_g / steps,
            "padded_tokens_per_sec": tokens_g * steps / (ms_g / 1e3), "padding_overhead": tokens_g / max(1, bases),
            "route": "batches packed into launches, every row left-padded as in its own batch (group_batches)",
            "e2e": {"value": bases * steps / e2e_g, "unit": UNIT, "api": "dcb200_predict_batch_host_rows, one call per launch"},
            "one_batch_per_launch": {"value": bases * steps / (ms / 1e3), "unit": UNIT, "ms_per_batch": ms / steps / len(batches),
                                     "gpu_launches": int(launches), "padding_overhead": tokens / max(1, bases),
                                     "e2e": {"value": bases * steps / e2e_s, "unit": UNIT,
                                             "api": "dcb200_predict_batch_host, one call per batch"}}}


def extra_smooth_only(args, dev, peaks):
    """BASELINE configs[4]: the GPU smooth / interval / chop-coordinate pass alone over 10M reads of precomputed int8
    labels (SURVEY §8d.5 recipe, generated on the device in chunks of 1M reads).  Device-resident and H2D-inclusive."""
    import ctypes as C
    from deepchopper_b200 import _native
    from deepchopper_b200._native import ChopParams, check, lib
    R_total, chunk = args.smooth_reads, 1_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed + 4)
    ctx = _native.torch_context(dev)
    params = ChopParams.default()
    ap = int(params.approved_interval_number)
    chunks = []
    n_bases = 0
    for c0 in range(0, R_total, chunk):
        R = min(chunk, R_total - c0)
        lens = torch.clamp(torch.round(torch.exp(torch.randn(R, generator=g, device=dev) * 0.6 + np.log(1000.0))), 200, 8192).to(torch.int32)
        starts = torch.zeros(R, dtype=torch.int64, device=dev)
        starts[1:] = torch.cumsum(lens.to(torch.int64), 0)[:-1]
        N = int(lens.sum().item())
        rid = torch.repeat_interleave(torch.arange(R, device=dev, dtype=torch.int32), lens.to(torch.int64)).long()
        pos = torch.arange(N, device=dev, dtype=torch.int64) - starts[rid]
        ln = lens[rid].to(torch.int64)
        labels = torch.rand(N, generator=g, device=dev) < 0.02
        term = torch.rand(R, generator=g, device=dev) < 0.6
        tlen = torch.randint(30, 121, (R,), generator=g, device=dev)
        in_run = term[rid] & (pos >= ln - tlen[rid])
        for _ in range(2):  # internal runs
            has = torch.rand(R, generator=g, device=dev) < 0.3
            rl = torch.randint(30, 121, (R,), generator=g, device=dev)
            st = (torch.rand(R, generator=g, device=dev) * (lens - 150).clamp(min=1)).to(torch.int64)
            in_run |= has[rid] & (pos >= st[rid]) & (pos < st[rid] + rl[rid])
        keep1 = torch.rand(N, generator=g, device=dev) >= 0.05
        labels = torch.where(in_run, keep1, labels).to(torch.int8)
        del rid, pos, ln, in_run, keep1
        outs = (torch.empty(R, dtype=torch.int32, device=dev), torch.empty((R, ap, 2), dtype=torch.int32, device=dev),
                torch.empty(R, dtype=torch.int32, device=dev), torch.empty((R, ap + 1, 2), dtype=torch.int32, device=dev),
                torch.empty(R, dtype=torch.uint8, device=dev))
        chunks.append((labels, starts, lens, outs))
        n_bases += N
    torch.cuda.synchronize(dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def run_chunk(labels, starts, lens, outs):
        check(lib().dcb200_smooth_chop(ctx.handle, p(labels), labels.numel(), p(starts), p(lens), None, lens.numel(),
                                       C.byref(params), *[p(o) for o in outs]))

    for ch in chunks:
        run_chunk(*ch)
    torch.cuda.synchronize(dev)
    ctx.profile(True)
    ctx.profile_read(reset=True)
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for ch in chunks:
            run_chunk(*ch)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)
    kms, kcnt = prof.get("smooth_chop", (ms * steps, len(chunks) * steps))
    ach = n_bases * steps / (kms / 1e3) / 1e9
    adapters = int(sum(int(ch[3][0].sum().item()) for ch in chunks))
    # parity spot check of one chunk's head against the C oracle (bit-exact), outside the timed region
    from oracle import cref
    lab0, st0, ln0, out0 = chunks[0]
    k = min(20000, int(ln0.numel()))
    nb0 = int((st0[k - 1] + ln0[k - 1]).item())
    chk = cref.load().smooth_chop(lab0[:nb0].cpu().numpy(), st0[:k].cpu().numpy(), ln0[:k].cpu().numpy())
    ok = all(np.array_equal(out0[i][:k].cpu().numpy(), chk[nm]) for i, nm in ((0, "n_adapter"), (2, "n_keep"), (4, "action")))
    # H2D-inclusive: pinned host labels of ONE chunk -> device -> kernel -> coordinate tables back to the host
    host_lab = lab0.cpu().pin_memory()
    host_outs = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in out0]
    dev_lab = torch.empty_like(lab0)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(2):
        dev_lab.copy_(host_lab, non_blocking=True)
        run_chunk(dev_lab, st0, ln0, out0)
        for h, o in zip(host_outs, out0):
            h.copy_(o, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / 2
    n0 = int(lab0.numel())
    res = {"workload": f"configs[4] post-processing only: {R_total} reads of precomputed int8 labels ({n_bases} bases), "
                       f"log-normal lengths, planted adapter runs, resident in HBM, {len(chunks)} launches of <= {chunk} reads",
           "value": n_bases / (ms / 1e3), "unit": UNIT, "reads_per_sec": R_total / (ms / 1e3), "ms_per_step": ms, "steps": steps,
           "adapters_found": adapters, "parity_vs_c_oracle_first_reads": {"reads": k, "bit_exact": bool(ok)},
           "roofline": {"kernel": "smooth_chop", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": None, "avg_launch_ms": kms / max(1, kcnt),
                        "note": "1 B per base (int8 label in); coordinate tables out are O(reads)"},
           "e2e": {"value": n0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n0,
                   "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_outs)),
                   "sample": f"one chunk of {int(ln0.numel())} reads from pinned host memory: H2D of the labels + kernel + "
                             "D2H of the coordinate tables"}}
    del chunks, host_lab, dev_lab
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, local, world):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; deepchopper_b200 has no CPU fallback")
    from deepchopper_b200.init_weights import random_state_dict
    from deepchopper_b200.model import DeepChopper
    from deepchopper_b200.predict import HostPipeline, plan_batches, shard_batches

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    sd = random_state_dict(0)
    model = DeepChopper.from_state_dict(sd, device=dev)
    # ---- ONE global job; this rank's share of its batches --------------------------------------------------
    n_global = args.reads if args.reads else (100_000 if world == 1 else 2_000_000)
    lens = synth_lengths(n_global, args.seed, args.workload)
    batches = plan_batches(lens, token_budget=args.token_budget)
    index_of = {id(b): i for i, b in enumerate(batches)}
    mine = shard_batches(batches, rank, world)
    items = make_items(lens, mine, args.seed, index_of)
    bases = int(sum(int(it[4].sum()) for it in items))
    reads = int(sum(it[0].rows.size for it in items))
    padded_tokens = int(sum(b.rows.size * b.Lrow for b in mine))

    sampler = ClockSampler(local)
    sampler.start()   # brackets warm-up + timed region
    ms_rank, prof, launches, pipe = run_pass(model, items, args.steps, args.warmup, dev, world)
    barrier(world)
    clocks = sampler.stop()
    rank_ms = gather_over_ranks(ms_rank, world, dev)
    ms_total = max(rank_ms)
    bases_all = sum(gather_over_ranks(bases, world, dev))
    reads_all = sum(gather_over_ranks(reads, world, dev))
    rank_tokens = gather_over_ranks(padded_tokens, world, dev)
    tokens_all = sum(rank_tokens)
    value = bases_all * args.steps / (ms_total / 1e3)
    del pipe
    torch.cuda.empty_cache()

    # ---- e2e through the C ABI on pinned host buffers ------------------------------------------------
    hp = HostPipeline(model)
    hp.pack_items(items)
    h2d, d2h = hp.bytes_per_pass()
    hp.run_all()  # warm
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier(world)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hp.run_all()
    torch.cuda.synchronize(dev)
    e2e_s = max(gather_over_ranks(time.perf_counter() - t0, world, dev))
    e2e_value = bases_all * e2e_steps / e2e_s
    h2d = int(sum(gather_over_ranks(h2d, world, dev)))  # whole-job bytes per step, like `value`
    d2h = int(sum(gather_over_ranks(d2h, world, dev)))
    del hp

    # ---- per-kernel roofline (CUDA events on the launching stream, timed region above; this rank) ----
    from deepchopper_b200 import _native
    conv = conv_work(mine, _native.torch_context(dev).get_option("fft_min_len"))
    kernels = kernel_table(prof, ms_rank, padded_tokens * args.steps, bases * args.steps, peaks, conv, args.steps)
    layer = layer_summary(kernels, padded_tokens * args.steps, peaks)
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    roofline = None
    if dom:
        k = kernels[dom]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture of the
            # SAME build (profiles/r02_*), per token, scaled to this run's average launch
            ent = json.load(open(tpath)).get("kernels", {}).get(dom)
            if ent:
                launches_per_step = k["launches"] / max(1, args.steps)
                tokens_per_launch = padded_tokens * (N_LAYERS if dom in PER_LAYER else 1) / max(1.0, launches_per_step)
                traffic = ent["dram_bytes_per_token"] * tokens_per_launch
        roofline = {"kernel": dom, "bound": k.get("bound"), "achieved": k.get("achieved"), "peak": k.get("peak"),
                    "unit": k.get("unit"), "frac": k.get("frac"), "traffic": traffic, "peak_source": peaks["source"],
                    "avg_launch_ms": k["ms_total"] / max(1, k["launches"])}
    if rank != 0:
        return
    cfg_name = {"configs1": "configs[1]" if world == 1 else "configs[2]", "stress": "configs[3] long-read stress"}[args.workload]
    dist_txt = ("log-normal length (median 1 kb, clipped to [200, 8000])" if args.workload == "configs1"
                else "length U[16384, 32766]")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{cfg_name}: ONE global job of {n_global} synthetic reads, {dist_txt}, random-init "
                               "HyenaDNA-small-32k + DeepChopper head" +
                               (f"; batches dealt to {world} ranks by shard_batches (fixed job: strong scaling)" if world > 1 else ""),
                   "global_reads": int(n_global), "global_batches": len(batches),
                   "batching": f"length-bucketed, <= {args.token_budget} padded tokens per batch, left-pad to batch max",
                   "batches_per_step": len(mine), "l2": "inputs_larger_than_l2", "parallelism": f"read-sharded x{world}"},
        "reads_per_sec": reads_all * args.steps / (ms_total / 1e3),
        "padded_tokens_per_sec": tokens_all * args.steps / (ms_total / 1e3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "dcb200_predict_batch_host (pinned host buffers)",
                "excludes": "FASTQ parse / index and any result file (see tools/bench_cli.py for file to file)"},
        "gpu_launches": int(launches),
        "sharding": {"tokens_per_rank": [int(t) for t in rank_tokens],
                     "token_imbalance_max_over_mean": max(rank_tokens) / (sum(rank_tokens) / len(rank_tokens)),
                     "rank_ms_max": max(rank_ms), "rank_ms_min": min(rank_ms)},
        "roofline": roofline,
        "hyena_layer": layer,
        "kernels": kernels,
    }
    if world == 1 and not args.no_extras:
        extras = {}
        for name, fn in (("stress", lambda: extra_stress(model, args, dev, peaks)),
                         ("smooth_only", lambda: extra_smooth_only(args, dev, peaks)),
                         ("reference_batching", lambda: extra_reference_batching(model, args, dev))):
            try:
                extras[name] = fn()
            except Exception as e:  # noqa: BLE001
                extras[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        line["extra_configs"] = extras
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(sd, args.seed, dev)
        except Exception as e:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": repr(e)[:300]}
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
    ap.add_argument("--reads", type=int, default=0, help="reads of the global job (default 100k at N=1, 2M at N>1)")
    ap.add_argument("--token-budget", type=int, default=1024 * 1024)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-reads", type=int, default=192)
    ap.add_argument("--ref-reads", type=int, default=96)
    ap.add_argument("--stress-reads", type=int, default=1024)
    ap.add_argument("--smooth-reads", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra_configs and the GPU eager baseline (N=1 only)")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "stress"],
                    help="configs1 = BASELINE configs[1] / [2] (the headline); stress = configs[3], 16-32 kb reads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        # rank 0 alone times the CPU reference; other ranks exit 0 without work (no process group needed)
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, local, world = dist_setup(args.gpus)
    try:
        run_ours(args, rank, local, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
