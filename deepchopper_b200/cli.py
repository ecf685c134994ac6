"""``deepchopper`` CLI mirror (deepchopper/cli.py:66-198): ``predict`` and ``chop`` with the reference's
flags and defaults.  ``python -m deepchopper_b200.cli predict x.fastq -o predictions`` then
``python -m deepchopper_b200.cli chop predictions/0 x.fastq``."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

MODELS = {"rna002": "yangliz5/deepchopper", "rna004": "yangliz5/deepchopper-rna004"}   # cli.py:96-103


def _load_model(args, device):
    from .model import DeepChopper
    if args.random_init:
        from .init_weights import random_state_dict
        return DeepChopper.from_state_dict(random_state_dict(args.seed), device=device)
    if args.checkpoint:
        return DeepChopper.from_checkpoint(args.checkpoint, device=device)
    return DeepChopper.from_pretrained(MODELS[args.model], device=device)


def _predict_worker(rank: int, world: int, args):
    import torch
    from . import encode, writer
    from .predict import Batch, DevicePipeline, plan_batches, shard_batches
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    model = _load_model(args, dev)
    buf = encode.read_fastq_bytes(args.data_path)
    ix = encode.index_fastq(buf)
    n = len(ix)
    if args.max_sample:
        n = min(n, args.max_sample)
    lens = np.minimum(ix.seq_len[:n].astype(np.int64), encode.MAX_TOKENS - 1)
    truncated = ix.seq_len[:n] >= encode.MAX_TOKENS                      # tokenizer.py:154-156
    if args.bucket:
        batches = plan_batches(lens, token_budget=args.token_budget)
    else:
        bs = max(1, args.batch_size // world)                           # only_fq.py:198-202
        batches = plan_batches(lens, token_budget=1 << 62, max_rows=bs, sort=False)
    if args.limit_batches:
        batches = batches[: args.limit_batches]
    mine = [(i, b) for i, b in enumerate(batches) if i % world == rank] if not args.bucket else \
        list(enumerate(shard_batches(batches, rank, world)))
    pipe = DevicePipeline(model)
    pipe.upload(buf, ix.seq_off[:n], ix.qual_off[:n], lens, [b for _, b in mine])
    t0 = time.time()
    for (idx, b), item in zip(mine, pipe.items):
        from .encode import encode_batch_device
        _, so, qo, ln, st = item
        tok, qual = encode_batch_device(pipe.blob, so, qo, ln, b.Lpad, None, b.Lrow)
        if args.compact:   # labels only, bit-packed: read back by this package's `chop`
            _, labels = model.forward_tokens(tok, qual, False, True)
            writer.write_batch_compact(args.output, rank, idx, labels, lens[b.rows], b.Lpad, [ix.name(r) for r in b.rows],
                                       truncated[b.rows])
            continue
        logits, _ = model.forward_tokens(tok, qual, True, False)
        d = writer.batch_dict(logits, tok, qual, encode.id_rows(ix, b.rows, truncated[b.rows]), lens[b.rows], b.Lpad)
        writer.write_batch(args.output, rank, idx, d)
    if args.verbose:
        print(f"[rank {rank}] {len(mine)} batches in {time.time() - t0:.2f}s", file=sys.stderr)


def cmd_predict(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("deepchopper_b200 predict needs a B200 (no CPU fallback)")
    world = max(1, min(args.gpus or 1, torch.cuda.device_count()))       # cli.py:126-129
    if world == 1:
        _predict_worker(0, 1, args)
    else:
        import torch.multiprocessing as mp
        mp.spawn(_predict_worker, args=(world, args), nprocs=world, join=True)


def cmd_chop(args):
    from .chop import chop_fastq, params_from_cli
    p = params_from_cli(args.smooth_window, args.min_interval_size, args.approved_intervals, args.max_process_intervals,
                        args.min_read_length, args.output_chopped, args.chop_type)
    out, npred, nrec = chop_fastq(args.predicts, args.fq, p, args.output, args.max_batch, threads=args.threads,
                                  level=args.compression_level, verbose=args.verbose)
    print(f"Wrote {nrec} records to {out} ({npred} predictions)")


def build_parser():
    ap = argparse.ArgumentParser(prog="deepchopper-b200", description="B200-native DeepChopper predict / chop")
    sub = ap.add_subparsers(dest="cmd", required=True)
    pr = sub.add_parser("predict", help="per-base adapter prediction (cli.py:66-152)")
    pr.add_argument("data_path")
    pr.add_argument("--gpus", "-g", type=int, default=0)
    pr.add_argument("--output", "-o", default="predictions")
    pr.add_argument("--batch-size", "-b", type=int, default=12)
    pr.add_argument("--workers", "-w", type=int, default=0)
    pr.add_argument("--model", "-m", default="rna002", choices=sorted(MODELS))
    pr.add_argument("--limit-batches", type=int, default=None)
    pr.add_argument("--max-sample", type=int, default=None)
    pr.add_argument("--verbose", "-v", action="store_true")
    pr.add_argument("--checkpoint", default=None, help="local .ckpt / .safetensors / state-dict file")
    pr.add_argument("--random-init", action="store_true", help="seeded random weights (no network for the hub)")
    pr.add_argument("--seed", type=int, default=0)
    pr.add_argument("--bucket", action="store_true", help="length-bucketed batches instead of FASTQ-order batches")
    pr.add_argument("--token-budget", type=int, default=512 * 1024)
    pr.add_argument("--compact", action="store_true",
                    help="write bit-packed label sidecars (1 bit per base) instead of the reference's .pt dicts "
                         "(28 bytes per token); only this package's `chop` reads them")
    pr.set_defaults(fn=cmd_predict)
    ch = sub.add_parser("chop", help="smooth predictions and cut reads (cli.py:155-198 -> deepchopper-chop)")
    ch.add_argument("predicts", nargs="+")
    ch.add_argument("fq")
    ch.add_argument("--smooth-window", type=int, default=21)
    ch.add_argument("--min-interval-size", type=int, default=13)
    ch.add_argument("--approved-intervals", type=int, default=20)
    ch.add_argument("--max-process-intervals", type=int, default=4)
    ch.add_argument("--min-read-length", type=int, default=20)
    ch.add_argument("--output-chopped", action="store_true")
    ch.add_argument("--chop-type", default="all", choices=["terminal", "internal", "all"])
    ch.add_argument("--threads", "-t", type=int, default=2)
    ch.add_argument("--compression-level", type=int, default=6,
                    help="BGZF deflate level 1-9 (6 = the reference's default); 0 = Huffman-only, ~8x faster, ~5 %% larger")
    ch.add_argument("--output", "-o", default=None)
    ch.add_argument("--max-batch", type=int, default=None)
    ch.add_argument("--verbose", "-v", action="store_true")     # cli.py:155-198 / src/bin/predict.rs:76-77
    ch.set_defaults(fn=cmd_chop)
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main()
