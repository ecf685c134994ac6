"""``deepchopper`` CLI mirror (deepchopper/cli.py:66-198): ``predict`` and ``chop`` with the reference's
flags and defaults.  ``python -m deepchopper_b200.cli predict x.fastq -o predictions`` then
``python -m deepchopper_b200.cli chop predictions/0 x.fastq``; ``predict --chop`` does both in one pass without
prediction files (deepchopper_b200/fused.py)."""
from __future__ import annotations

import argparse
import os
import queue
import sys
import threading
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


def gather_ranges(buf: np.ndarray, off: np.ndarray, ln: np.ndarray) -> np.ndarray:
    """Concatenation of buf[off[k] : off[k] + ln[k]] (vectorised)."""
    ln = ln.astype(np.int64)
    tot = int(ln.sum())
    if tot == 0:
        return np.zeros(0, dtype=buf.dtype)
    starts = np.cumsum(ln) - ln
    return buf[np.repeat(off.astype(np.int64) - starts, ln) + np.arange(tot, dtype=np.int64)]


def plan_rank_batches(lens: np.ndarray, args, rank: int, world: int):
    """[(file index, Batch)] of one rank.  FASTQ-order mode mirrors the reference: the per-device batch is
    batch_size // world and must divide (only_fq.py:198-202), and Lightning's DistributedSampler deals read i to rank
    i % world (its padding of the last round by repeating reads is not reproduced: chop de-duplicates by id anyway).
    --bucket: length-bucketed batches dealt greedily by padded tokens (predict.shard_batches)."""
    from .predict import Batch, plan_batches, shard_batches
    if args.bucket:
        batches = plan_batches(lens, token_budget=args.token_budget)
        if args.limit_batches:
            batches = batches[: args.limit_batches]
        index_of = {id(b): i for i, b in enumerate(batches)}
        return [(index_of[id(b)], b) for b in shard_batches(batches, rank, world)]
    if args.batch_size % world != 0:
        raise ValueError(f"Batch size {args.batch_size} must be divisible by the number of devices {world}")
    bs = max(1, args.batch_size // world)
    mine = np.arange(rank, lens.size, world)
    out = []
    for k, i in enumerate(range(0, mine.size, bs)):
        rows = mine[i:i + bs]
        lpad = int(lens[rows].max()) + 1
        out.append((k, Batch(rows.copy(), lpad, (lpad + 127) // 128 * 128)))
    return out[: args.limit_batches] if args.limit_batches else out


class BatchWriter:
    """Writes prediction batches on a background thread: the device -> host copy of batch i (pinned, on its own stream,
    after the batch's event) and ``torch.save`` / the sidecar write run while the GPU computes the batches behind it.
    At most ``depth`` batches are in flight."""

    def __init__(self, device, depth: int = 2):
        import torch
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.err = None
        self.stream = torch.cuda.Stream(device=device)
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        import torch
        while True:
            job = self.q.get()
            if job is None:
                return
            ev, fn = job
            try:
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(ev)
                    fn()
            except BaseException as e:  # noqa: BLE001
                self.err = e

    def submit(self, fn):
        import torch
        if self.err:
            raise self.err
        ev = torch.cuda.Event()
        ev.record()
        self.q.put((ev, fn))

    def close(self):
        self.q.put(None)
        self.t.join()
        if self.err:
            raise self.err


def _predict_worker(rank: int, world: int, args, shared):
    import torch
    from . import encode, writer
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    model = _load_model(args, dev)
    # the parent read, inflated and indexed the FASTQ ONCE; ranks see it through shared memory and upload only the bytes
    # of their own reads
    ix = encode.FastqIndex(*[shared[k].numpy() for k in ("buf", "name_off", "name_len", "head_len", "seq_off", "seq_len",
                                                          "qual_off", "qual_len")])
    n = int(shared["n"])
    lens = np.minimum(ix.seq_len[:n].astype(np.int64), encode.MAX_TOKENS - 1)
    truncated = ix.seq_len[:n] >= encode.MAX_TOKENS                      # tokenizer.py:154-156
    mine = plan_rank_batches(lens, args, rank, world)
    t0 = time.time()
    bw = BatchWriter(dev)
    from . import ops  # noqa: F401  (registers torch.ops.dcb200.*)
    from .predict import Launch, group_batches
    if args.bucket:
        launches = [Launch(b.rows, np.full(b.rows.size, b.Lpad, np.int32), b.Lpad, b.Lrow, [(k, 0, b.rows.size, b)])
                    for k, (_, b) in enumerate(mine)]
    else:
        # the reference's FASTQ-order batches of --batch-size reads, many per launch: every row keeps the pad count of
        # its own batch (the pads are semantic), the kernels get full row tiles, and one .pt per reference batch is written
        launches = group_batches([b for _, b in mine], args.token_budget)
    for g in launches:
        ln = lens[g.rows]
        tot = int(ln.sum())
        host = torch.empty(2 * tot, dtype=torch.uint8).pin_memory()
        hb = host.numpy()
        hb[:tot] = gather_ranges(ix.buf, ix.seq_off[g.rows], ln)
        hb[tot:] = gather_ranges(ix.buf, ix.qual_off[g.rows], ln)
        so = np.cumsum(ln) - ln
        blob = host.to(dev, non_blocking=True)
        tok, qual = torch.ops.dcb200.encode_rows(blob, torch.from_numpy(so).to(dev), torch.from_numpy(so + tot).to(dev),
                                                 torch.from_numpy(ln.astype(np.int32)).to(dev),
                                                 torch.from_numpy(g.lpad).to(dev), int(g.Lpad), int(g.Lrow))
        if args.compact:   # labels only, bit-packed: read back by this package's `chop`
            logits, labels = None, model.forward_tokens(tok, qual, False, True)[1]
        else:
            logits, labels = model.forward_tokens(tok, qual, True, False)[0], None
        for pos, r0, r1, b in g.members:
            idx = mine[pos][0]
            lnb = ln[r0:r1]
            if args.compact:
                ids = [ix.name(r) for r in b.rows]
                bw.submit(lambda labels=labels[r0:r1], lnb=lnb, b=b, idx=idx, ids=ids, host=host: writer.write_batch_compact(
                    args.output, rank, idx, labels, lnb, b.Lpad, ids, truncated[b.rows]))
                continue
            id_rows = encode.id_rows(ix, b.rows, truncated[b.rows])
            bw.submit(lambda logits=logits[r0:r1], tok=tok[r0:r1], qual=qual[r0:r1], id_rows=id_rows, lnb=lnb, b=b, idx=idx,
                      host=host: writer.write_batch(args.output, rank, idx, writer.batch_dict(logits, tok, qual, id_rows, lnb, b.Lpad)))
    bw.close()
    if args.verbose:
        print(f"[rank {rank}] {len(mine)} batches in {len(launches)} launches, {time.time() - t0:.2f}s", file=sys.stderr)


def cmd_predict(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("deepchopper_b200 predict needs a B200 (no CPU fallback)")
    world = max(1, min(args.gpus or 1, torch.cuda.device_count()))       # cli.py:126-129
    if args.chop:
        from .chop import params_from_cli
        from .fused import predict_chop_fastq
        p = params_from_cli(args.smooth_window, args.min_interval_size, args.approved_intervals, args.max_process_intervals,
                            args.min_read_length, args.output_chopped, args.chop_type)
        # the FASTQ is read, inflated and indexed (native code, GIL released) while CUDA and the weights come up
        from . import encode
        loaded = {}

        def _load():
            try:
                buf = encode.read_fastq_bytes(args.data_path)
                loaded["v"] = (buf, encode.index_fastq(buf))
            except BaseException as e:  # noqa: BLE001
                loaded["e"] = e

        th = threading.Thread(target=_load)
        th.start()
        model = _load_model(args, torch.device("cuda", 0))
        th.join()
        if "e" in loaded:
            raise loaded["e"]
        out, npred, nrec = predict_chop_fastq(args.data_path, model, p, args.chop_output, args.token_budget,
                                              None if args.bucket else args.batch_size, args.threads,
                                              args.compression_level, args.max_sample, args.verbose, loaded["v"])
        print(f"Wrote {nrec} records to {out} ({npred} predictions)")
        return
    from . import encode
    buf = encode.read_fastq_bytes(args.data_path)
    ix = encode.index_fastq(buf)
    n = len(ix) if not args.max_sample else min(len(ix), args.max_sample)
    shared = {k: torch.from_numpy(np.ascontiguousarray(getattr(ix, k))) for k in
              ("buf", "name_off", "name_len", "head_len", "seq_off", "seq_len", "qual_off", "qual_len")}
    shared["n"] = n
    if world == 1:
        _predict_worker(0, 1, args, shared)
    else:
        import torch.multiprocessing as mp
        for k, v in shared.items():
            if torch.is_tensor(v):
                v.share_memory_()
        mp.spawn(_predict_worker, args=(world, args, shared), nprocs=world, join=True)


def cmd_chop(args):
    from .chop import chop_fastq, params_from_cli
    p = params_from_cli(args.smooth_window, args.min_interval_size, args.approved_intervals, args.max_process_intervals,
                        args.min_read_length, args.output_chopped, args.chop_type)
    out, npred, nrec = chop_fastq(args.predicts, args.fq, p, args.output, args.max_batch, threads=args.threads,
                                  level=args.compression_level, verbose=args.verbose,
                                  chunk_bytes=(args.chunk_mb << 20) if args.chunk_mb else None)
    print(f"Wrote {nrec} records to {out} ({npred} predictions)")


def _add_chop_flags(ch, output_flag: str):
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
    ch.add_argument(*output_flag.split(), default=None, dest="chop_output" if "chop" in output_flag else "output",
                    help="output prefix of the chopped FASTQ")


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
    pr.add_argument("--bucket", action="store_true",
                    help="length-bucketed batches instead of FASTQ-order batches of --batch-size reads.  Left pads are "
                         "semantic (the model has no attention mask), so logits depend on the batching: only single-GPU "
                         "FASTQ-order mode reproduces the batches of a reference run with the same -b")
    pr.add_argument("--token-budget", type=int, default=512 * 1024)
    pr.add_argument("--compact", action="store_true",
                    help="write bit-packed label sidecars (1 bit per base) instead of the reference's .pt dicts "
                         "(28 bytes per token); only this package's `chop` reads them")
    pr.add_argument("--chop", action="store_true",
                    help="predict and chop in one pass: no prediction files, the chopped FASTQ is written directly "
                         "(the chop flags below apply)")
    _add_chop_flags(pr, "--chop-output")
    pr.set_defaults(fn=cmd_predict)
    ch = sub.add_parser("chop", help="smooth predictions and cut reads (cli.py:155-198 -> deepchopper-chop)")
    ch.add_argument("predicts", nargs="+")
    ch.add_argument("fq")
    _add_chop_flags(ch, "--output -o")
    ch.add_argument("--max-batch", type=int, default=None)
    ch.add_argument("--chunk-mb", type=int, default=0,
                    help="stream the FASTQ in pieces of this many MB of text (src/bin/predict.rs:275-316 streams 10 000 "
                         "records at a time); default: whole file below 2 GB, 256 MB pieces above")
    ch.add_argument("--verbose", "-v", action="store_true")     # cli.py:155-198 / src/bin/predict.rs:76-77
    ch.set_defaults(fn=cmd_chop)
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main()
