"""-m gpu end-to-end parity: prediction directory -> chop output (decompressed bytes identical to the
oracle's restatement of deepchopper-chop), CLI predict -> .pt layout, host vs device pipelines."""
import gzip
import os

import numpy as np
import pytest
import torch

from deepchopper_b200 import synth
from oracle import hyena_ref as H
from oracle import smooth_ref as S

pytestmark = pytest.mark.gpu


def _planted_batches(recs, rng, batch=16):
    from deepchopper_b200 import encode, writer
    buf = np.frombuffer(synth.fastq_text(recs), dtype=np.uint8)
    ix = encode.index_fastq(buf)
    lens_all = ix.seq_len.astype(np.int64)
    lab, starts, _ = synth.planted_labels(rng, lens_all)
    table = np.zeros(256, np.uint8)
    for k, v in {ord("A"): 7, ord("C"): 8, ord("G"): 9, ord("T"): 10, ord("N"): 11}.items():
        table[k] = v
    dicts = []
    for i in range(0, len(recs), batch):
        rows = np.arange(i, min(len(recs), i + batch))
        lens = lens_all[rows]
        Lpad = int(lens.max()) + 1
        tok = torch.full((rows.size, Lpad), 4, dtype=torch.uint8)
        qual = torch.zeros((rows.size, Lpad))
        logits = torch.zeros((rows.size, Lpad, 2))
        for k, r in enumerate(rows):
            n = int(lens[k])
            tok[k, Lpad - 1 - n:Lpad - 1] = torch.from_numpy(table[np.frombuffer(ix.seq(r), dtype=np.uint8)])
            tok[k, Lpad - 1] = 1
            l = torch.from_numpy(lab[starts[r]:starts[r] + n].astype(np.float32))
            m = torch.from_numpy(rng.random(n).astype(np.float32)) + 0.05
            logits[k, Lpad - 1 - n:Lpad - 1, 1] = l * m
            logits[k, Lpad - 1 - n:Lpad - 1, 0] = (1 - l) * m
        logits[:, 0, :] = 0.3                                    # ties / junk at ignored positions must not matter
        dicts.append(writer.batch_dict(logits, tok, qual, encode.id_rows(ix, rows, np.zeros(rows.size, bool)), lens, Lpad))
    return dicts


def _oracle_text(dicts, recs, opt):
    preds = {}
    for d in dicts:
        preds.update(S.load_predicts_from_batch(d["prediction"].numpy(), d["target"].numpy(), d["seq"].numpy(), d["id"].numpy()))
    fq = [S.FastqRecord(rid.split(" ")[0], rid.partition(" ")[2], s, q) for rid, s, q in recs]
    out = S.chop_records(fq, preds, opt)
    return "".join(r.to_text() for r in out), len(preds), len(out)


@pytest.mark.parametrize("chop_type,ocq", [("all", False), ("terminal", False), ("internal", False), ("all", True)])
def test_chop_output_bit_exact(tmp_path, chop_type, ocq):
    from deepchopper_b200 import writer
    from deepchopper_b200.chop import chop_fastq, params_from_cli
    rng = np.random.default_rng(42)
    lens = np.concatenate([synth.read_lengths(rng, 150, hi=3000), rng.integers(20, 150, 10)])
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    recs[3] = (recs[3][0] + " some description", recs[3][1], recs[3][2])     # description survives only on passthrough
    dicts = _planted_batches(recs, rng)
    # a FASTQ record without prediction is dropped; a truncated prediction passes through
    fq_recs = list(recs) + [("nopred", "ACGT" * 50, "I" * 200)]
    long_id = recs[7][0]
    fq_recs[7] = (long_id, recs[7][1] + "ACGTACGT", recs[7][2] + "IIIIIIII")
    fq_path = tmp_path / "reads.fq"
    fq_path.write_bytes(synth.fastq_text(fq_recs))
    pdir = tmp_path / "predictions"
    for i, d in enumerate(dicts):
        writer.write_batch(str(pdir), 0, i, d)
    params = params_from_cli(chop_type=chop_type, output_chopped=ocq)
    out, npred, nrec = chop_fastq([str(pdir / "0")], str(fq_path), params, output_prefix=str(tmp_path / "out"))
    want, want_pred, want_rec = _oracle_text(dicts, fq_recs, S.ChopOptions(chop_type=chop_type, output_chopped_seqs=ocq))
    assert os.path.basename(out) == f"out.{want_pred}pd.{want_rec}record.chop.fq.gz"   # src/bin/predict.rs:342-352
    got = gzip.open(out, "rb").read().decode()
    assert (npred, nrec) == (want_pred, want_rec)
    assert got == want
    # the compact sidecar (labels only, bit-packed) gives the same output as the .pt dicts
    cdir = tmp_path / "compact"
    for i, d in enumerate(dicts):
        lab = (d["prediction"][..., 1] > d["prediction"][..., 0]).to(torch.uint8)
        keep = (d["target"] != -100)
        lens_b = keep.sum(dim=1).numpy()
        ids = [bytes(d["id"][b, 2:2 + int(d["id"][b, 0])].to(torch.uint8).numpy()).decode("latin1") for b in range(lab.shape[0])]
        writer.write_batch_compact(str(cdir), 0, i, lab, lens_b, lab.shape[1], ids, d["id"][:, 1].numpy())
    out2, npred2, nrec2 = chop_fastq([str(cdir / "0")], str(fq_path), params, output_prefix=str(tmp_path / "out2"))
    assert (npred2, nrec2) == (want_pred, want_rec)
    assert gzip.open(out2, "rb").read().decode() == want
    size_pt = sum(os.path.getsize(pdir / "0" / f) for f in os.listdir(pdir / "0"))
    size_c = sum(os.path.getsize(cdir / "0" / f) for f in os.listdir(cdir / "0"))
    assert size_c * 50 < size_pt
    if not ocq and chop_type == "all":
        assert "|T\n" in got and "|I\n" in got                   # the planted set exercises both chop kinds


def test_cli_predict_layout_and_logits(tmp_path):
    from deepchopper_b200 import cli
    from deepchopper_b200.init_weights import random_state_dict
    rng = np.random.default_rng(5)
    recs = synth.fastq_reads(rng, 30, 150, 900)
    fq = tmp_path / "x.fastq"
    fq.write_bytes(synth.fastq_text(recs))
    out = tmp_path / "predictions"
    cli.main(["predict", str(fq), "-o", str(out), "--random-init", "-b", "12", "--gpus", "1"])
    files = sorted(os.listdir(out / "0"))
    assert files == ["0_0.pt", "0_1.pt", "0_2.pt"]              # {global_rank}_{batch_idx}.pt, callbacks.py:25
    ref = H.make_reference_model(0)
    ref.load_state_dict(random_state_dict(0))
    d = torch.load(out / "0" / "0_1.pt")
    feats = [H.tokenize_read(*recs[i]) for i in range(12, 24)]
    batch = H.collate(feats)
    assert torch.equal(d["seq"], batch["input_ids"])
    assert torch.equal(d["target"], batch["labels"].to(torch.int64))
    assert torch.equal(d["id"], batch["id"].to(torch.int64))
    np.testing.assert_allclose(d["qual"].numpy(), batch["input_quals"].numpy(), rtol=1e-6, atol=1e-9)
    with torch.no_grad():
        want = ref(batch["input_ids"], batch["input_quals"])
    assert (d["prediction"] - want).abs().max() < 2e-2
    # chop on those predictions == oracle chop on the same predictions
    from deepchopper_b200.chop import chop_fastq
    o, npred, nrec = chop_fastq([str(out / "0")], str(fq), output_prefix=str(tmp_path / "y"))
    dicts = [torch.load(out / "0" / f) for f in files]
    want_txt, wp, wr = _oracle_text(dicts, recs, S.ChopOptions())
    assert gzip.open(o, "rb").read().decode() == want_txt and (npred, nrec) == (wp, wr)


def test_host_pipeline_equals_device_pipeline():
    from deepchopper_b200.init_weights import random_state_dict
    from deepchopper_b200.model import DeepChopper
    from deepchopper_b200.predict import DevicePipeline, HostPipeline, plan_batches
    rng = np.random.default_rng(8)
    lens = synth.read_lengths(rng, 300, hi=2500)
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    blob = np.frombuffer(("".join(r[1] for r in recs) + "".join(r[2] for r in recs)).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
    batches = plan_batches(lens, token_budget=64 * 1024)
    dp = DevicePipeline(model)
    dp.upload(blob, off, off + int(lens.sum()), lens, batches)
    hp = HostPipeline(model)
    hp.pack(blob, off, off + int(lens.sum()), lens, batches)
    for di, hi in zip(dp.items, hp.items):
        b = di[0]
        labels_h = np.zeros((b.rows.size, b.Lpad), np.uint8)
        o = hp.run_batch(hi, labels_h)
        _, labels_d, (n_ad, ad, n_keep, keep, act) = dp.run_batch(di)
        torch.cuda.synchronize()
        assert np.array_equal(labels_d.cpu().numpy()[:, :b.Lpad], labels_h)
        assert np.array_equal(o["n_adapter"], n_ad.cpu().numpy()) and np.array_equal(o["action"], act.cpu().numpy())
        assert np.array_equal(o["adapter_iv"], ad.cpu().numpy()) and np.array_equal(o["keep_iv"], keep.cpu().numpy())


def _write_fastq(tmp_path, recs, name="reads.fq"):
    p = tmp_path / name
    p.write_bytes(synth.fastq_text(recs))
    return p


def test_fused_predict_chop_equals_two_step(tmp_path):
    """`predict --chop` (no prediction files) writes the same bytes as `predict --compact` + `chop` on the same batches."""
    from deepchopper_b200 import cli
    rng = np.random.default_rng(17)
    lens = np.concatenate([synth.read_lengths(rng, 400, hi=4000), rng.integers(20, 150, 12)])
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    recs[5] = (recs[5][0] + " a description", recs[5][1], recs[5][2])
    fq = _write_fastq(tmp_path, recs)
    common = ["--random-init", "--bucket", "--token-budget", "65536"]
    cli.main(["predict", str(fq), "-o", str(tmp_path / "pred"), "--compact"] + common)
    cli.main(["chop", str(tmp_path / "pred" / "0"), str(fq), "-o", str(tmp_path / "two"), "-t", "4"])
    cli.main(["predict", str(fq), "--chop", "--chop-output", str(tmp_path / "one"), "-t", "4"] + common)
    two = [f for f in os.listdir(tmp_path) if f.startswith("two.")]
    one = [f for f in os.listdir(tmp_path) if f.startswith("one.")]
    assert len(two) == 1 and len(one) == 1 and two[0][3:] == one[0][3:]      # same {n}pd.{m}record counts
    a = gzip.open(tmp_path / two[0], "rb").read()
    b = gzip.open(tmp_path / one[0], "rb").read()
    assert a == b and len(a) > 0


def test_grouped_reference_batches_equal_one_launch_per_batch(tmp_path):
    """The reference's FASTQ-order batches packed into launches (predict.group_batches + dcb200_encode_batch_rows): a row's
    tokens / quality equal those of its own batch's collation, and -- with one long-convolution kernel for both runs --
    `predict --chop -b 16` writes the same bytes whether a launch holds one batch or forty (causal model, independent
    rows: the right filler and the neighbours cannot reach a row's first Lpad columns)."""
    from deepchopper_b200 import _native, cli, ops  # noqa: F401
    from deepchopper_b200.encode import encode_batch_device
    from deepchopper_b200.predict import Batch, group_batches
    rng = np.random.default_rng(29)
    lens = np.concatenate([synth.read_lengths(rng, 500, hi=5000), rng.integers(20, 150, 20)])
    rng.shuffle(lens)
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    # (a) encode_rows == encode per batch
    blob = np.frombuffer(("".join(r[1] for r in recs) + "".join(r[2] for r in recs)).encode(), dtype=np.uint8)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    dev = torch.device("cuda", 0)
    d_blob = torch.from_numpy(blob.copy()).to(dev)
    batches = []
    for i in range(0, lens.size, 16):
        rows = np.arange(i, min(i + 16, lens.size))
        lp = int(lens[rows].max()) + 1
        batches.append(Batch(rows, lp, (lp + 127) // 128 * 128))
    launches = group_batches(batches, 256 * 1024)
    assert len(launches) < len(batches) / 4
    tot = int(lens.sum())
    for g in launches[:3]:
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)  # noqa: E731
        tok, qual = torch.ops.dcb200.encode_rows(d_blob, t(off[g.rows], np.int64), t(off[g.rows] + tot, np.int64),
                                                 t(lens[g.rows], np.int32), t(g.lpad, np.int32), int(g.Lpad), int(g.Lrow))
        for pos, r0, r1, b in g.members:
            tk, ql = encode_batch_device(d_blob, t(off[b.rows], np.int64), t(off[b.rows] + tot, np.int64),
                                         t(lens[b.rows], np.int32), b.Lpad, None, b.Lrow)
            assert torch.equal(tok[r0:r1, :b.Lrow], tk) and torch.equal(qual[r0:r1, :b.Lrow], ql)
            assert (tok[r0:r1, b.Lrow:] == 4).all() and (qual[r0:r1, b.Lrow:] == 0).all()
    # (a') the host-buffer C entry for such a launch (dcb200_predict_batch_host_rows) == the three device ops
    from deepchopper_b200.init_weights import random_state_dict
    from deepchopper_b200.model import DeepChopper
    from deepchopper_b200.predict import HostPipeline
    from deepchopper_b200.smooth import smooth_chop_device
    model = DeepChopper.from_state_dict(random_state_dict(0), device=0)
    hp = HostPipeline(model)
    items = []
    for g in launches[:2]:
        ln = lens[g.rows].astype(np.int32)
        so = (np.cumsum(ln) - ln).astype(np.int64)
        buf = np.concatenate([blob[off[r]:off[r] + lens[r]] for r in g.rows] + [blob[tot + off[r]:tot + off[r] + lens[r]] for r in g.rows])
        items.append((g, buf, so, so + int(ln.sum()), ln))
    hp.pack_items(items, pin=False)
    for g, hi in zip(launches[:2], hp.items):
        labels_h = np.zeros((g.rows.size, g.Lpad), np.uint8)
        o = hp.run_batch(hi, labels_h)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)  # noqa: E731
        ln_d, lp_d = t(lens[g.rows], np.int32), t(g.lpad, np.int32)
        tok, qual = torch.ops.dcb200.encode_rows(d_blob, t(off[g.rows], np.int64), t(off[g.rows] + tot, np.int64), ln_d, lp_d,
                                                 int(g.Lpad), int(g.Lrow))
        _, labels_d = model.forward_tokens(tok, qual, False, True)
        st = torch.arange(g.rows.size, dtype=torch.int64, device=dev) * g.Lrow + (lp_d.to(torch.int64) - 1) - ln_d.to(torch.int64)
        n_ad, ad, n_keep, keep, act = smooth_chop_device(labels_d.view(-1), st, ln_d)
        torch.cuda.synchronize()
        assert np.array_equal(labels_d.cpu().numpy()[:, :g.Lpad], labels_h)
        assert np.array_equal(o["n_adapter"], n_ad.cpu().numpy()) and np.array_equal(o["action"], act.cpu().numpy())
        assert np.array_equal(o["adapter_iv"], ad.cpu().numpy()) and np.array_equal(o["keep_iv"], keep.cpu().numpy())
    # (b) the fused route, one batch per launch vs grouped launches, Toeplitz kernel for both
    fq = _write_fastq(tmp_path, recs)
    ctx = _native.torch_context(dev)
    old = ctx.get_option("fft_min_len")
    ctx.set_option("fft_min_len", 1 << 30)
    try:
        common = ["predict", str(fq), "--chop", "--random-init", "-b", "16", "-t", "4"]
        cli.main(common + ["--chop-output", str(tmp_path / "single"), "--token-budget", "1"])
        cli.main(common + ["--chop-output", str(tmp_path / "grouped"), "--token-budget", "262144"])
    finally:
        ctx.set_option("fft_min_len", old)
    a = [f for f in os.listdir(tmp_path) if f.startswith("single.")]
    b = [f for f in os.listdir(tmp_path) if f.startswith("grouped.")]
    assert len(a) == 1 and len(b) == 1 and a[0][len("single"):] == b[0][len("grouped"):]
    ta, tb = gzip.open(tmp_path / a[0], "rb").read(), gzip.open(tmp_path / b[0], "rb").read()
    assert ta == tb and len(ta) > 0


def test_truncated_read_passes_through(tmp_path):
    """A read of >= 32768 bases is cut to the model's window and flagged (tokenizer.py:154-163); `chop` then refuses to
    cut it (prediction length != FASTQ quality length, src/bin/predict.rs:160-164) and emits the record verbatim.  The
    `.pt` batch equals the oracle's tokenisation, the logits the fp32 oracle's, both routes write the oracle's bytes."""
    from deepchopper_b200 import cli
    from deepchopper_b200.chop import chop_fastq
    from deepchopper_b200.init_weights import random_state_dict
    rng = np.random.default_rng(23)
    lens = np.array([700, 33000, 32768, 32767, 900])
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    fq = _write_fastq(tmp_path, recs)
    out = tmp_path / "predictions"
    cli.main(["predict", str(fq), "-o", str(out), "--random-init", "-b", "5", "--gpus", "1"])
    d = torch.load(out / "0" / "0_0.pt")
    feats = [H.tokenize_read(*r) for r in recs]
    batch = H.collate(feats)
    assert d["seq"].shape == (5, 32768) and torch.equal(d["seq"], batch["input_ids"])
    assert torch.equal(d["id"], batch["id"].to(torch.int64)) and d["id"][:, 1].tolist() == [0, 1, 1, 0, 0]
    assert torch.equal(d["target"], batch["labels"].to(torch.int64))
    np.testing.assert_allclose(d["qual"].numpy(), batch["input_quals"].numpy(), rtol=1e-6, atol=1e-9)
    ref = H.make_reference_model(0)
    ref.load_state_dict(random_state_dict(0))
    with torch.no_grad():
        want = ref(batch["input_ids"], batch["input_quals"])
    assert (d["prediction"] - want).abs().max() < 5e-2
    o, npred, nrec = chop_fastq([str(out / "0")], str(fq), output_prefix=str(tmp_path / "y"))
    want_txt, wp, wr = _oracle_text([d], recs, S.ChopOptions())
    got = gzip.open(o, "rb").read().decode()
    assert got == want_txt and (npred, nrec) == (wp, wr)
    for k in (1, 2):   # the two truncated reads come out verbatim, whatever the labels say
        rid, s, q = recs[k]
        assert f"@{rid}\n{s}\n+\n{q}\n" in got
    cli.main(["predict", str(fq), "--chop", "--chop-output", str(tmp_path / "z"), "--random-init", "-b", "5"])
    z = [f for f in os.listdir(tmp_path) if f.startswith("z.")]
    assert gzip.open(tmp_path / z[0], "rb").read().decode() == want_txt


def test_predict_cli_and_loaders_on_gpu(tmp_path):
    """The PyO3-named entry points: predict_cli (src/cli.rs:57-165) writes `.chop.fq.bgz` with the same records as the
    chop driver; a prediction without a FASTQ record is an error there; Predict objects from load_predicts_from_batch_pts
    smooth on the GPU like the oracle."""
    import deepchopper_b200 as D
    from deepchopper_b200 import writer
    from deepchopper_b200.chop import chop_fastq
    rng = np.random.default_rng(31)
    lens = synth.read_lengths(rng, 60, hi=2500)
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    dicts = _planted_batches(recs, rng)
    fq = _write_fastq(tmp_path, recs)
    pdir = tmp_path / "predictions"
    for i, d in enumerate(dicts):
        writer.write_batch(str(pdir), 0, i, d)
    D.predict_cli([str(pdir / "0")], str(fq), output_prefix=str(tmp_path / "pc"), threads=2)
    o, npred, nrec = chop_fastq([str(pdir / "0")], str(fq), output_prefix=str(tmp_path / "cf"))
    pc = [f for f in os.listdir(tmp_path) if f.startswith("pc.")]
    assert pc == [f"pc.{npred}pd.{nrec}record.chop.fq.bgz"]                 # src/cli.rs:143-157
    assert gzip.open(tmp_path / pc[0], "rb").read() == gzip.open(o, "rb").read()
    fq2 = _write_fastq(tmp_path, recs[:-1], "fewer.fq")
    with pytest.raises(KeyError, match="id not found"):
        D.predict_cli([str(pdir / "0")], str(fq2), output_prefix=str(tmp_path / "bad"))
    preds = D.load_predicts_from_batch_pts(str(pdir))
    assert len(preds) == len(recs)
    for rid in list(preds)[:10]:
        p = preds[rid]
        assert p.smooth_and_select_intervals(21, 13, 20) == [tuple(x) for x in S.smooth_label_region(p.prediction, 21, 13, 20)]
        assert p.smooth_label(21) == S.majority_voting(p.prediction, 21)


def test_torch_ops_direct():
    """torch.ops.dcb200.* called directly: same results as the host mirror's functions; approved_interval_number = 0 is
    legal (the reference then treats every read as having no interval -> passthrough)."""
    import deepchopper_b200.ops as ops  # noqa: F401
    from deepchopper_b200._native import ChopParams
    from oracle import cref
    rng = np.random.default_rng(2)
    lens = synth.read_lengths(rng, 50, hi=1500)
    lab, starts, ln = synth.planted_labels(rng, lens)
    dev = torch.device("cuda", 0)
    args = (torch.from_numpy(lab).to(dev), torch.from_numpy(starts).to(dev), torch.from_numpy(ln).to(dev),
            torch.empty(0, dtype=torch.int32, device=dev))
    p = ChopParams.default()
    n_ad, ad, n_keep, keep, act = torch.ops.dcb200.smooth_chop(*args, ops.params_list(p))
    chk = cref.load().smooth_chop(lab, starts, ln)
    assert np.array_equal(n_ad.cpu().numpy(), chk["n_adapter"]) and np.array_equal(act.cpu().numpy(), chk["action"])
    p0 = ChopParams.default(approved_interval_number=0)
    n_ad0, ad0, n_keep0, keep0, act0 = torch.ops.dcb200.smooth_chop(*args, ops.params_list(p0))
    assert ad0.shape == (50, 0, 2) and int(n_ad0.sum()) == 0 and int(act0.sum()) == 0
    # logits form: fp32 [N, 2] -> argmax fused
    lg = torch.zeros((lab.size, 2), device=dev)
    lg[:, 1] = torch.from_numpy(lab.astype(np.float32)).to(dev) - 0.5
    out2 = torch.ops.dcb200.smooth_chop(lg, *args[1:], ops.params_list(p))
    assert torch.equal(out2[0], n_ad) and torch.equal(out2[4], act) and torch.equal(out2[1], ad)


@pytest.mark.parametrize("ocq", [False, True])
def test_streaming_chop_equals_whole_file(tmp_path, ocq):
    """`chop --chunk-mb` (FASTQ streamed in pieces, decisions matched by id, BGZF parts appended) writes the same
    decompressed bytes and counts as the whole-file driver, for .pt dicts and for compact sidecars; covers a truncated
    prediction, a read without prediction and a description line."""
    from deepchopper_b200 import writer
    from deepchopper_b200.chop import chop_fastq, params_from_cli
    rng = np.random.default_rng(77)
    lens = np.concatenate([synth.read_lengths(rng, 200, hi=2500), rng.integers(20, 150, 10)])
    recs = synth.fastq_reads(rng, lens.size, lengths=lens)
    recs[3] = (recs[3][0] + " some description", recs[3][1], recs[3][2])
    dicts = _planted_batches(recs, rng)
    fq_recs = list(recs) + [("nopred", "ACGT" * 50, "I" * 200)]
    fq_recs[7] = (recs[7][0], recs[7][1] + "ACGTACGT", recs[7][2] + "IIIIIIII")        # prediction shorter than the record
    fq_path = tmp_path / "reads.fq"
    fq_path.write_bytes(synth.fastq_text(fq_recs))
    params = params_from_cli(output_chopped=ocq)
    pdir, cdir = tmp_path / "predictions", tmp_path / "compact"
    for i, d in enumerate(dicts):
        writer.write_batch(str(pdir), 0, i, d)
        lab = (d["prediction"][..., 1] > d["prediction"][..., 0]).to(torch.uint8)
        lens_b = (d["target"] != -100).sum(dim=1).numpy()
        ids = [bytes(d["id"][b, 2:2 + int(d["id"][b, 0])].to(torch.uint8).numpy()).decode("latin1") for b in range(lab.shape[0])]
        writer.write_batch_compact(str(cdir), 0, i, lab, lens_b, lab.shape[1], ids, d["id"][:, 1].numpy())
    for name, src in (("pt", pdir), ("compact", cdir)):
        whole, np1, nr1 = chop_fastq([str(src / "0")], str(fq_path), params, output_prefix=str(tmp_path / f"w_{name}"))
        part, np2, nr2 = chop_fastq([str(src / "0")], str(fq_path), params, output_prefix=str(tmp_path / f"s_{name}"),
                                    chunk_bytes=30000, threads=3)
        assert (np1, nr1) == (np2, nr2)
        a, b = gzip.open(whole, "rb").read(), gzip.open(part, "rb").read()
        assert a == b and len(a) > 0
        assert os.path.basename(part) == f"s_{name}.{np1}pd.{nr1}record.chop.fq.gz"
