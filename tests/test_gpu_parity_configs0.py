"""-m gpu: BASELINE configs[0] (the reference's own CPU-runnable case) end to end -- synthetic dRNA reads of 0.5-2 kb,
FASTQ order, batch 16 -- with EACH SIDE USING ITS OWN LABELS: the GPU path's labels -> GPU smoothing vs the fp32
oracle's labels -> the C oracle's smoothing.  Labels may differ at near-tie positions (bf16 operands); the majority
vote absorbs most isolated flips, so what is reported and bounded is the fraction of READS whose smoothed adapter
intervals and chop decision come out identical.  tools/parity_configs0.py runs the full 1000 reads and prints the same
figures (committed under profiles/)."""
import numpy as np
import pytest
import torch

from deepchopper_b200 import synth
from oracle import cref, hyena_ref as H

pytestmark = pytest.mark.gpu


def compare_configs0(n_reads: int, seed: int = 20261018, batch: int = 16):
    from deepchopper_b200.init_weights import random_state_dict
    from deepchopper_b200.model import DeepChopper
    from deepchopper_b200.smooth import smooth_chop_device
    sd = random_state_dict(0)
    ref = H.make_reference_model(0)
    ref.load_state_dict(sd)
    gpu = DeepChopper.from_state_dict(sd, device=0)
    rng = np.random.default_rng(seed)
    recs = synth.fastq_reads(rng, n_reads, 500, 2000)
    c = cref.load()
    same_iv = same_act = 0
    flips = positions = 0
    max_err = 0.0
    for i in range(0, n_reads, batch):
        feats = [H.tokenize_read(*r) for r in recs[i:i + batch]]
        b = H.collate(feats)
        with torch.no_grad():
            want = ref(b["input_ids"], b["input_quals"])
        got = gpu(b["input_ids"].cuda(), b["input_quals"].cuda())
        L = want.shape[1]
        ln = np.array([len(f["input_ids"]) - 1 for f in feats], dtype=np.int32)
        starts = (np.arange(len(feats)) * L + (L - 1) - ln).astype(np.int64)
        lab_ref = (want[..., 1] > want[..., 0]).to(torch.int8).numpy()
        lab_gpu = (got[..., 1] > got[..., 0]).to(torch.int8)
        r_ref = c.smooth_chop(lab_ref.reshape(-1), starts, ln)
        n_ad, ad, n_keep, keep, act = smooth_chop_device(lab_gpu.reshape(-1).contiguous(), torch.from_numpy(starts).cuda(),
                                                         torch.from_numpy(ln).cuda())
        n_ad, ad, act = n_ad.cpu().numpy(), ad.cpu().numpy(), act.cpu().numpy()
        for k in range(len(feats)):
            eq = n_ad[k] == r_ref["n_adapter"][k] and np.array_equal(ad[k, :n_ad[k]], r_ref["adapter_iv"][k, :n_ad[k]])
            same_iv += int(eq)
            same_act += int(eq and act[k] == r_ref["action"][k])
            sl = slice(L - 1 - ln[k], L - 1)
            flips += int((lab_gpu[k, sl].cpu().numpy() != lab_ref[k, sl]).sum())
            positions += int(ln[k])
        max_err = max(max_err, float((got.cpu() - want).abs().max()))
    return {"reads": n_reads, "reads_with_identical_smoothed_intervals": same_iv / n_reads,
            "reads_with_identical_intervals_and_action": same_act / n_reads, "label_flip_fraction": flips / positions,
            "max_abs_logit_error": max_err}


def test_configs0_own_labels_interval_agreement():
    res = compare_configs0(160)
    print(res)
    assert res["max_abs_logit_error"] < 5e-2
    assert res["label_flip_fraction"] < 5e-3       # configs[0] batches: measured 2-3e-4
    # Random-init weights put ~10 % of all positions within 2e-2 of a tie, so raw labels are noise-like and a single
    # flipped base next to a window tie moves an interval edge.  Measured: 96.9 % of 160 reads, 95.2 % of 1000 reads
    # (profiles/r02_summary.md) come out with identical intervals AND chop decision; the bound leaves room for the seed.
    assert res["reads_with_identical_smoothed_intervals"] >= 0.85
    assert res["reads_with_identical_intervals_and_action"] >= 0.85
