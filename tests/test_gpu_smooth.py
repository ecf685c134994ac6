"""-m gpu parity tests of the smoothing / interval / chop-coordinate kernel against the oracle,
through the C ABI (dcb200_smooth_chop*, dcb200_majority_voting*)."""
import json
import os

import numpy as np
import pytest

from oracle import smooth_ref as S
from deepchopper_b200 import synth
from helpers_kats import MV_KATS, REGION_KATS, _random_labels

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sm():
    from deepchopper_b200 import smooth
    return smooth


@pytest.fixture(params=[1, 2], ids=["warp_kernel", "tile_kernel"], autouse=True)
def smooth_kernel(request):
    """Every test of this module runs once per int8-label kernel (ctx option smooth_warp_kernel: 1 = warp per read,
    2 = tile; the default picks by launch size and would leave the tile kernel out of the small cases)."""
    import torch
    from deepchopper_b200 import _native
    ctxs = [_native.default_context(), _native.torch_context(torch.device("cuda", 0))]
    for c in ctxs:
        c.set_option("smooth_warp_kernel", request.param)
    yield request.param
    for c in ctxs:
        c.set_option("smooth_warp_kernel", 0)


@pytest.mark.parametrize("labels,window,expected", MV_KATS)
def test_majority_voting_kats(sm, labels, window, expected):
    assert sm.majority_voting(labels, window) == expected


@pytest.mark.parametrize("labels,expected", REGION_KATS)
def test_get_label_region_kats(sm, labels, expected):
    assert sm.get_label_region(labels) == expected


def test_predict_class_surface(sm):
    lab = [0] * 50 + [1] * 40 + [0] * 100 + [1] * 5 + [0] * 30
    p = sm.Predict(lab, "A" * len(lab), "x", False)
    o = S.Predict(lab, "A" * len(lab), "x")
    assert p.prediction_region() == o.prediction_region()
    assert p.smooth_prediction(21) == o.smooth_prediction(21)
    assert p.smooth_label(21) == o.smooth_label(21)
    assert p.smooth_and_select_intervals(21, 13, 20) == o.smooth_and_select_intervals(21, 13, 20)
    assert p.seq_len() == len(lab)


def test_fixture72_bit_exact(sm, dcref):
    z = np.load(os.path.join(GOLD, "smooth_fixture.npz"))
    meta = json.load(open(os.path.join(GOLD, "smooth_fixture.json")))
    offs = z["offsets"]
    lens = np.diff(offs).astype(np.int32)
    res = sm.smooth_chop_host(z["labels"], offs[:-1], lens)
    ref = dcref.smooth_chop(z["labels"], offs[:-1], lens)
    for r in range(lens.size):
        want = [tuple(x) for x in meta["intervals"][r]]
        if lens[r] >= 150:
            assert res.adapters(r) == want, meta["ids"][r]
    for k in ("n_adapter", "n_keep", "action"):
        assert np.array_equal(getattr(res, k), ref[k]), k
    for r in range(lens.size):
        assert np.array_equal(res.adapter_iv[r, :res.n_adapter[r]], ref["adapter_iv"][r, :ref["n_adapter"][r]])
        assert np.array_equal(res.keep_iv[r, :res.n_keep[r]], ref["keep_iv"][r, :ref["n_keep"][r]])


def _compare(res, ref):
    for k in ("n_adapter", "n_keep", "action"):
        bad = np.nonzero(getattr(res, k) != ref[k])[0]
        assert bad.size == 0, (k, bad[:10], getattr(res, k)[bad[:10]], ref[k][bad[:10]])
    ap = res.adapter_iv.shape[1]
    m = np.arange(ap)[None, :] < res.n_adapter[:, None]
    assert np.array_equal(res.adapter_iv[m], ref["adapter_iv"][m])
    m = np.arange(ap + 1)[None, :] < res.n_keep[:, None]
    assert np.array_equal(res.keep_iv[m], ref["keep_iv"][m])


PARAM_SETS = [
    dict(window=21, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=150, chop_type=2, ocq=0),
    dict(window=21, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=150, chop_type=0, ocq=0),
    dict(window=21, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=150, chop_type=1, ocq=0),
    dict(window=21, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=150, chop_type=2, ocq=1),
    dict(window=1, min_interval=0, approved=64, max_process=64, min_after_chop=0, min_read_len=0, chop_type=2, ocq=0),
    dict(window=2, min_interval=2, approved=1, max_process=1, min_after_chop=50, min_read_len=20, chop_type=2, ocq=0),
    dict(window=11, min_interval=5, approved=3, max_process=2, min_after_chop=10, min_read_len=0, chop_type=0, ocq=0),
    dict(window=51, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=150, chop_type=2, ocq=0),
    dict(window=63, min_interval=1, approved=20, max_process=20, min_after_chop=1, min_read_len=0, chop_type=2, ocq=0),
    dict(window=101, min_interval=13, approved=20, max_process=4, min_after_chop=20, min_read_len=0, chop_type=2, ocq=0),
]


@pytest.mark.parametrize("ps", PARAM_SETS)
def test_random_ragged_reads_bit_exact(sm, dcref, ps):
    from deepchopper_b200 import ChopParams
    rng = np.random.default_rng(11)
    lens = np.concatenate([rng.integers(0, 70, 300), rng.integers(100, 3000, 300), [0, 1, 2, 31, 32, 33, 1023, 1024, 1025, 2047, 2048, 2049, 8192]])
    labs = [_random_labels(rng, int(n)) if n else np.zeros(0, np.int8) for n in lens]
    # unaligned starts: put a random gap in front of every read
    gaps = rng.integers(0, 37, lens.size)
    starts = np.cumsum(gaps + np.concatenate([[0], lens[:-1]])).astype(np.int64)
    buf = (rng.random(int(starts[-1] + lens[-1] + 5)) < 0.5).astype(np.int8)  # junk between reads must not leak in
    for s, l in zip(starts, labs):
        buf[s:s + l.size] = l
    qual_lens = lens.astype(np.int32).copy()
    qual_lens[::13] += 3
    p = ChopParams.default(smooth_window_size=ps["window"], min_interval_size=ps["min_interval"],
                           approved_interval_number=ps["approved"], max_process_intervals=ps["max_process"],
                           min_read_length_after_chop=ps["min_after_chop"], min_read_length=ps["min_read_len"],
                           chop_type=ps["chop_type"], output_chopped_seqs=ps["ocq"])
    res = sm.smooth_chop_host(buf, starts, lens.astype(np.int32), p, qual_lens)
    ref = dcref.smooth_chop(buf, starts, lens.astype(np.int32), qual_lens, **ps)
    _compare(res, ref)
    sm_gpu = sm.majority_voting_host(buf, starts, lens.astype(np.int32), ps["window"])
    for s, l in zip(starts[::7], labs[::7]):
        assert np.array_equal(sm_gpu[s:s + l.size], dcref.majority_voting(l, ps["window"]))


def test_tile_kernel_sub_batches_and_warp_kernel(sm, dcref):
    """The int8 path is the tile kernel (thread per 32-base word, 64 reads per CTA): tiles of long reads are split into
    sub-batches of whole reads, slotless reads (empty / below min_read_length) sit between members, and the last read of
    a sub-batch may reach beyond its window.  Bit-exact against the C oracle and against the warp-per-read kernel
    (ctx option smooth_warp_kernel)."""
    from deepchopper_b200 import ChopParams
    from deepchopper_b200._native import default_context
    rng = np.random.default_rng(77)
    lens = np.concatenate([rng.integers(20000, 32769, 70), rng.integers(0, 200, 40), rng.integers(3000, 9000, 150),
                           [32768, 0, 32768, 149, 32767, 150, 32736, 1, 32737]])
    rng.shuffle(lens)
    lab, starts, ln = synth.planted_labels_fast(rng, lens)
    lab = lab.copy()
    lab[rng.random(lab.size) < 0.3] ^= 1          # many short runs as well: every word has starts and ends
    big = rng.integers(0, lens.size, 12)
    for r in big:                                  # and some reads that are one long run (the backward search for a start)
        lab[starts[r]:starts[r] + ln[r]] = 1
    for ps in (dict(), dict(smooth_window_size=1, min_interval_size=1, approved_interval_number=64, max_process_intervals=64,
                            min_read_length=0), dict(smooth_window_size=41, min_read_length=0)):
        p = ChopParams.default(**ps)
        ctx = default_context()
        ctx.set_option("smooth_warp_kernel", 2)
        res = sm.smooth_chop_host(lab, starts, ln, p)
        kw = dict(window=p.smooth_window_size, min_interval=p.min_interval_size, approved=p.approved_interval_number,
                  max_process=p.max_process_intervals, min_after_chop=p.min_read_length_after_chop,
                  min_read_len=p.min_read_length, chop_type=p.chop_type, ocq=p.output_chopped_seqs)
        _compare(res, dcref.smooth_chop(lab, starts, ln, None, **kw))
        ctx.set_option("smooth_warp_kernel", 1)
        res2 = sm.smooth_chop_host(lab, starts, ln, p)
        for k in ("n_adapter", "adapter_iv", "n_keep", "keep_iv", "action"):
            a, b = getattr(res, k), getattr(res2, k)
            if k == "adapter_iv":
                m = np.arange(a.shape[1])[None, :] < res.n_adapter[:, None]
                assert np.array_equal(a[m], b[m])
            elif k == "keep_iv":
                m = np.arange(a.shape[1])[None, :] < res.n_keep[:, None]
                assert np.array_equal(a[m], b[m])
            else:
                assert np.array_equal(a, b), k


def test_reads_longer_than_the_model_window(sm, dcref):
    """The smoothing-only entry points take reads of any length (the tile kernel's shared arrays hold up to 32768 bases per
    read: longer ones go through the warp-per-read kernel in a second launch behind it)."""
    from deepchopper_b200._native import default_context
    rng = np.random.default_rng(31)
    lens = np.concatenate([rng.integers(200, 3000, 300), [32768, 32769, 40000, 100001, 65536], rng.integers(200, 3000, 100)])
    lab, starts, ln = synth.planted_labels_fast(rng, lens)
    ctx = default_context()
    ctx.set_option("smooth_warp_kernel", 2)
    res = sm.smooth_chop_host(lab, starts, ln)
    _compare(res, dcref.smooth_chop(lab, starts, ln))
    assert res.n_adapter[300:305].sum() > 0


def test_label_domain_other_values_count_as_zero(sm):
    """include/dcb200.h: a label is an adapter base iff it equals 1.  Bytes such as -100 (the ignore index), 2, 3, -1 or 127
    take the exact packing path (the 0/1 fast path is checked per 32 labels, not assumed) and must behave like 0."""
    rng = np.random.default_rng(9)
    lens = synth.read_lengths(rng, 400)
    lab, starts, ln = synth.planted_labels_fast(rng, lens)
    weird = lab.copy()
    zero = np.nonzero(lab == 0)[0]
    pick = rng.choice(zero, zero.size // 3, replace=False)
    weird[pick] = rng.choice(np.array([-100, 2, 3, -1, 127, -128, 0x11], dtype=np.int8), pick.size)
    a = sm.smooth_chop_host(lab, starts, ln)
    b = sm.smooth_chop_host(weird, starts, ln)
    for k in ("n_adapter", "adapter_iv", "n_keep", "keep_iv", "action"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(sm.majority_voting_host(weird, starts, ln, 21), sm.majority_voting_host(lab, starts, ln, 21))


def test_logits_variant_matches_label_variant(sm):
    import torch
    from deepchopper_b200 import ChopParams
    rng = np.random.default_rng(5)
    lens = synth.read_lengths(rng, 500)
    lab, starts, lens32 = synth.planted_labels(rng, lens)
    margin = rng.random(lab.size).astype(np.float32) + 0.01
    logits = np.stack([np.where(lab == 1, 0, margin), np.where(lab == 1, margin, 0)], 1).astype(np.float32)
    logits[::97] = 0.5   # exact ties -> class 0 (argmax first index)
    lab2 = (logits[:, 1] > logits[:, 0]).astype(np.int8)
    a = sm.smooth_chop_device(torch.from_numpy(lab2).cuda(), torch.from_numpy(starts).cuda(), torch.from_numpy(lens32).cuda())
    b = sm.smooth_chop_device(torch.from_numpy(logits).cuda(), torch.from_numpy(starts).cuda(), torch.from_numpy(lens32).cuda(), logits=True)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_full_size_properties(sm, dcref):
    """BASELINE config-5 sized slice (2M reads here; 10M in bench) through size-independent properties:
    idempotent action histogram vs a 20k-read oracle sample, interval sanity on everything."""
    import torch
    rng = np.random.default_rng(2026)
    lens = synth.read_lengths(rng, 2_000_000)
    lab, starts, lens32 = synth.planted_labels_fast(rng, lens)
    n_ad, ad, n_keep, keep, act = [t.cpu().numpy() for t in sm.smooth_chop_device(
        torch.from_numpy(lab).cuda(), torch.from_numpy(starts).cuda(), torch.from_numpy(lens32).cuda())]
    # properties: intervals sorted, disjoint, inside the read, long enough; kept pieces complement adapters
    m = np.arange(20)[None, :] < n_ad[:, None]
    s, e = ad[..., 0], ad[..., 1]
    assert (e[m] - s[m] >= 13).all() and (s[m] >= 1).all()
    assert (e[m] <= np.broadcast_to(lens32[:, None], e.shape)[m]).all()
    m2 = m[:, 1:] & m[:, :-1]
    assert (s[:, 1:][m2] > e[:, :-1][m2]).all()
    assert ((act == 0) | (n_keep > 0) | (n_ad > 0)).all() and (act <= 4).all()
    # exact agreement with the oracle on a sample
    idx = rng.choice(lens.size, 20000, replace=False)
    ref = dcref.smooth_chop(lab, starts[idx], lens32[idx])
    assert np.array_equal(n_ad[idx], ref["n_adapter"]) and np.array_equal(act[idx], ref["action"])
    assert np.array_equal(n_keep[idx], ref["n_keep"])
    mm = np.arange(20)[None, :] < ref["n_adapter"][:, None]
    assert np.array_equal(ad[idx][mm], ref["adapter_iv"][mm])
    mk = np.arange(21)[None, :] < ref["n_keep"][:, None]
    assert np.array_equal(keep[idx][mk], ref["keep_iv"][mk])


def test_full_size_properties_chunking_and_order():
    """Size-independent properties at a size the Python oracle cannot follow (300 k reads, 360 M labels): the result of a
    read does not depend on how the reads are split over launches nor on their order in the buffer; the C oracle checks
    a sample bit for bit."""
    import torch
    from deepchopper_b200.smooth import smooth_chop_device
    from oracle import cref
    rng = np.random.default_rng(123)
    R = 300_000
    lens = synth.read_lengths(rng, R)
    lab, starts, ln = synth.planted_labels_fast(rng, lens)
    dev = torch.device("cuda", 0)
    d_lab, d_st, d_ln = torch.from_numpy(lab).to(dev), torch.from_numpy(starts).to(dev), torch.from_numpy(ln).to(dev)
    whole = [t.cpu().numpy() for t in smooth_chop_device(d_lab, d_st, d_ln)]
    # (a) ten launches of 30 k reads each == one launch
    parts = [[t.cpu().numpy() for t in smooth_chop_device(d_lab, d_st[i:i + 30_000].contiguous(), d_ln[i:i + 30_000].contiguous())]
             for i in range(0, R, 30_000)]
    for k in range(5):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k])
    # (b) reads visited in a random order == the same rows permuted
    perm = rng.permutation(R)
    shuffled = [t.cpu().numpy() for t in smooth_chop_device(d_lab, d_st[torch.from_numpy(perm).to(dev)].contiguous(),
                                                            d_ln[torch.from_numpy(perm).to(dev)].contiguous())]
    for k in range(5):
        assert np.array_equal(shuffled[k], whole[k][perm])
    # (c) a sample against the C oracle
    idx = np.sort(rng.choice(R, 5000, replace=False))
    sub_lab = np.concatenate([lab[starts[i]:starts[i] + ln[i]] for i in idx])
    sub_st = np.concatenate([[0], np.cumsum(ln[idx])[:-1]]).astype(np.int64)
    chk = cref.load().smooth_chop(sub_lab, sub_st, ln[idx])
    assert np.array_equal(whole[0][idx], chk["n_adapter"]) and np.array_equal(whole[4][idx], chk["action"])
    assert np.array_equal(whole[2][idx], chk["n_keep"])
    assert int(whole[0].sum()) > R // 2       # the planted runs are found
