"""Pins the smoothing/chop oracle (oracle/smooth_ref.{py,c}) to the reference's own Rust unit-test
vectors and to the 72 real prediction reads of the reference's fixtures (tests/golden)."""
import json
import os

import numpy as np
import pytest

from oracle import smooth_ref as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")

from helpers_kats import MV_KATS, REGION_KATS, _random_labels  # noqa: E402


@pytest.mark.parametrize("labels,window,expected", MV_KATS)
def test_majority_voting_kats(labels, window, expected, dcref):
    assert S.majority_voting(labels, window) == expected
    assert dcref.majority_voting(labels, window).tolist() == expected


@pytest.mark.parametrize("labels,expected", REGION_KATS)
def test_get_label_region_kats(labels, expected, dcref):
    assert S.get_label_region(labels) == expected
    assert dcref.get_label_region(labels) == expected


def test_summary_predict_kat():
    # src/utils.rs:742-751
    p, l = S.summary_predict([[0, 0, 1], [1, 1, 1]], [[0, -100, 1], [-100, 1, -100]], -100)
    assert p == [[0, 1], [1]] and l == [[0, 1], [1]]


def test_remove_intervals_and_keep_left_kat():
    # src/output/split.rs:326-345
    seq = "abcdefghijklmnopqrstuvwxyz"
    assert S.remove_intervals_and_keep_left(seq, [(1, 5), (10, 15), (20, 25)])[0] == ["a", "fghij", "pqrst"]
    assert S.remove_intervals_and_keep_left(seq, [(5, 10), (15, 20)])[0] == ["abcde", "klmno", "uvwxy"]
    assert S.remove_intervals_and_keep_left(seq, [])[0] == [seq]


def test_generate_unmaped_intervals_kat():
    # src/output/split.rs:347-353
    assert S.generate_unmaped_intervals([(8100, 8123)], 32768) == [(0, 8100), (8123, 32767)]


def test_id_list2seq_kats():
    # src/smooth/utils.rs:139-165
    assert S.id_list2seq([7, 8, 9, 10, 11]) == "ACGTN"
    assert S.id_list2seq([0, 1, 6, 7, 8, 9, 10, 11]) == "NNNACGTN"


# ---- SURVEY.md Appendix B (an independent derivation from the same fixtures) -----------------
APPENDIX_B = """
03078328 750 [(681,750)] · 02c7fc86 399 [(350,399)] · 10dcea02 1315 [(1256,1315)] · 110d30bb 989 [(924,989)] · 11ca9f84 552 [(4,34),(486,552)] · 00b8e566 842 [(766,842)]
281f4601 733 [(682,733)] · 10e2e940 789 [(733,789)] · 1f422fe8 840 [(778,840)] · 2737297b 3274 [(2260,2325),(3195,3274)] · 1c4186e8 1293 [(1217,1293)] · 24c74bb7 1502 [(1445,1502)]
252a78c3 745 [(687,745)] · 3600765b 351 [] · f2c8bcef 1411 [] · b67bbaf7 248 [(183,248)] · 45438708 878 [(803,878)] · e419d3ac 668 [] · 4e153a35 311 [] · de850030 1338 [(1289,1338)]
1786f59d 462 [(401,462)] · fbd9086a 767 [(717,767)] · 3d412ef6 573 [(491,573)] · 7f89b31f 54 []
48d501d6 679 [(610,679)] · 16a24e10 323 [] · f85d8652 642 [(562,642)] · a245f4a0 1110 [(1043,1110)] · fff7d335 982 [(933,982)] · e8215296 729 [(653,729)] · 00db816c 161 [(112,161)]
f1f86f15 972 [(893,972)] · 4f22f193 870 [(815,870)] · 5080ae86 428 [(354,428)] · 50d4b15d 706 [(643,706)] · 51752410 1114 [(1071,1114)]
1ea5f4fe 777 [(706,777)] · 1dca3750 2478 [(2404,2478)] · 5588d71b 2965 [(2879,2965)] · 3c7a761b 268 [] · 47a6097e 505 [(445,505)] · 3fdfefac 774 [(706,774)] · 4ce4477f 784 [(725,784)]
1f0a8025 120 [] · 34d25864 301 [(191,301)] · 5718ff5b 1636 [(19,121),(1567,1636)] · 43e1313c 284 [(227,284)] · 614f0b81 574 [(514,574)]
65edf921 796 [(730,796)] · 6c13aa1b 1252 [(979,1042),(1200,1252)] · 76943960 535 [(467,535)] · 7469da48 630 [(570,630)] · 749eee80 748 [(688,748)] · 78d916ef 906 [(836,906)]
901b9fde 378 [(308,378)] · 8f8b4d17 381 [(313,381)] · fcf0469c 456 [] · 868324b8 1456 [(1375,1456)] · 093448b4 156 [(90,156)] · 75918ce1 410 [(346,410)]
9d995e00 564 [(504,564)] · 9db50907 615 [(566,615)] · c1e96f7a 1361 [(1290,1361)] · c8ec883f 1768 [(1693,1768)] · 9f97bf3f 521 [(461,521)] · aacf7b7d 1023 [(965,1023)]
aa9008d4 745 [(683,745)] · 913371a0 259 [(214,259)] · 917970d5 580 [(528,580)] · b80361f4 643 [(565,643)] · 6c4c3dbe 177 [(103,177)] · abd1a6a5 498 [(416,498)]
"""


def _appendix_b():
    out = {}
    for line in APPENDIX_B.strip().splitlines():
        for item in line.split("·"):
            pref, n, ivs = item.strip().split(" ", 2)
            out[pref] = (int(n), [tuple(t) for t in eval(ivs)])
    return out


@pytest.fixture(scope="module")
def fixture72():
    z = np.load(os.path.join(GOLD, "smooth_fixture.npz"))
    meta = json.load(open(os.path.join(GOLD, "smooth_fixture.json")))
    return z, meta


def test_fixture_reads_match_appendix_b(fixture72, dcref):
    z, meta = fixture72
    table = _appendix_b()
    assert len(meta["ids"]) == 72            # 6 files x 12 reads; test_load_predict asserts 12 for chunk0/0.pt
    offs = z["offsets"]
    seen = 0
    for r, rid in enumerate(meta["ids"]):
        lab = z["labels"][offs[r]:offs[r + 1]]
        got = S.smooth_label_region(lab.tolist(), 21, 13, 20)
        assert [tuple(x) for x in meta["intervals"][r]] == got
        n, ivs = table[rid[:8]]
        assert n == lab.size and ivs == got, rid
        seen += 1
        # C oracle, same answer through its own code path
        sm = dcref.majority_voting(lab, 21)
        reg = [iv for iv in dcref.get_label_region(sm) if iv[1] - iv[0] >= 13]
        assert reg == got
    assert seen == 72


def test_c_oracle_batch_equals_python_on_fixture(fixture72, dcref):
    z, meta = fixture72
    offs = z["offsets"]
    lens = np.diff(offs).astype(np.int32)
    for ocq in (0, 1):
        for ct in (0, 1, 2):
            res = dcref.smooth_chop(z["labels"], offs[:-1], lens, chop_type=ct, ocq=ocq)
            opt = S.ChopOptions(chop_type=["terminal", "internal", "all"][ct], output_chopped_seqs=bool(ocq))
            for r in range(lens.size):
                act, ad, keep = S.chop_coordinates(z["labels"][offs[r]:offs[r + 1]].tolist(), None, opt)
                assert res["action"][r] == act
                assert res["n_adapter"][r] == len(ad)
                assert [tuple(x) for x in res["adapter_iv"][r][:len(ad)]] == ad
                assert res["n_keep"][r] == len(keep)
                assert [tuple(x) for x in res["keep_iv"][r][:len(keep)]] == keep


def test_c_oracle_equals_python_random():
    from oracle import cref
    c = cref.load()
    rng = np.random.default_rng(7)
    for trial in range(300):
        n = int(rng.integers(0, 400)) if trial % 3 else int(rng.integers(0, 30))
        lab = _random_labels(rng, n) if n else np.zeros(0, np.int8)
        for w in (1, 2, 3, 4, 21, 22, 51):
            assert c.majority_voting(lab, w).tolist() == S.majority_voting(lab.tolist(), w)
        assert c.get_label_region(lab) == S.get_label_region(lab.tolist())
    # batched chop coordinates incl. odd parameter settings and truncated reads
    labs = [_random_labels(rng, int(rng.integers(1, 900))) for _ in range(200)]
    lens = np.array([l.size for l in labs], np.int32)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    qual_lens = lens.copy()
    qual_lens[::17] += 5
    for (w, mi, ap, mp, mc, mr, ct, ocq) in [(21, 13, 20, 4, 20, 150, 2, 0), (11, 5, 3, 2, 10, 0, 0, 0),
                                             (21, 13, 20, 4, 20, 150, 1, 0), (5, 1, 20, 20, 1, 10, 2, 1),
                                             (2, 2, 1, 1, 50, 20, 2, 0)]:
        res = c.smooth_chop(np.concatenate(labs), starts, lens, qual_lens, w, mi, ap, mp, mc, mr, ct, ocq)
        opt = S.ChopOptions(w, mi, ap, mp, mc, bool(ocq), ["terminal", "internal", "all"][ct])
        old = S.MIN_READ_LEN
        S.MIN_READ_LEN = mr
        try:
            for r, lab in enumerate(labs):
                act, ad, keep = S.chop_coordinates(lab.tolist(), int(qual_lens[r]), opt)
                assert res["action"][r] == act, (r, w)
                assert res["n_adapter"][r] == len(ad)
                assert [tuple(x) for x in res["adapter_iv"][r][:len(ad)]] == ad
                assert res["n_keep"][r] == len(keep)
                assert [tuple(x) for x in res["keep_iv"][r][:len(keep)]] == keep
        finally:
            S.MIN_READ_LEN = old


def test_chop_records_end_to_end_small():
    """process_record gating (src/bin/predict.rs:137-187) + record naming (split.rs:109-117,203-223)."""
    n = 400
    lab = [0] * n
    for i in range(300, 400):
        lab[i] = 1
    seq = "ACGT" * 100
    qual = "I" * n
    pred = {"r1": S.Predict(lab, seq, "r1"), "short": S.Predict([1] * 100, "A" * 100, "short")}
    fq = [S.FastqRecord("r1", "desc here", seq, qual), S.FastqRecord("nopred", "", "AC", "II"),
          S.FastqRecord("short", "d", "A" * 100, "I" * 100)]
    out = S.chop_records(fq, pred, S.ChopOptions())
    assert [r.name for r in out] == ["r1|0:300|T", "short"]
    assert out[0].seq == seq[:300] and out[0].qual == qual[:300] and out[0].description == ""
    assert out[1].description == "d"      # untouched reads are passed through verbatim
    out = S.chop_records(fq, pred, S.ChopOptions(output_chopped_seqs=True))
    assert [r.name for r in out] == ["r1|300:400", "short"]
    out = S.chop_records(fq, pred, S.ChopOptions(chop_type="internal"))
    assert [r.name for r in out] == ["r1", "short"] and out[0].description == ""
    # internal adapter: two kept pieces, trailing piece loses its last base (T7)
    lab2 = [0] * n
    for i in range(100, 150):
        lab2[i] = 1
    out = S.chop_records([fq[0]], {"r1": S.Predict(lab2, seq, "r1")}, S.ChopOptions())
    assert [r.name for r in out] == ["r1|0:100|I", "r1|150:399|I"]
    assert out[1].seq == seq[150:399]
