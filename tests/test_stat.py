"""StatResult / collect_statistics_for_predicts (src/smooth/stat.rs): oracle restatement on hand-checked cases (CPU) and
the GPU-backed host mirror against the oracle on the 72 fixture reads and random planted reads."""
import pickle

import numpy as np
import pytest

from oracle import smooth_ref as S


def _mk(labels, seq=None, rid="r", trunc=False):
    seq = seq or ("C" * len(labels))
    return S.Predict(list(labels), seq, rid, trunc)


def test_oracle_statistics_hand_checked():
    n = 200
    lab = [0] * n
    for i in range(100, 130):
        lab[i] = 1                      # one clean 30-base run ending at 130: relative position 0.65
    seq = list("C" * n)
    seq[95:100] = "AACAA"               # 4 A in the 5 bases before the run
    a = _mk(lab, "".join(seq), "a")
    short = _mk([1] * 100, "C" * 100, "short")                       # < MIN_READ_LEN: skipped entirely
    lab2 = [0] * n
    lab2[0] = 1                                                      # T5 quirk: vanishes
    lab2[50:53] = [1, 1, 1]                                          # raw run, too short after smoothing
    b = _mk(lab2, "G" * n, "b", trunc=True)
    r = S.collect_statistics_for_predicts([a, short, b], 21, 13, 20, 0.9, 3)
    assert r["total_predicts"] == 2 and r["total_truncated"] == 1
    assert r["predicts_with_chop"] == ["a", "b"]
    assert r["original_intervals"] == {"a": [(100, 130)], "b": [(50, 53)]}
    assert r["smooth_predicts_with_chop"] == ["a"] and r["smooth_intervals"] == {"a": [(100, 130)]}
    assert r["smooth_only_one"] == ["a"] and r["smooth_only_one_with_ploya"] == ["a"]
    assert r["smooth_intervals_relative_pos"] == [float(np.float32(130) / np.float32(200))]
    assert r["smooth_internal_predicts"] == ["a"]
    assert S.collect_statistics_for_predicts([a], 21, 13, 20, 0.5, 5)["smooth_only_one_with_ploya"] == []
    assert S.collect_statistics_for_predicts([a], 21, 13, 20, 0.5, 5)["smooth_internal_predicts"] == []


def _as_dict(st):
    from dataclasses import asdict
    d = asdict(st)
    d["smooth_intervals"] = {k: [tuple(x) for x in v] for k, v in d["smooth_intervals"].items()}
    d["original_intervals"] = {k: [tuple(x) for x in v] for k, v in d["original_intervals"].items()}
    return d


@pytest.mark.gpu
def test_statresult_matches_oracle_on_fixture_and_random():
    from deepchopper_b200 import synth
    from deepchopper_b200.smooth import Predict
    from deepchopper_b200.stat import StatResult, collect_statistics_for_predicts
    import json
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "smooth_fixture.npz"))
    meta = json.load(open(os.path.join(gold, "smooth_fixture.json")))
    offs = z["offsets"]
    rng = np.random.default_rng(17)
    reads = []
    for r, rid in enumerate(meta["ids"]):                # the 72 real prediction reads of tests/data/eval
        labels = z["labels"][offs[r]:offs[r + 1]]
        seq = "".join(rng.choice(list("ACGT"), len(labels)))
        reads.append((rid, list(map(int, labels)), seq, bool(z["truncated"][r])))
    lens = np.concatenate([synth.read_lengths(rng, 300, hi=3000), rng.integers(10, 200, 40)])
    lab, starts, _ = synth.planted_labels(rng, lens)
    for i, n in enumerate(lens.tolist()):
        l = lab[starts[i]:starts[i] + n].astype(int).tolist()
        if i % 9 == 0 and n:
            l[0] = 1                                   # runs touching index 0 (T5 quirk)
        reads.append((f"syn{i}", l, "".join(rng.choice(list("AAACGT"), n)), False))
    ours = [Predict(l, s, i, t) for i, l, s, t in reads]
    want_in = [S.Predict(l, s, i, t) for i, l, s, t in reads]
    for args in [(21, 13, 20, 0.9, 3), (11, 5, 2, 0.5, 2), (1, 1, 100, 1.0, 0)]:
        got = collect_statistics_for_predicts(ours, *args)
        want = S.collect_statistics_for_predicts(want_in, *args)
        assert _as_dict(got) == want, args
    # methods + state
    st = collect_statistics_for_predicts(ours, 21, 13, 20, 0.9, 3)
    assert st.number_smooth_predicts_with_chop() == [len(st.smooth_intervals[i]) for i in st.smooth_predicts_with_chop]
    assert sum(st.length_predicts_with_chop()) == sum(e - s for v in st.original_intervals.values() for s, e in v)
    assert set(st.selected_predict_by_intervals(2)) == {i for i, v in st.smooth_intervals.items() if len(v) >= 2}
    clone = pickle.loads(pickle.dumps(st))
    assert _as_dict(clone) == _as_dict(st)
    both = StatResult()
    both.merge(st)
    both.merge(clone)
    assert both.total_predicts == 2 * st.total_predicts and len(both.smooth_only_one) == 2 * len(st.smooth_only_one)
