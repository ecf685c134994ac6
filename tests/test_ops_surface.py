"""CPU checks of the host surface added around the C ABI: torch.library ops are registered and CUDA-only, the PyO3-named
loaders decode the reference's ``.pt`` layout like the oracle restatement, the package re-exports the reference names."""
import numpy as np
import pytest
import torch

from oracle import smooth_ref as R


def test_custom_ops_registered_cuda_only():
    import deepchopper_b200.ops  # noqa: F401
    for name in ("encode", "forward", "smooth_chop"):
        assert hasattr(torch.ops.dcb200, name)
    with pytest.raises(NotImplementedError):      # no CPU kernel, no fallback
        torch.ops.dcb200.smooth_chop(torch.zeros(4, dtype=torch.int8), torch.zeros(1, dtype=torch.int64),
                                     torch.tensor([4], dtype=torch.int32), torch.empty(0, dtype=torch.int32),
                                     [21, 13, 20, 4, 20, 150, 2, 0])
    with pytest.raises(NotImplementedError):
        torch.ops.dcb200.forward(torch.zeros((1, 128), dtype=torch.uint8), torch.zeros((1, 128)), 1, True, True)
    # fake (meta) implementations give the output shapes without a device
    tok, q = torch.ops.dcb200.encode(torch.empty(10, dtype=torch.uint8, device="meta"),
                                     torch.empty(3, dtype=torch.int64, device="meta"),
                                     torch.empty(3, dtype=torch.int64, device="meta"),
                                     torch.empty(3, dtype=torch.int32, device="meta"), 100, 128)
    assert tok.shape == (3, 128) and q.shape == (3, 128) and q.dtype == torch.float32
    lg, lb = torch.ops.dcb200.forward(torch.empty((2, 256), dtype=torch.uint8, device="meta"),
                                      torch.empty((2, 256), device="meta"), 1, True, True)
    assert lg.shape == (2, 256, 2) and lb.shape == (2, 256) and lb.dtype == torch.uint8


def _fake_pt(tmp_path, name, rng, B=5, L=300):
    lens = rng.integers(20, L - 1, B)
    pred = torch.from_numpy(rng.standard_normal((B, L, 2)).astype(np.float32))
    pred[0, :, 1] = pred[0, :, 0]                     # exact ties -> class 0
    target = torch.full((B, L), -100, dtype=torch.int64)
    seq = torch.full((B, L), 4, dtype=torch.int64)
    idt = torch.zeros((B, 256), dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        target[b, L - 1 - n:L - 1] = 0
        seq[b, L - 1 - n:L - 1] = torch.from_numpy(rng.integers(6, 12, n))
        seq[b, L - 1] = 1
        rid = f"read_{name}_{b}"
        idt[b, 0], idt[b, 1] = len(rid), int(b == 2)
        idt[b, 2:2 + len(rid)] = torch.tensor([ord(c) for c in rid])
    d = {"prediction": pred, "target": target, "seq": seq, "qual": torch.zeros((B, L)), "id": idt}
    path = tmp_path / f"0_{name}.pt"
    torch.save(d, path)
    return path, d


def test_load_predicts_names_match_oracle(tmp_path):
    import deepchopper_b200 as D
    rng = np.random.default_rng(3)
    p0, d0 = _fake_pt(tmp_path, "a", rng)
    p1, d1 = _fake_pt(tmp_path, "b", rng)
    got = D.load_predicts_from_batch_pt(p0, -100)
    want = R.load_predicts_from_batch(d0["prediction"].numpy(), d0["target"].numpy(), d0["seq"].numpy(), d0["id"].numpy())
    assert set(got) == set(want) and len(got) == 5
    for k in want:
        assert list(got[k].prediction) == list(want[k].prediction)
        assert got[k].seq == want[k].seq and got[k].is_truncated == want[k].is_truncated and got[k].id == k
    assert sum(got["read_a_0"].prediction) == 0       # ties -> class 0 (src/smooth/predict.rs:275)
    (tmp_path / "broken.pt").write_bytes(b"not a pickle")
    both = D.load_predicts_from_batch_pts(tmp_path)    # reports and skips the broken file
    assert len(both) == 10
    assert len(D.load_predicts_from_batch_pts(tmp_path, -100, 1)) == 5


def test_package_reexports_reference_names():
    import deepchopper_b200 as D
    for name in ("majority_voting", "get_label_region", "smooth_label_region", "remove_intervals_and_keep_left",
                 "summary_predict", "id_list2seq", "Predict", "encode_qual", "normalize_seq",
                 "load_predicts_from_batch_pt", "load_predicts_from_batch_pts", "predict_cli", "StatResult",
                 "DeepChopper"):
        assert getattr(D, name) is not None
    import inspect
    sig = inspect.signature(D.predict_cli)            # src/python.rs:829-842
    assert list(sig.parameters) == ["predicts", "fq", "smooth_window_size", "min_interval_size", "approved_interval_number",
                                    "max_process_intervals", "min_read_length_after_chop", "output_chopped_seqs",
                                    "chop_type", "threads", "output_prefix", "max_batch_size"]
    assert [sig.parameters[k].default for k in list(sig.parameters)[2:]] == [21, 13, 20, 4, 20, False, "all", 2, None, None]
    sig = inspect.signature(D.load_predicts_from_batch_pts)
    assert [p.default for p in sig.parameters.values()][1:] == [-100, None]
