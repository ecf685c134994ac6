"""Pins what can be pinned of the model oracle (oracle/hyena_ref.py):
 * RefHead == the reference's own deepchopper/models/llm/head.py (golden made by tests/golden/make_golden.py,
   and live when /root/reference is present),
 * tokenise/collate layout == the reference's collated fixture batch (left pad 4, SEP 1, unit-norm quals),
 * structural facts of HyenaDNA-small-32k (param counts, causal prefix property of the implicit filter)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import hyena_ref as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_head_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "head_golden.npz"))
    torch.manual_seed(int(g["seed"]))
    head = H.RefHead(H.HyenaConfig()).eval()
    with torch.no_grad():
        y = head(torch.from_numpy(g["x"]), torch.from_numpy(g["q"]))
    assert torch.equal(y, torch.from_numpy(g["y"]))   # same ops, same order -> bit-exact on CPU


@pytest.mark.skipif(not os.path.exists("/root/reference/deepchopper/models/llm/head.py"), reason="reference checkout absent")
def test_head_matches_reference_live():
    spec = importlib.util.spec_from_file_location("ref_head", "/root/reference/deepchopper/models/llm/head.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.TokenClassificationHead(256, 2, 1024, 1024, use_identity_layer_for_qual=True, use_qual=True).eval()
    mine = H.RefHead(H.HyenaConfig()).eval()
    mine.load_state_dict(ref.state_dict())
    x, q = torch.randn(3, 17, 256), torch.rand(3, 17)
    with torch.no_grad():
        assert torch.equal(ref(x, q), mine(x, q))


def test_param_counts():
    m = H.make_reference_model(0)
    assert sum(p.numel() for p in m.net.backbone.parameters()) == 3_933_992
    assert sum(p.numel() for p in m.net.head.parameters()) == 1_314_818       # from the real head.py
    keys = list(m.state_dict().keys())
    assert "net.backbone.backbone.embeddings.word_embeddings.weight" in keys
    assert "net.backbone.backbone.layers.3.mixer.filter_fn.implicit_filter.6.weight" in keys
    assert "net.head.linear3.bias" in keys


def test_filter_prefix_property():
    # filter(L) is the prefix of filter(Lmax) -> compute once per weight set
    m = H.make_reference_model(0)
    f = m.net.backbone.backbone.layers[1].mixer.filter_fn
    assert torch.equal(f.filter(300)[0], f.filter(1000)[0][:300])


def test_collate_layout_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "collate_golden.npz"))
    assert (g["last_tok"] == 1).all()                        # SEP appended, right-aligned (T1)
    assert np.allclose(g["norms"], 1.0, atol=1e-5)           # T3
    assert (g["pad_quals_max"] == 0).all()
    # rebuild row 0 through our tokenise+collate from decoded bases and a matching quality vector
    ids0 = g["row0_ids"]
    first = int(g["first_real"][0])
    bases = "".join({7: "A", 8: "C", 9: "G", 10: "T", 11: "N"}[int(t)] for t in ids0[first:-1])
    q = g["row0_quals"][first:-1].astype(np.float64)
    # recover integer phred up to the common scale: ratios to the max are rational with small ints
    scale = 1.0 / q[q > 0].min()
    cand = None
    for k in range(1, 94):
        v = q * scale * k
        if np.abs(v - np.round(v)).max() < 1e-3 and np.round(v).max() <= 93:
            cand = np.round(v).astype(int)
            break
    assert cand is not None
    qual = "".join(chr(int(v) + 33) for v in cand)
    feat = H.tokenize_read("x", bases, qual)
    batch = H.collate([feat], pad_to=ids0.size)
    assert np.array_equal(batch["input_ids"][0].numpy().astype(np.uint8), ids0)
    assert np.allclose(batch["input_quals"][0].numpy(), g["row0_quals"], rtol=1e-5, atol=1e-8)


def test_pad_sensitivity_is_real():
    # no attention mask -> left pads change the logits; the padded batch is the parity unit
    m = H.make_reference_model(0)
    ids = torch.randint(7, 11, (1, 200))
    ids[0, -1] = 1
    q = torch.rand(1, 200)
    a = m(ids, q)
    b = m(torch.cat([torch.full((1, 56), 4), ids], 1), torch.cat([torch.zeros(1, 56), q], 1))[:, 56:]
    assert (a - b).abs().max() > 1e-3
