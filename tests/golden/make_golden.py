"""Regenerates tests/golden/*.npz|json from the read-only reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

  python tests/golden/make_golden.py

* smooth_fixture.npz  -- the 72 real prediction reads of /root/reference/tests/data/eval/chunk{0,1}/{0,1,2}.pt
  reduced to what the hot path consumes: per-read argmax labels (target != -100 positions only), decoded token
  ids, read ids, truncated flag, plus the intervals the oracle derives (window 21, min 13, approved 20).
  The float logits themselves are not committed (MBs); argmax is `l1 > l0` (src/smooth/predict.rs:275).
* head_golden.npz     -- outputs of the reference's OWN deepchopper/models/llm/head.py (imported by file path)
  on seeded inputs/weights; pins oracle.hyena_ref.RefHead.
* collate_golden.npz  -- tests/data/input_ids.pt / input_quals.pt row facts (left pad, SEP, unit norm).
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import smooth_ref  # noqa: E402
from oracle.hyena_ref import HyenaConfig, RefHead  # noqa: E402


def smooth_fixture():
    labels, toks, ids, trunc, offs, ivs, files = [], [], [], [], [0], [], []
    for chunk in ("chunk0", "chunk1"):
        for i in range(3):
            path = f"{REF}/tests/data/eval/{chunk}/{i}.pt"
            d = torch.load(path)
            pred = d["prediction"].numpy()
            target = d["target"].to(torch.int64).numpy()     # chunk1/2.pt is still int8 (scripts/convert_pt_dtype.py)
            seq = d["seq"].numpy()
            idarr = d["id"].to(torch.int64).numpy()
            preds = smooth_ref.load_predicts_from_batch(pred, target, seq, idarr)
            lab = (pred[..., 1] > pred[..., 0]).astype(np.int8)
            for b in range(lab.shape[0]):
                keep = target[b] != -100
                n_id = int(idarr[b][0])
                rid = "".join(chr(int(c)) for c in idarr[b][2:2 + n_id])
                p = preds[rid]
                assert p.prediction == lab[b][keep].tolist()
                labels.append(lab[b][keep])
                toks.append(seq[b][keep].astype(np.uint8))
                ids.append(rid)
                trunc.append(bool(idarr[b][1]))
                offs.append(offs[-1] + int(keep.sum()))
                ivs.append(p.smooth_and_select_intervals(21, 13, 20))
                files.append(f"{chunk}/{i}.pt")
    np.savez_compressed(f"{OUT}/smooth_fixture.npz", labels=np.concatenate(labels), tokens=np.concatenate(toks),
                        offsets=np.array(offs, dtype=np.int64), truncated=np.array(trunc))
    with open(f"{OUT}/smooth_fixture.json", "w") as f:
        json.dump({"ids": ids, "files": files, "intervals": ivs,
                   "params": {"window": 21, "min_interval": 13, "approved": 20}}, f, indent=0)
    print("smooth fixture:", len(ids), "reads", offs[-1], "bases")


def head_golden():
    spec = importlib.util.spec_from_file_location("ref_head", f"{REF}/deepchopper/models/llm/head.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(1234)
    ref = mod.TokenClassificationHead(256, 2, 1024, 1024, use_identity_layer_for_qual=True, use_qual=True).eval()
    x = torch.randn(2, 24, 256)
    q = torch.rand(2, 24) * 0.1
    with torch.no_grad():
        y = ref(x, q)
    # weights are regenerated in the test from the same seed through RefHead (same ctor order: linear1,2,3)
    torch.manual_seed(1234)
    mine = RefHead(HyenaConfig()).eval()
    for (ka, va), (kb, vb) in zip(sorted(ref.state_dict().items()), sorted(mine.state_dict().items())):
        assert ka == kb and torch.equal(va, vb), ka
    np.savez_compressed(f"{OUT}/head_golden.npz", x=x.numpy(), q=q.numpy(), y=y.numpy(), seed=1234)
    print("head golden written", y.shape)


def collate_golden():
    ids = torch.load(f"{REF}/tests/data/input_ids.pt").numpy()
    quals = torch.load(f"{REF}/tests/data/input_quals.pt").numpy()
    first_real = (ids != 4).argmax(axis=1)
    np.savez_compressed(f"{OUT}/collate_golden.npz", first_real=first_real, last_tok=ids[:, -1],
                        norms=np.sqrt((quals.astype(np.float64) ** 2).sum(axis=1)),
                        row0_ids=ids[0].astype(np.uint8), row0_quals=quals[0],
                        pad_quals_max=np.array([np.abs(quals[b, :first_real[b]]).max() if first_real[b] else 0.0
                                                for b in range(ids.shape[0])]))
    print("collate golden written", ids.shape)


if __name__ == "__main__":
    smooth_fixture()
    head_golden()
    collate_golden()
