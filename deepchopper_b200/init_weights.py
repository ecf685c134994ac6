"""Random-init state dict with the exact tensor names/shapes of the reference checkpoint
(``TokenClassificationLit`` -> ``net.backbone`` HyenaDNA-small-32k + ``net.head``).
Used by bench.py / the CLI's ``--random-init`` (no network: the trained ``yangliz5/deepchopper``
weights cannot be fetched here).  Initialisers follow torch's module defaults."""
from __future__ import annotations

import math
from typing import Dict

import torch

D, LAYERS, INNER, ORDER, EMB, VOCAB, LMAX, HEAD = 256, 4, 1024, 64, 5, 16, 32770, 1024


def _linear(g, out_f, in_f, bias=True):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand((out_f, in_f), generator=g) * 2 - 1) * bound
    b = (torch.rand((out_f,), generator=g) * 2 - 1) * bound if bias else None
    return w, b


def positional_table(lmax: int = LMAX, emb_dim: int = EMB):
    """HyenaPositionalEmbedding: z = [t, Re exp(-j f w), Im exp(-j f w)]."""
    t = torch.linspace(0, 1, lmax)[None, :, None]
    bands = (emb_dim - 1) // 2
    t_rescaled = torch.linspace(0, lmax - 1, lmax)[None, :, None]
    w = 2 * math.pi * t_rescaled / lmax
    f = torch.linspace(1e-4, bands - 1, bands)[None, None]
    z = torch.exp(-1j * f * w)
    return torch.cat([t, z.real, z.imag], dim=-1), t


def random_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    p = "net.backbone.backbone."
    sd[p + "embeddings.word_embeddings.weight"] = torch.randn((VOCAB, D), generator=g)
    z, t = positional_table()
    for l in range(LAYERS):
        q = f"{p}layers.{l}."
        w, b = _linear(g, 3 * D, D)
        sd[q + "mixer.in_linear.weight"], sd[q + "mixer.in_linear.bias"] = w, b
        w, b = _linear(g, D, D)
        sd[q + "mixer.out_linear.weight"], sd[q + "mixer.out_linear.bias"] = w, b
        bound = 1.0 / math.sqrt(3)
        sd[q + "mixer.short_filter.weight"] = (torch.rand((3 * D, 1, 3), generator=g) * 2 - 1) * bound
        sd[q + "mixer.short_filter.bias"] = (torch.rand((3 * D,), generator=g) * 2 - 1) * bound
        f = q + "mixer.filter_fn."
        sd[f + "bias"] = torch.randn((D,), generator=g)
        sd[f + "pos_emb.z"] = z.clone()
        sd[f + "pos_emb.t"] = t.clone()
        w, b = _linear(g, ORDER, EMB)
        sd[f + "implicit_filter.0.weight"], sd[f + "implicit_filter.0.bias"] = w, b
        for i in (1, 3, 5):
            sd[f + f"implicit_filter.{i}.freq"] = 10.0 * torch.ones((1, ORDER))
        for i in (2, 4):
            w, b = _linear(g, ORDER, ORDER)
            sd[f + f"implicit_filter.{i}.weight"], sd[f + f"implicit_filter.{i}.bias"] = w, b
        sd[f + "implicit_filter.6.weight"] = _linear(g, D, ORDER, bias=False)[0]
        sd[f + "modulation.deltas"] = torch.linspace(math.log(1e-2) / 1.5, math.log(1e-2) / 0.3, D)[None, None]
        for n in ("norm1", "norm2"):
            sd[q + n + ".weight"] = torch.ones(D)
            sd[q + n + ".bias"] = torch.zeros(D)
        w, b = _linear(g, INNER, D)
        sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = w, b
        w, b = _linear(g, D, INNER)
        sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = w, b
    sd[p + "ln_f.weight"] = torch.ones(D)
    sd[p + "ln_f.bias"] = torch.zeros(D)
    h = "net.head."
    w, b = _linear(g, HEAD, D)
    sd[h + "linear1.weight"], sd[h + "linear1.bias"] = w, b
    w, b = _linear(g, HEAD, HEAD)
    sd[h + "linear2.weight"], sd[h + "linear2.bias"] = w, b
    w, b = _linear(g, 2, HEAD)
    sd[h + "linear3.weight"], sd[h + "linear3.bias"] = w, b
    return sd
