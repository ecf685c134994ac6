"""CPU ORACLE (test infrastructure, NOT the product): fp32 PyTorch restatement of DeepChopper's
``predict`` model path -- HyenaDNA-small-32k backbone + quality-aware token-classification head.

PARITY UNPINNED for the backbone: the arithmetic lives in third-party HuggingFace remote code
(``LongSafari/hyenadna-small-32k-seqlen-hf`` ``modeling_hyena.py``, revision unpinned by the
reference, loaded at deepchopper/models/llm/hyena.py:22) which is absent from /root/reference and
unreachable (no network).  This file restates its published algorithm; the
reference's own tests never run the model.  The head IS pinned: tests/test_oracle_model.py checks
``RefHead`` against the reference's own deepchopper/models/llm/head.py when /root/reference exists.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class HyenaConfig:
    # configuration_hyena.py of hyenadna-small-32k-seqlen-hf
    d_model: int = 256
    n_layer: int = 4
    d_inner: int = 1024
    vocab_size: int = 12
    pad_vocab_size_multiple: int = 8
    emb_dim: int = 5
    filter_order: int = 64
    num_inner_mlps: int = 2
    hyena_order: int = 2
    short_filter_order: int = 3
    max_seq_len: int = 32770
    activation_freq: float = 10.0
    layer_norm_epsilon: float = 1e-5
    # head: deepchopper/models/dc_hg.py:119-163 / configs/model/hyena.yaml:16-28
    head_hidden: int = 1024
    num_class: int = 2

    @property
    def vocab_padded(self) -> int:
        m = self.pad_vocab_size_multiple
        return (self.vocab_size + m - 1) // m * m


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


class RefSin(nn.Module):
    """HyenaSin: one learnable ``freq [1, filter_order]`` shared by all sine activations."""

    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.freq = nn.Parameter(cfg.activation_freq * torch.ones(1, cfg.filter_order))

    def forward(self, x):
        return torch.sin(self.freq * x)


class RefPositionalEmbedding(nn.Module):
    """HyenaPositionalEmbedding: z = [t, Re exp(-j f w), Im exp(-j f w)], built for max_seq_len."""

    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        L = cfg.max_seq_len
        t = torch.linspace(0, 1, L)[None, :, None]
        bands = (cfg.emb_dim - 1) // 2
        t_rescaled = torch.linspace(0, L - 1, L)[None, :, None]
        w = 2 * math.pi * t_rescaled / L
        f = torch.linspace(1e-4, bands - 1, bands)[None, None]
        z = torch.exp(-1j * f * w)
        z = torch.cat([t, z.real, z.imag], dim=-1)
        self.z = nn.Parameter(z)
        self.register_buffer("t", t)

    def forward(self, L):
        return self.z[:, :L], self.t[:, :L]


class RefModulation(nn.Module):
    def __init__(self, d_model, fast_decay_pct=0.3, slow_decay_pct=1.5, target=1e-2, shift=0.05):
        super().__init__()
        self.shift = shift
        max_decay = math.log(target) / fast_decay_pct
        min_decay = math.log(target) / slow_decay_pct
        self.deltas = nn.Parameter(torch.linspace(min_decay, max_decay, d_model)[None, None])

    def forward(self, t, x):
        return x * (torch.exp(-t * self.deltas.abs()) + self.shift)


class RefFilter(nn.Module):
    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.bias = nn.Parameter(torch.randn(cfg.d_model))
        act = RefSin(cfg)
        self.pos_emb = RefPositionalEmbedding(cfg)
        layers = [nn.Linear(cfg.emb_dim, cfg.filter_order), act]
        for _ in range(cfg.num_inner_mlps):
            layers += [nn.Linear(cfg.filter_order, cfg.filter_order), act]
        layers.append(nn.Linear(cfg.filter_order, cfg.d_model, bias=False))
        self.implicit_filter = nn.Sequential(*layers)
        self.modulation = RefModulation(cfg.d_model)

    def filter(self, L):
        z, t = self.pos_emb(L)
        return self.modulation(t, self.implicit_filter(z))  # [1, L, d]


def fftconv_ref(u, k, D):
    """modeling_hyena.py fftconv: causal long conv via rfft of size 2L, plus the ``u * D`` skip."""
    L = u.shape[-1]
    n = 2 * L
    k_f = torch.fft.rfft(k, n=n) / n
    u_f = torch.fft.rfft(u.to(k.dtype), n=n)
    y = torch.fft.irfft(u_f * k_f, n=n, norm="forward")[..., :L]
    return (y + u * D.unsqueeze(-1)).to(u.dtype)


class RefOperator(nn.Module):
    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        d = cfg.d_model
        self.d_model = d
        self.l_max = cfg.max_seq_len
        inner = d * (cfg.hyena_order + 1)
        self.in_linear = nn.Linear(d, inner)
        self.out_linear = nn.Linear(d, d)
        self.short_filter = nn.Conv1d(inner, inner, cfg.short_filter_order, padding=cfg.short_filter_order - 1,
                                      groups=inner)
        self.filter_fn = RefFilter(cfg)

    def forward(self, u, emulate_bf16=False):
        L = u.size(-2)
        lf = min(L, self.l_max)
        if emulate_bf16:
            z = F.linear(_bf16(u), _bf16(self.in_linear.weight), self.in_linear.bias)
            z = _bf16(z)                      # in_proj output is stored bf16, channel-major
        else:
            z = self.in_linear(u)
        z = z.transpose(1, 2)                 # b d l
        zc = self.short_filter(z)[..., :lf]
        x0, x1, v = zc.split(self.d_model, dim=1)
        k = self.filter_fn.filter(lf)[0].transpose(0, 1)  # d l
        v = v * x1
        v = fftconv_ref(v, k, self.filter_fn.bias)
        y = (v * x0).transpose(1, 2)
        if emulate_bf16:
            return F.linear(_bf16(y), _bf16(self.out_linear.weight), self.out_linear.bias)
        return self.out_linear(y)


class RefMlp(nn.Module):
    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.fc1 = nn.Linear(cfg.d_model, cfg.d_inner)
        self.fc2 = nn.Linear(cfg.d_inner, cfg.d_model)

    def forward(self, x, emulate_bf16=False):
        if emulate_bf16:
            g = F.gelu(F.linear(_bf16(x), _bf16(self.fc1.weight), self.fc1.bias), approximate="tanh")
            return F.linear(_bf16(g), _bf16(self.fc2.weight), self.fc2.bias)
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


class RefBlock(nn.Module):
    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.mixer = RefOperator(cfg)
        self.norm1 = nn.LayerNorm(cfg.d_model, eps=cfg.layer_norm_epsilon)
        self.mlp = RefMlp(cfg)
        self.norm2 = nn.LayerNorm(cfg.d_model, eps=cfg.layer_norm_epsilon)

    def forward(self, h, emulate_bf16=False):
        r = h
        h = self.mixer(self.norm1(r), emulate_bf16) + r
        r = h
        return self.mlp(self.norm2(r), emulate_bf16) + r


class RefEmbeddings(nn.Module):
    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_padded, cfg.d_model)

    def forward(self, ids):
        return self.word_embeddings(ids)


class RefBackbone(nn.Module):
    """HyenaLMBackbone: embeddings -> n_layer x HyenaBlock -> ln_f."""

    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.embeddings = RefEmbeddings(cfg)
        self.layers = nn.ModuleList([RefBlock(cfg) for _ in range(cfg.n_layer)])
        self.ln_f = nn.LayerNorm(cfg.d_model, eps=cfg.layer_norm_epsilon)

    def forward(self, ids, emulate_bf16=False):
        h = self.embeddings(ids)
        for layer in self.layers:
            h = layer(h, emulate_bf16)
        return self.ln_f(h)


class RefHyenaDNAModel(nn.Module):
    """HyenaDNAModel: ``.backbone`` holds the HyenaLMBackbone (state-dict prefix ``backbone.``)."""

    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.backbone = RefBackbone(cfg)

    def forward(self, ids, emulate_bf16=False):
        return self.backbone(ids, emulate_bf16)


class RefHead(nn.Module):
    """deepchopper/models/llm/head.py:39-102 with use_identity_layer_for_qual=True, use_qual=True."""

    def __init__(self, cfg: HyenaConfig):
        super().__init__()
        self.linear1 = nn.Linear(cfg.d_model, cfg.head_hidden)
        self.linear2 = nn.Linear(cfg.head_hidden, cfg.head_hidden)
        self.linear3 = nn.Linear(cfg.head_hidden, cfg.num_class)

    def forward(self, x, input_quals, emulate_bf16=False):
        if emulate_bf16:
            o = F.relu(F.linear(_bf16(x), _bf16(self.linear1.weight), self.linear1.bias))
            r = o + input_quals.unsqueeze(-1)
            o = F.relu(F.linear(_bf16(r), _bf16(self.linear2.weight), self.linear2.bias) + r)
            return self.linear3(o)
        o = F.relu(self.linear1(x))                                   # head.py:94
        r = o + input_quals.unsqueeze(-1)                             # head.py:97
        o = F.relu(self.linear2(r) + r)                               # head.py:98
        return self.linear3(o)                                        # head.py:102


class RefTokenClassificationModule(nn.Module):
    """deepchopper/models/llm/hyena.py:9-41: ``head(backbone(input_ids)[0], input_quals)``."""

    def __init__(self, cfg: HyenaConfig | None = None):
        super().__init__()
        self.cfg = cfg or HyenaConfig()
        self.backbone = RefHyenaDNAModel(self.cfg)
        self.head = RefHead(self.cfg)

    def forward(self, input_ids, input_quals, emulate_bf16=False):
        return self.head(self.backbone(input_ids, emulate_bf16), input_quals, emulate_bf16)


class RefDeepChopper(nn.Module):
    """TokenClassificationLit (deepchopper/models/basic_module.py:34-100,197-207), inference part:
    state-dict prefix ``net.``; ``predict_step`` returns ``(logits, batch['labels'])``."""

    def __init__(self, cfg: HyenaConfig | None = None):
        super().__init__()
        self.net = RefTokenClassificationModule(cfg)

    def forward(self, input_ids, input_quals, emulate_bf16=False):
        return self.net(input_ids, input_quals, emulate_bf16)

    def predict_step(self, batch, batch_idx=0):
        return self.forward(batch["input_ids"], batch["input_quals"]), batch["labels"]


def make_reference_model(seed: int = 0, cfg: HyenaConfig | None = None) -> RefDeepChopper:
    """Deterministic random init (torch default initialisers under ``torch.manual_seed(seed)``)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = RefDeepChopper(cfg).eval()
    torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


# ---- host-side tokenise / collate restatement (tokenizer.py:145-178, :34-93) -------------------

CHAR2ID = {"A": 7, "C": 8, "G": 9, "T": 10, "N": 11}
PAD_ID, SEP_ID, UNK_ID = 4, 1, 6


def normalize_seq(seq: str) -> str:
    """src/python.rs:272-275 -> needletail Sequence::normalize(iupac=false): upper-case, U->T,
    '.'/'~' -> '-', other non-ACGTN -> N.  (third-party needletail 0.7.3; no reference test pins it)."""
    out = []
    for ch in seq.upper():
        if ch in "ACGTN-":
            out.append(ch)
        elif ch == "U":
            out.append("T")
        elif ch in ".~":
            out.append("-")
        elif ch in " \t\r\n":
            continue
        else:
            out.append("N")
    return "".join(out)


def encode_qual(qual: str, offset: int = 33):
    """src/python.rs:25-35."""
    return [ord(c) - offset for c in qual]


def tokenize_read(read_id: str, seq: str, qual: str, max_length: int = 32768, max_id_length: int = 256):
    """tokenizer.py:145-178 for the predict path (target == [0, 0])."""
    seq = normalize_seq(seq)
    ids = [CHAR2ID.get(c, UNK_ID) for c in seq]
    q = torch.tensor(encode_qual(qual), dtype=torch.float32)
    truncation = len(seq) >= max_length
    if truncation:
        ids = ids[: max_length - 1]
        q = q[: max_length - 1]
    n = len(ids)
    ids = ids + [SEP_ID]
    labels = [0] * n + [-100]
    quals = F.normalize(torch.cat((q, torch.tensor([0.0]))).float(), dim=0)
    new_id = [len(read_id), int(truncation)] + [ord(c) for c in read_id]
    new_id = new_id[:max_id_length] + [0] * max(0, max_id_length - len(new_id))
    return {"input_ids": ids, "labels": labels, "input_quals": quals, "id": new_id}


def collate(features, pad_to: int | None = None):
    """DataCollatorForTokenClassificationWithQual.torch_call (tokenizer.py:34-93), LEFT padding."""
    L = max(len(f["input_ids"]) for f in features)
    if pad_to is not None:
        assert pad_to >= L
        L = pad_to
    ids = torch.full((len(features), L), PAD_ID, dtype=torch.int64)
    labels = torch.full((len(features), L), -100, dtype=torch.int8)
    quals = torch.zeros((len(features), L), dtype=torch.float32)
    idt = torch.zeros((len(features), len(features[0]["id"])), dtype=torch.int8)
    for b, f in enumerate(features):
        n = len(f["input_ids"])
        ids[b, L - n:] = torch.tensor(f["input_ids"], dtype=torch.int64)
        labels[b, L - n:] = torch.tensor(f["labels"], dtype=torch.int8)
        quals[b, L - n:] = f["input_quals"]
        idt[b] = torch.tensor(f["id"], dtype=torch.int64).to(torch.int8)
    return {"input_ids": ids, "labels": labels, "input_quals": quals, "id": idt}
