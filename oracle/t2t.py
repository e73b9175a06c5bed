"""Oracle: the reference's TensorFlow T2T-ViT restated in torch (test infrastructure).

PARITY UNPINNED: TensorFlow is not installed here; the reference's ``modeling/models/t2t_vit.py``
cannot run and ships no fixture.  Line-by-line restatement of

* ``tf_Unfold.call``        ``modeling/models/t2t_vit.py:20-40``  (channel-last, patch depth order ``(kh, kw, c)``
                            as ``tf.image.extract_patches`` lays it out; zero padding first)
* ``T2T_module.call``       ``modeling/models/t2t_vit.py:63-88``
* ``TokenPerformer``        ``modeling/layers/transformer_encoder.py:39-101``
* ``T2T_ViT``               ``modeling/models/t2t_vit.py:91-135``;  ``get_t2t_vit_14`` ``:147-148``
* sinusoid table            ``modeling/layers/embedding.py:4-15``

Weight names: ``t2t.performer{1,2}.{norm1,norm2}.{gamma,beta}``, ``.kqv.{kernel,bias}``, ``.w`` ([m, emb],
already multiplied by sqrt(m) as at ``transformer_encoder.py:65``), ``.attn_output.{kernel,bias}``,
``.mlp.fc{1,2}.{kernel,bias}``; ``t2t.project.{kernel,bias}``; ``cls_tokens``; ``pos_embedding``;
``layers.{l}.*`` as in ``oracle.tf_vit``; ``norm.{gamma,beta}``; ``classifier_head.{kernel,bias}``.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .tf_vit import TF_EPS, dense, glorot, init_encoder, tf_encoder, tf_ffn


def unfold_nhwc(x: torch.Tensor, k: int, s: int, p: int) -> torch.Tensor:
    """tf_Unfold(channel_last=True): [B,H,W,C] -> [B, oh*ow, k*k*C] with depth order (kh, kw, c)."""
    B, H, W, C = x.shape
    x = F.pad(x, (0, 0, p, p, p, p))
    oh = (H + 2 * p - k) // s + 1
    ow = (W + 2 * p - k) // s + 1
    win = x.unfold(1, k, s).unfold(2, k, s)              # [B, oh, ow, C, kh, kw]
    return win.permute(0, 1, 2, 4, 5, 3).reshape(B, oh * ow, k * k * C)


def prm_exp(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """transformer_encoder.py:67-81: exp(w x - |x|^2/2) / sqrt(m)."""
    m = w.shape[0]
    xd = (x * x).sum(-1, keepdim=True) / 2
    wtd = torch.einsum("bti,mi->btm", x, w)
    return torch.exp(wtd - xd) / math.sqrt(m)


def token_performer(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """TokenPerformer.call, transformer_encoder.py:96-101 (dropout is identity at inference)."""
    emb = sd[p + ".attn_output.kernel"].shape[0]
    x = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm1.gamma"], sd[p + ".norm1.beta"], TF_EPS)
    k, q, v = torch.split(dense(x, sd, p + ".kqv"), emb, dim=-1)           # :84
    kp, qp = prm_exp(k, sd[p + ".w"]), prm_exp(q, sd[p + ".w"])
    D = torch.einsum("bti,bi->bt", qp, kp.sum(dim=1)).unsqueeze(2)         # :86-87
    kptv = torch.einsum("bin,bim->bnm", v, kp)                             # :88
    y = torch.einsum("bti,bni->btn", qp, kptv) / (D + 1e-8)                # :90
    y = v + dense(y, sd, p + ".attn_output")                               # :93
    z = F.layer_norm(y, (emb,), sd[p + ".norm2.gamma"], sd[p + ".norm2.beta"], TF_EPS)
    return y + tf_ffn(sd, p + ".mlp", z)                                   # :99


def t2t_module(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """T2T_module.call, t2t_vit.py:63-88; x NHWC [B,224,224,3] -> [B,196,D]."""
    B = x.shape[0]
    x = unfold_nhwc(x, 7, 4, 2)
    x = token_performer(sd, "t2t.performer1", x)
    hw = int(math.isqrt(x.shape[1]))
    x = unfold_nhwc(x.reshape(B, hw, hw, -1), 3, 2, 1)
    x = token_performer(sd, "t2t.performer2", x)
    hw = int(math.isqrt(x.shape[1]))
    x = unfold_nhwc(x.reshape(B, hw, hw, -1), 3, 2, 1)
    return dense(x, sd, "t2t.project")


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """modeling/layers/embedding.py:4-15."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab).float()


@torch.no_grad()
def t2t_vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, num_heads: int, return_tokens=False):
    """T2T_ViT.call, t2t_vit.py:117-135; x NHWC."""
    D = sd["cls_tokens"].shape[-1]
    tok = t2t_module(sd, x.float())
    B = tok.shape[0]
    h = torch.cat((sd["cls_tokens"].expand(B, 1, -1), tok), dim=1) + sd["pos_embedding"]
    h = tf_encoder(sd, h, [num_heads] * depth, D // num_heads)
    h = F.layer_norm(h, (D,), sd["norm.gamma"], sd["norm.beta"], TF_EPS)
    logits = dense(h[:, 0], sd, "classifier_head")
    return (logits, tok) if return_tokens else logits


def init_t2t_vit(hidden=384, depth=14, num_heads=6, mlp_ratio=3.0, token_size=64, num_classes=1000, seed=0,
                 stress=False):
    """Seeded weights with Keras default initialisers; ``w`` = sqrt(m) * orthogonal [m, emb]
    (transformer_encoder.py:59-65).  Defaults = get_t2t_vit_14 (t2t_vit.py:147-148)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def vec(n, base):
        return base + (torch.randn(n, generator=g) * 0.1 if stress else torch.zeros(n))

    m = int(token_size * 0.5)
    for name, in_dim in (("t2t.performer1", 7 * 7 * 3), ("t2t.performer2", 3 * 3 * token_size)):
        sd[name + ".norm1.gamma"], sd[name + ".norm1.beta"] = vec(in_dim, 1.0), vec(in_dim, 0.0)
        sd[name + ".kqv.kernel"], sd[name + ".kqv.bias"] = glorot(g, in_dim, 3 * token_size), vec(3 * token_size, 0.0)
        q, _ = torch.linalg.qr(torch.randn(token_size, m, generator=g))
        sd[name + ".w"] = q.t().contiguous() * math.sqrt(m)
        sd[name + ".attn_output.kernel"], sd[name + ".attn_output.bias"] = glorot(g, token_size, token_size), vec(token_size, 0.0)
        sd[name + ".norm2.gamma"], sd[name + ".norm2.beta"] = vec(token_size, 1.0), vec(token_size, 0.0)
        sd[name + ".mlp.fc1.kernel"], sd[name + ".mlp.fc1.bias"] = glorot(g, token_size, token_size), vec(token_size, 0.0)
        sd[name + ".mlp.fc2.kernel"], sd[name + ".mlp.fc2.bias"] = glorot(g, token_size, token_size), vec(token_size, 0.0)
    sd["t2t.project.kernel"], sd["t2t.project.bias"] = glorot(g, 9 * token_size, hidden), vec(hidden, 0.0)
    sd["cls_tokens"] = torch.randn(1, 1, hidden, generator=g) * 0.05
    sd["pos_embedding"] = sinusoid_table(197, hidden)
    init_encoder(sd, g, hidden, [num_heads] * depth, [int(mlp_ratio * hidden)] * depth, hidden // num_heads, stress=stress)
    sd["norm.gamma"], sd["norm.beta"] = vec(hidden, 1.0), vec(hidden, 0.0)
    sd["classifier_head.kernel"], sd["classifier_head.bias"] = glorot(g, hidden, num_classes), vec(num_classes, 0.0)
    return sd
