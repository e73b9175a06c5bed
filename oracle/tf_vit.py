"""Oracle: the reference's TensorFlow-dialect DeiT restated in torch (test infrastructure).

PARITY UNPINNED: TensorFlow is not installed in the build image, so the reference's own
``modeling/models/vit.py`` cannot run here and it ships no fixture.  This file restates it
line by line:

* ``ViT.call``                         ``modeling/models/vit.py:41-55``
* ``ViT_Pruned``                       ``modeling/models/vit.py:58-97``
* ``TransformerEncoderBlock[_Pruned]`` ``modeling/layers/transformer_encoder.py:9-36``
* ``LayerNorm(fn, pre)``               ``modeling/layers/norm.py:3-14``  (eps 1e-5)
* ``Residual``                         ``modeling/layers/residual.py:3-9``
* ``Attention``                        ``modeling/layers/attention.py:5-36`` (fused no-bias qkv,
                                        column order ``(qkv, head, d)``)
* ``FeedForward`` / ``gelu``           ``modeling/layers/ffn.py:5-12``, ``activation.py:4-15`` (tanh)

Dialect facts that differ from HF:  ``LayerNorm(Residual(f), pre=True)`` evaluates
``f(LN(x)) + LN(x)`` -- the skip connection carries the NORMALISED activations; patch pixels are
flattened ``(p1 p2 c)``; Keras ``Dense`` kernels are ``[in, out]``; no final LayerNorm; the head is
``Dense(mlp_dim, gelu) -> Dense(num_classes)``.

Weight names are this repo's own flat scheme (Keras checkpoints have no stable names):
``pos_embedding [S,D]``, ``cls_token [1,1,D]``, ``patch_to_embedding.{kernel,bias}``,
``layers.{l}.attn.norm.{gamma,beta}``, ``layers.{l}.attn.to_qkv.kernel``, ``layers.{l}.attn.to_out.{kernel,bias}``,
``layers.{l}.ffn.norm.{gamma,beta}``, ``layers.{l}.ffn.fc{1,2}.{kernel,bias}``, ``mlp_head.{0,1}.{kernel,bias}``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .vit import gelu_tanh

TF_EPS = 1e-5


def dense(x, sd, name, bias=True):
    y = x @ sd[name + ".kernel"]
    return y + sd[name + ".bias"] if bias else y


def tf_attention(sd, prefix, x, num_heads: int, h_k: int):
    """modeling/layers/attention.py:23-36."""
    B, n, _ = x.shape
    qkv = dense(x, sd, prefix + ".to_qkv", bias=False)                       # :24
    qkv = qkv.view(B, n, 3, num_heads, h_k).permute(2, 0, 3, 1, 4)            # 'b n (qkv h d) -> qkv b h n d' :20
    q, k, v = qkv[0], qkv[1], qkv[2]
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * (h_k ** -0.5)             # :30
    attn = torch.softmax(dots, dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, n, num_heads * h_k)             # 'b h n d -> b n (h d)'
    return dense(out, sd, prefix + ".to_out")


def tf_ffn(sd, prefix, x):
    return dense(gelu_tanh(dense(x, sd, prefix + ".fc1")), sd, prefix + ".fc2")


def tf_encoder(sd, x, heads: List[int], h_k: int, prefix: str = "layers"):
    """TransformerEncoderBlock(norm_first=True): each sub-block is fn(LN(x)) with fn = Residual(f)."""
    D = x.shape[-1]
    for l, nh in enumerate(heads):
        p = f"{prefix}.{l}"
        y = F.layer_norm(x, (D,), sd[p + ".attn.norm.gamma"], sd[p + ".attn.norm.beta"], TF_EPS)
        x = tf_attention(sd, p + ".attn", y, nh, h_k) + y
        y = F.layer_norm(x, (D,), sd[p + ".ffn.norm.gamma"], sd[p + ".ffn.norm.beta"], TF_EPS)
        x = tf_ffn(sd, p + ".ffn", y) + y
    return x


@torch.no_grad()
def tf_vit_forward(sd: Dict[str, torch.Tensor], img: torch.Tensor, heads: List[int], h_k: int = 64, patch: int = 16):
    """ViT.call (modeling/models/vit.py:41-55); img is NCHW as the Rearrange pattern demands."""
    B, C, H, W = img.shape
    gh, gw = H // patch, W // patch
    # 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'
    x = img.reshape(B, C, gh, patch, gw, patch).permute(0, 2, 4, 3, 5, 1).reshape(B, gh * gw, patch * patch * C)
    x = dense(x, sd, "patch_to_embedding")
    x = torch.cat((sd["cls_token"].expand(B, 1, -1), x), dim=1)
    x = x + sd["pos_embedding"]
    x = tf_encoder(sd, x, heads, h_k)
    x = x[:, 0]
    x = gelu_tanh(dense(x, sd, "mlp_head.0"))
    return dense(x, sd, "mlp_head.1")


# ------------------------------------------------------------------ seeded init (Keras defaults)

def glorot(g, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * lim


def init_encoder(sd, g, D, heads, inter, h_k, prefix="layers", stress=False):
    def vec(n, base):
        return base + (torch.randn(n, generator=g) * 0.1 if stress else torch.zeros(n))
    for l, (nh, i) in enumerate(zip(heads, inter)):
        p = f"{prefix}.{l}"
        a = nh * h_k
        sd[p + ".attn.norm.gamma"], sd[p + ".attn.norm.beta"] = vec(D, 1.0), vec(D, 0.0)
        sd[p + ".attn.to_qkv.kernel"] = glorot(g, D, 3 * a)
        sd[p + ".attn.to_out.kernel"], sd[p + ".attn.to_out.bias"] = glorot(g, a, D), vec(D, 0.0)
        sd[p + ".ffn.norm.gamma"], sd[p + ".ffn.norm.beta"] = vec(D, 1.0), vec(D, 0.0)
        sd[p + ".ffn.fc1.kernel"], sd[p + ".ffn.fc1.bias"] = glorot(g, D, i), vec(i, 0.0)
        sd[p + ".ffn.fc2.kernel"], sd[p + ".ffn.fc2.bias"] = glorot(g, i, D), vec(D, 0.0)


def init_tf_vit(dim=192, depth=12, heads: Optional[List[int]] = None, inter: Optional[List[int]] = None,
                mlp_dim: Optional[int] = None, h_k=64, num_classes=1000, patch=16, image=224, seed=0, stress=False):
    """Random weights with Keras' default initialisers (glorot-uniform kernels, zero biases,
    RandomNormal(0.05) cls/pos).  get_deit_{tiny,small,base}: modeling/models/vit.py:100-109."""
    g = torch.Generator().manual_seed(seed)
    mlp_dim = mlp_dim or 4 * dim
    heads = heads or [dim // h_k] * depth
    inter = inter or [mlp_dim] * depth
    S = (image // patch) ** 2 + 1
    sd = {}
    sd["pos_embedding"] = torch.randn(S, dim, generator=g) * 0.05
    sd["cls_token"] = torch.randn(1, 1, dim, generator=g) * 0.05
    sd["patch_to_embedding.kernel"] = glorot(g, patch * patch * 3, dim)
    sd["patch_to_embedding.bias"] = torch.randn(dim, generator=g) * 0.1 if stress else torch.zeros(dim)
    init_encoder(sd, g, dim, heads, inter, h_k, stress=stress)
    sd["mlp_head.0.kernel"] = glorot(g, dim, mlp_dim)
    sd["mlp_head.0.bias"] = torch.randn(mlp_dim, generator=g) * 0.1 if stress else torch.zeros(mlp_dim)
    sd["mlp_head.1.kernel"] = glorot(g, mlp_dim, num_classes)
    sd["mlp_head.1.bias"] = torch.randn(num_classes, generator=g) * 0.1 if stress else torch.zeros(num_classes)
    return sd, heads, inter
