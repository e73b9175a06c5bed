"""Second restatement of the reference's TensorFlow-dialect DeiT in NumPy float64 (test infrastructure).

PARITY UNPINNED like ``oracle/tf_vit.py`` (TensorFlow is not installed, ``modeling/models/vit.py`` cannot run here).
What this file adds: every tensor re-layout goes through ``einops.rearrange`` with the reference's OWN pattern strings,
copied verbatim -- ``einops`` is the library the reference calls (``einops.layers.tensorflow.Rearrange``) and its pattern
semantics do not depend on the backend, so the three places where a reading error would silently permute data (patch
pixel order, the fused ``(qkv h d)`` column order, the head merge) are pinned to the real implementation; the arithmetic
around them is written from the reference's text with no code shared with ``oracle/tf_vit.py``.
``tests/test_oracle.py`` requires the two restatements to agree.

* ``ViT.call``                       ``modeling/models/vit.py:33-34, 41-55``
* ``ViT_Pruned``                     ``modeling/models/vit.py:58-73`` (per-layer heads / FFN widths, head_size 64)
* ``LayerNorm(fn, pre=True)``        ``modeling/layers/norm.py:10-12``   -> ``fn(norm(x))``, epsilon 1e-5
* ``Residual(fn)``                   ``modeling/layers/residual.py:8-9`` -> ``fn(x) + x``  (x is ALREADY normalised here)
* ``Attention.call``                 ``modeling/layers/attention.py:19-36``
* ``FeedForward`` / ``gelu``         ``modeling/layers/ffn.py:8-12``, ``modeling/layers/activation.py:4-15``
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
from einops import rearrange

PATCHES = 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'      # modeling/models/vit.py:33-34
SPLIT_QKV = 'b n (qkv h d) -> qkv b h n d'              # modeling/layers/attention.py:19
MERGE_HEADS = 'b h n d -> b n (h d)'                    # modeling/layers/attention.py:20


def _f64(t) -> np.ndarray:
    return np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64)


def keras_layernorm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """tf.keras.layers.LayerNormalization(epsilon=1e-5) over the last axis: biased variance, eps inside the root."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5) * gamma + beta


def gelu(x: np.ndarray) -> np.ndarray:
    """modeling/layers/activation.py:13-15."""
    cdf = 0.5 * (1.0 + np.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * np.power(x, 3))))
    return x * cdf


def attention(w: Dict[str, np.ndarray], p: str, x: np.ndarray, heads: int, h_k: int) -> np.ndarray:
    qkv = x @ w[p + ".to_qkv.kernel"]                                   # Dense(use_bias=False), kernel [in, out]
    qkv = rearrange(qkv, SPLIT_QKV, qkv=3, h=heads)
    q, k, v = qkv[0], qkv[1], qkv[2]
    dots = np.einsum('bhid,bhjd->bhij', q, k) * (h_k ** -0.5)
    dots = dots - dots.max(axis=-1, keepdims=True)                      # tf.nn.softmax is shift-invariant
    e = np.exp(dots)
    attn = e / e.sum(axis=-1, keepdims=True)
    out = rearrange(np.einsum('bhij,bhjd->bhid', attn, v), MERGE_HEADS)
    return out @ w[p + ".to_out.kernel"] + w[p + ".to_out.bias"]


def feed_forward(w: Dict[str, np.ndarray], p: str, x: np.ndarray) -> np.ndarray:
    h = gelu(x @ w[p + ".fc1.kernel"] + w[p + ".fc1.bias"])
    return h @ w[p + ".fc2.kernel"] + w[p + ".fc2.bias"]


def tf_vit_forward_np(sd, img, heads: List[int], h_k: int = 64, patch: int = 16) -> np.ndarray:
    """``ViT.call`` on an NCHW batch; ``sd`` uses the flat names documented in ``oracle/tf_vit.py``."""
    w = {k: _f64(v) for k, v in sd.items()}
    x = rearrange(_f64(img), PATCHES, p1=patch, p2=patch)               # vit.py:44
    x = x @ w["patch_to_embedding.kernel"] + w["patch_to_embedding.bias"]   # :45
    cls = np.broadcast_to(w["cls_token"], (x.shape[0], 1, x.shape[2]))  # :47-48
    x = np.concatenate((cls, x), axis=1) + w["pos_embedding"]           # :49-50
    for l, nh in enumerate(heads):                                      # Sequential of LayerNorm(Residual(.), pre=True) pairs
        p = f"layers.{l}"
        n = keras_layernorm(x, w[p + ".attn.norm.gamma"], w[p + ".attn.norm.beta"])
        x = attention(w, p + ".attn", n, nh, h_k) + n                   # Residual adds its INPUT, the normalised rows
        n = keras_layernorm(x, w[p + ".ffn.norm.gamma"], w[p + ".ffn.norm.beta"])
        x = feed_forward(w, p + ".ffn", n) + n
    x = x[:, 0]                                                         # :53, no final LayerNorm
    x = gelu(x @ w["mlp_head.0.kernel"] + w["mlp_head.0.bias"])         # :38-39
    return x @ w["mlp_head.1.kernel"] + w["mlp_head.1.bias"]
