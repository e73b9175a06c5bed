"""Second, independently written restatement of the reference's tokens-to-token front-end (test infrastructure).

PARITY UNPINNED like ``oracle/t2t.py`` (TensorFlow is not installed).  Its purpose is to catch a SHARED misreading: it
is written from the reference's text in NumPy float64 with explicit index arithmetic -- no ``unfold``, no ``einsum``, no
code shared with ``oracle/t2t.py`` -- and ``tests/test_oracle.py`` requires both restatements to agree.

* soft split      ``modeling/models/t2t_vit.py:20-40`` (``tf.pad`` then ``tf.image.extract_patches(..., 'VALID')``: output
                  position (oy, ox) holds the k x k x C window whose top-left input pixel is (oy*s, ox*s) of the PADDED
                  image, flattened row-major over (window row, window column, channel))
* TokenPerformer  ``modeling/layers/transformer_encoder.py:39-101``
* T2T_module      ``modeling/models/t2t_vit.py:63-88``
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

EPS_LN = 1e-5       # tf.keras.layers.LayerNormalization(epsilon=1e-5), transformer_encoder.py:49-50
EPS_DIV = 1e-8      # transformer_encoder.py:51


def soft_split(x: np.ndarray, k: int, s: int, p: int) -> np.ndarray:
    """x [B,H,W,C] -> [B, oh*ow, k*k*C]; element ((oy*ow + ox), (ky*k + kx)*C + c) = padded[oy*s + ky, ox*s + kx, c]."""
    B, H, W, C = x.shape
    Hp, Wp = H + 2 * p, W + 2 * p
    padded = np.zeros((B, Hp, Wp, C), dtype=np.float64)
    padded[:, p:p + H, p:p + W, :] = x
    oh, ow = (Hp - k) // s + 1, (Wp - k) // s + 1
    out = np.empty((B, oh * ow, k * k * C), dtype=np.float64)
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(k):
                for kx in range(k):
                    d0 = (ky * k + kx) * C
                    out[:, oy * ow + ox, d0:d0 + C] = padded[:, oy * s + ky, ox * s + kx, :]
    return out


def _layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray) -> np.ndarray:
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + EPS_LN) * gamma + beta


def _gelu(x: np.ndarray) -> np.ndarray:
    """modeling/layers/activation.py:13-15 (tanh form)."""
    return 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _positive_random_features(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """prm_exp: exp(w.x - |x|^2 / 2) / sqrt(m), per token (transformer_encoder.py:67-81); x [T, e], w [m, e]."""
    m = w.shape[0]
    out = np.empty((x.shape[0], m), dtype=np.float64)
    for t in range(x.shape[0]):
        half_sq = 0.5 * float(np.dot(x[t], x[t]))
        out[t] = np.exp(w @ x[t] - half_sq) / math.sqrt(m)
    return out


def performer(sd: Dict[str, np.ndarray], p: str, x: np.ndarray) -> np.ndarray:
    """TokenPerformer.call for one image; x [T, in_dim] -> [T, emb]."""
    emb = sd[p + ".attn_output.kernel"].shape[0]
    x = _layer_norm(x, sd[p + ".norm1.gamma"], sd[p + ".norm1.beta"])
    kqv = x @ sd[p + ".kqv.kernel"] + sd[p + ".kqv.bias"]
    k, q, v = kqv[:, :emb], kqv[:, emb:2 * emb], kqv[:, 2 * emb:]           # tf.split(..., 3): k first, then q, then v
    kp, qp = _positive_random_features(k, sd[p + ".w"]), _positive_random_features(q, sd[p + ".w"])
    ksum = kp.sum(axis=0)                                                    # [m]
    kptv = v.T @ kp                                                          # [emb, m] = sum_t v[t, n] kp[t, m]
    y = np.empty((x.shape[0], emb), dtype=np.float64)
    for t in range(x.shape[0]):
        denom = float(np.dot(qp[t], ksum)) + EPS_DIV
        y[t] = (kptv @ qp[t]) / denom
    y = v + (y @ sd[p + ".attn_output.kernel"] + sd[p + ".attn_output.bias"])
    z = _layer_norm(y, sd[p + ".norm2.gamma"], sd[p + ".norm2.beta"])
    h = _gelu(z @ sd[p + ".mlp.fc1.kernel"] + sd[p + ".mlp.fc1.bias"])
    return y + (h @ sd[p + ".mlp.fc2.kernel"] + sd[p + ".mlp.fc2.bias"])


def t2t_tokens(sd_t: Dict[str, object], x_nhwc) -> np.ndarray:
    """T2T_module.call: [B,224,224,3] -> [B,196,D] (float64).  sd_t: the torch state dict of oracle/t2t.py's naming."""
    sd = {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float64) for k, v in sd_t.items()
          if k.startswith("t2t.")}
    x = np.asarray(x_nhwc.detach().cpu().numpy() if hasattr(x_nhwc, "detach") else x_nhwc, dtype=np.float64)
    B = x.shape[0]
    t = soft_split(x, 7, 4, 2)                                               # [B, 56*56, 147]
    t = np.stack([performer(sd, "t2t.performer1", t[b]) for b in range(B)])  # [B, 3136, 64]
    side = int(round(math.sqrt(t.shape[1])))
    t = soft_split(t.reshape(B, side, side, -1), 3, 2, 1)                    # [B, 28*28, 576]
    t = np.stack([performer(sd, "t2t.performer2", t[b]) for b in range(B)])  # [B, 784, 64]
    side = int(round(math.sqrt(t.shape[1])))
    t = soft_split(t.reshape(B, side, side, -1), 3, 2, 1)                    # [B, 196, 576]
    return t @ sd["t2t.project.kernel"] + sd["t2t.project.bias"]
