"""Oracle: ``modeling/torch_layers`` op-level models restated functionally (test infrastructure).

Follows ``modeling/torch_layers/attention.py:4-48``, ``ffn.py:7-17``, ``norm.py:4-15``,
``residual.py:4-10``, ``activation.py:4-7`` as composed by ``utils.get_attention`` /
``utils.get_ffn(is_tf=False)`` (``utils.py:322-365``):  ``LN_[n,h](x + sub(x))`` with the
LayerNorm taken jointly over the last TWO dims.  Pinned against the reference modules
imported unchanged (``tests/golden/torch_layers_*.npz``).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .vit import gelu_tanh


def attention_fwd(p: Dict[str, torch.Tensor], x: torch.Tensor, num_heads: int, head_size: int) -> torch.Tensor:
    """Attention.forward, modeling/torch_layers/attention.py:29-48 (prefix-free param names)."""
    B, n, _ = x.shape

    def heads(t):  # transpose_for_scores :24-27
        return t.view(B, n, num_heads, head_size).permute(0, 2, 1, 3)

    q = heads(F.linear(x, p["to_query.weight"], p["to_query.bias"]))
    k = heads(F.linear(x, p["to_key.weight"], p["to_key.bias"]))
    v = heads(F.linear(x, p["to_value.weight"], p["to_value.bias"]))
    s = torch.matmul(q, k.transpose(-1, -2)) * (head_size ** -0.5)
    ctx = torch.matmul(torch.softmax(s, dim=-1), v).permute(0, 2, 1, 3).contiguous().view(B, n, num_heads * head_size)
    return F.linear(ctx, p["to_out.weight"], p["to_out.bias"])


def ffn_fwd(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """FeedForward.forward, modeling/torch_layers/ffn.py:13-17."""
    return F.linear(gelu_tanh(F.linear(x, p["linear1.weight"], p["linear1.bias"])), p["linear2.weight"], p["linear2.bias"])


def post_ln_residual(sub_out: torch.Tensor, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5):
    """LayerNorm([n,h], Residual(sub), is_pre=False): norm.py:11-15 + residual.py:9-10."""
    return F.layer_norm(x + sub_out, tuple(w.shape), w, b, eps)


def strip(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@torch.no_grad()
def get_attention_fwd(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_heads: int, head_size: int):
    """Forward of utils.get_attention(h, a, h_k, is_tf=False, n) given its state_dict."""
    sub = attention_fwd(strip(sd, "sub_layer.sub_layer."), x, num_heads, head_size)
    return post_ln_residual(sub, x, sd["layer_norm.weight"], sd["layer_norm.bias"])


@torch.no_grad()
def get_ffn_fwd(sd: Dict[str, torch.Tensor], x: torch.Tensor):
    """Forward of utils.get_ffn(h, i, is_tf=False, n) given its state_dict."""
    sub = ffn_fwd(strip(sd, "sub_layer.sub_layer."), x)
    return post_ln_residual(sub, x, sd["layer_norm.weight"], sd["layer_norm.bias"])
