"""Oracle: the timm / facebookresearch-deit ``VisionTransformer`` forward restated in torch (test infrastructure).

The reference loads these models with ``torch.hub.load('facebookresearch/deit:main', 'deit_{tiny,small,base}_patch16_224')``
(``utils.py:52-62`` ``get_torch_deit``; used by ``tools.py:244-263`` ``export_onnx_deit`` and ``eval_deit``).  Neither
``timm`` nor the hub repo is in the build image and there is no network, so the model code itself cannot run here:
this file restates the published ``timm.models.vision_transformer.VisionTransformer`` (timm 0.3.2, the version
facebookresearch/deit pins) --

    x = patch_embed(x)                                   # Conv2d(3, D, 16, 16) -> flatten(2).transpose(1, 2)
    x = cat(cls_token.expand(B, -1, -1), x) + pos_embed
    for blk: x = x + blk.attn(blk.norm1(x)); x = x + blk.mlp(blk.norm2(x))      # LayerNorm eps = 1e-6
    return head(norm(x)[:, 0])
    Attention: qkv = qkv(x).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4); attn = softmax(q k^T * hd^-0.5); proj(attn v)
    Mlp: fc2(GELU_erf(fc1(x)))

PINNING: the same function is HF ``ViTForImageClassification`` with ``layer_norm_eps=1e-6`` and the three q/k/v Linears
concatenated; ``tests/test_oracle.py`` pins this restatement against the installed HF forward through ``hf_to_timm``
(max-abs 1e-5), and ``tests/golden/timm_tiny_s5.npz`` holds logits of that HF forward.  Against timm itself: parity
unpinned (cannot be imported here).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

TIMM_EPS = 1e-6


def hf_to_timm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """HF ViT state dict -> timm VisionTransformer key names (fused qkv, rows ordered q | k | v)."""
    out = {
        "cls_token": sd["vit.embeddings.cls_token"],
        "pos_embed": sd["vit.embeddings.position_embeddings"],
        "patch_embed.proj.weight": sd["vit.embeddings.patch_embeddings.projection.weight"],
        "patch_embed.proj.bias": sd["vit.embeddings.patch_embeddings.projection.bias"],
        "norm.weight": sd["vit.layernorm.weight"], "norm.bias": sd["vit.layernorm.bias"],
        "head.weight": sd["classifier.weight"], "head.bias": sd["classifier.bias"],
    }
    l = 0
    while f"vit.encoder.layer.{l}.attention.attention.query.weight" in sd:
        p, q = f"vit.encoder.layer.{l}.", f"blocks.{l}."
        out[q + "attn.qkv.weight"] = torch.cat([sd[p + f"attention.attention.{n}.weight"] for n in ("query", "key", "value")], 0)
        out[q + "attn.qkv.bias"] = torch.cat([sd[p + f"attention.attention.{n}.bias"] for n in ("query", "key", "value")], 0)
        out[q + "attn.proj.weight"], out[q + "attn.proj.bias"] = sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]
        out[q + "norm1.weight"], out[q + "norm1.bias"] = sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"]
        out[q + "norm2.weight"], out[q + "norm2.bias"] = sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"]
        out[q + "mlp.fc1.weight"], out[q + "mlp.fc1.bias"] = sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]
        out[q + "mlp.fc2.weight"], out[q + "mlp.fc2.bias"] = sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]
        l += 1
    return {k: v.detach().clone() for k, v in out.items()}


def timm_vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_heads: int, patch: int = 16) -> torch.Tensor:
    B = x.shape[0]
    D = sd["cls_token"].shape[-1]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1) + sd["pos_embed"]
    hd = D // num_heads
    l = 0
    while f"blocks.{l}.attn.qkv.weight" in sd:
        p = f"blocks.{l}."
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], TIMM_EPS)
        N = y.shape[1]
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
        y = (attn @ qkv[2]).transpose(1, 2).reshape(B, N, D)
        t = t + F.linear(y, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], TIMM_EPS)
        y = F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        t = t + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        l += 1
    t = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], TIMM_EPS)
    return F.linear(t[:, 0], sd["head.weight"], sd["head.bias"])
