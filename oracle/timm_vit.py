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
    }
    if "cls_classifier.weight" in sd:     # HF DeiTForImageClassificationWithTeacher -> facebookresearch DistilledVisionTransformer
        out["dist_token"] = sd["vit.embeddings.distillation_token"]
        out["head.weight"], out["head.bias"] = sd["cls_classifier.weight"], sd["cls_classifier.bias"]
        out["head_dist.weight"], out["head_dist.bias"] = sd["distillation_classifier.weight"], sd["distillation_classifier.bias"]
    else:
        out["head.weight"], out["head.bias"] = sd["classifier.weight"], sd["classifier.bias"]
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
    # DistilledVisionTransformer.forward_features (facebookresearch/deit models.py): cls, dist, patches
    pre = [sd["cls_token"].expand(B, -1, -1)] + ([sd["dist_token"].expand(B, -1, -1)] if "dist_token" in sd else [])
    t = torch.cat(pre + [t], dim=1) + sd["pos_embed"]
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
    if "dist_token" in sd:                # eval mode: the two heads' outputs averaged
        return (F.linear(t[:, 0], sd["head.weight"], sd["head.bias"]) + F.linear(t[:, 1], sd["head_dist.weight"], sd["head_dist.bias"])) / 2
    return F.linear(t[:, 0], sd["head.weight"], sd["head.bias"])


def build_hf_distilled(seed: int = 31, layers: int = 2):
    """Seeded random-init HF ``DeiTForImageClassificationWithTeacher`` (DeiT-Tiny width, eps 1e-6 as in timm) with biases, LayerNorm
    affine and the token / position embeddings randomised: the checker of the distilled two-head layout."""
    from transformers import DeiTConfig, DeiTForImageClassificationWithTeacher
    torch.manual_seed(seed)
    hf = DeiTForImageClassificationWithTeacher(DeiTConfig(hidden_size=192, num_hidden_layers=layers, num_attention_heads=3,
                                                         intermediate_size=768, num_labels=1000, layer_norm_eps=1e-6,
                                                         attn_implementation="eager")).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias") or "layernorm" in n:
                p.add_(torch.randn_like(p) * 0.1)
        for t in (hf.deit.embeddings.cls_token, hf.deit.embeddings.distillation_token, hf.deit.embeddings.position_embeddings):
            t.normal_(0, 0.02)
    return hf
