"""Weight-layout adapters: the reference's TensorFlow-dialect models -> the canonical (HF-named,
``[out, in]`` Linear) layout libevt loads.

TF dialect facts handled here (modeling/models/vit.py:9-55, modeling/layers/attention.py:17-20):
  * Keras ``Dense`` kernels are ``[in, out]``                    -> transposed
  * fused no-bias ``to_qkv`` with columns ordered (qkv, head, d)   -> split into query / key / value
  * patch pixels flattened ``(p1 p2 c)``                           -> rows permuted to ``(c p1 p2)`` (our im2col order)
  * no final LayerNorm, head = Dense(mlp_dim, gelu) -> Dense(classes) -> pre_classifier + classifier
The different skip-connection semantics (skip carries LN(x)) is a runtime flag (dialect="tf").

timm / facebookresearch-deit dialect (``torch.hub.load('facebookresearch/deit:main', 'deit_*_patch16_224')``,
utils.py:52-62, tools.py:244-263): same function as HF ViT with a fused ``attn.qkv`` Linear (rows q | k | v, head-major
inside each), LayerNorm eps 1e-6 and flat key names -> pure renaming + split, see ``timm_vit_to_canonical``.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def _encoder_to_canonical(sd: Dict[str, torch.Tensor], out: Dict[str, torch.Tensor], heads: List[int], head_size: int,
                          prefix: str = "layers") -> None:
    for l, nh in enumerate(heads):
        p, q = f"{prefix}.{l}", f"vit.encoder.layer.{l}"
        a = nh * head_size
        wqkv = sd[p + ".attn.to_qkv.kernel"].t().contiguous()          # [3a, D], rows (qkv, head, d)
        for i, n in enumerate(("query", "key", "value")):
            out[f"{q}.attention.attention.{n}.weight"] = wqkv[i * a:(i + 1) * a].contiguous()
        out[f"{q}.attention.output.dense.weight"] = sd[p + ".attn.to_out.kernel"].t().contiguous()
        out[f"{q}.attention.output.dense.bias"] = sd[p + ".attn.to_out.bias"]
        out[f"{q}.layernorm_before.weight"] = sd[p + ".attn.norm.gamma"]
        out[f"{q}.layernorm_before.bias"] = sd[p + ".attn.norm.beta"]
        out[f"{q}.layernorm_after.weight"] = sd[p + ".ffn.norm.gamma"]
        out[f"{q}.layernorm_after.bias"] = sd[p + ".ffn.norm.beta"]
        out[f"{q}.intermediate.dense.weight"] = sd[p + ".ffn.fc1.kernel"].t().contiguous()
        out[f"{q}.intermediate.dense.bias"] = sd[p + ".ffn.fc1.bias"]
        out[f"{q}.output.dense.weight"] = sd[p + ".ffn.fc2.kernel"].t().contiguous()
        out[f"{q}.output.dense.bias"] = sd[p + ".ffn.fc2.bias"]


def tf_vit_to_canonical(sd: Dict[str, torch.Tensor], heads: List[int], head_size: int = 64, patch: int = 16
                        ) -> Tuple[Dict[str, torch.Tensor], dict]:
    """modeling/models/vit.py ViT / ViT_Pruned weights -> (canonical state dict, from_state_dict kwargs)."""
    out: Dict[str, torch.Tensor] = {}
    D = sd["cls_token"].shape[-1]
    k = sd["patch_to_embedding.kernel"]                                  # [(p1 p2 c), D]
    w = k.t().reshape(D, patch, patch, 3).permute(0, 3, 1, 2).contiguous()   # [D, c, p1, p2]
    out["vit.embeddings.patch_embeddings.projection.weight"] = w
    out["vit.embeddings.patch_embeddings.projection.bias"] = sd["patch_to_embedding.bias"]
    out["vit.embeddings.cls_token"] = sd["cls_token"].reshape(1, 1, D)
    out["vit.embeddings.position_embeddings"] = sd["pos_embedding"].reshape(1, -1, D)
    _encoder_to_canonical(sd, out, heads, head_size)
    out["pre_classifier.weight"] = sd["mlp_head.0.kernel"].t().contiguous()
    out["pre_classifier.bias"] = sd["mlp_head.0.bias"]
    out["classifier.weight"] = sd["mlp_head.1.kernel"].t().contiguous()
    out["classifier.bias"] = sd["mlp_head.1.bias"]
    kw = dict(dialect="tf", hidden_act="gelu_tanh", layer_norm_eps=1e-5, final_ln=False, head_size=head_size, patch_size=patch)
    return out, kw


def timm_vit_to_canonical(sd: Dict[str, torch.Tensor], patch: int = 16) -> Tuple[Dict[str, torch.Tensor], dict]:
    """timm ``VisionTransformer`` / facebookresearch DeiT state dict -> (canonical state dict, from_state_dict kwargs).

    Accepts the hub checkpoints' ``{'model': state_dict}`` wrapper.  The reference asks the hub for
    ``deit_{type}_patch16_{224,384}`` (utils.py:52-62); the hub's distilled variants (``DistilledVisionTransformer``:
    ``dist_token``, ``head_dist``, eval output ``(head(x[:, 0]) + head_dist(x[:, 1])) / 2``) map onto the two-head layout that
    ``modeling_vit.normalise_keys`` folds into one classifier.  Half of that layout is refused."""
    if "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    distilled = "dist_token" in sd
    if distilled != ("head_dist.weight" in sd):
        raise ValueError("distilled DeiT checkpoint needs both dist_token and head_dist")
    D = sd["cls_token"].shape[-1]
    out: Dict[str, torch.Tensor] = {
        "vit.embeddings.cls_token": sd["cls_token"].reshape(1, 1, D),
        "vit.embeddings.position_embeddings": sd["pos_embed"].reshape(1, -1, D),
        "vit.embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
        "vit.embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
        "vit.layernorm.weight": sd["norm.weight"], "vit.layernorm.bias": sd["norm.bias"],
    }
    if distilled:
        out["vit.embeddings.distillation_token"] = sd["dist_token"].reshape(1, 1, D)
        out["cls_classifier.weight"], out["cls_classifier.bias"] = sd["head.weight"], sd["head.bias"]
        out["distillation_classifier.weight"], out["distillation_classifier.bias"] = sd["head_dist.weight"], sd["head_dist.bias"]
    else:
        out["classifier.weight"], out["classifier.bias"] = sd["head.weight"], sd["head.bias"]
    l = 0
    while f"blocks.{l}.attn.qkv.weight" in sd:
        p, q = f"blocks.{l}.", f"vit.encoder.layer.{l}."
        w = sd[p + "attn.qkv.weight"]
        if w.shape[0] != 3 * D:
            raise ValueError(f"blocks.{l}.attn.qkv.weight has {w.shape[0]} rows, expected 3 x {D}")
        b = sd.get(p + "attn.qkv.bias")
        for i, n in enumerate(("query", "key", "value")):
            out[q + f"attention.attention.{n}.weight"] = w[i * D:(i + 1) * D].contiguous()
            if b is not None:
                out[q + f"attention.attention.{n}.bias"] = b[i * D:(i + 1) * D].contiguous()
        for src, dst in (("attn.proj", "attention.output.dense"), ("mlp.fc1", "intermediate.dense"), ("mlp.fc2", "output.dense"),
                         ("norm1", "layernorm_before"), ("norm2", "layernorm_after")):
            out[q + dst + ".weight"] = sd[p + src + ".weight"]
            out[q + dst + ".bias"] = sd[p + src + ".bias"]
        l += 1
    if l == 0:
        raise ValueError("not a timm VisionTransformer state dict (no blocks.0.attn.qkv.weight)")
    return out, dict(layer_norm_eps=1e-6, hidden_act="gelu", patch_size=patch)
