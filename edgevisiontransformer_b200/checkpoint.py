"""HF checkpoint directory loader for (pruned) ViT / DeiT classifiers -- the on-disk format on the input side
of the hot path (deit_pruning/src/train_main.py:392-396 writes it, deit_pruning/src/eval_main.py:87 and
are_16_heads/evaluate_iterative_pruned_deit.py:17 read it).

What a directory holds: ``config.json`` (hidden_size, num_hidden_layers, num_attention_heads, intermediate_size,
layer_norm_eps, hidden_act, image_size, patch_size, ``pruned_heads`` = {layer: [head, ...]}) and
``pytorch_model.bin`` or ``model.safetensors`` with HF key names.

Semantics restored here, because the installed transformers (5.x) dropped them:
  * ``config.pruned_heads``: transformers 4.7.0 re-applied it in ``init_weights`` so that a checkpoint saved
    AFTER ``model.prune_heads`` (physically smaller q/k/v/out-proj) loads; if instead the tensors are still
    full-size, the listed heads are removed here (index-select of rows / columns, original order kept).
  * nn_pruning FFN pruning leaves all-zero FC1 rows / FC2 columns in a full-size checkpoint;
    ``optimize_model(model, "dense")`` (inference_model_patcher.py:266-317) drops them at eval time.
    ``drop_zero_ffn`` does the same on the state dict (cross-zero, then drop, keep at least one unit).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Tuple

import torch

_QKV = ("query", "key", "value")


def prune_heads_(sd: Dict[str, torch.Tensor], pruned_heads: Dict[int, Iterable[int]], head_size: int,
                 n_orig: int = 0) -> None:
    """In place: remove the listed heads (numbered as in the ORIGINAL model) when they are still physically
    present in q/k/v/out-proj; a layer whose tensors were already shrunk by ``prune_heads`` is left alone."""
    for layer, heads in pruned_heads.items():
        layer = int(layer)
        p = f"vit.encoder.layer.{layer}.attention."
        qw = sd[p + "attention.query.weight"]
        n_now = qw.shape[0] // head_size
        heads = sorted(set(int(h) for h in heads))
        if not heads:
            continue
        orig = n_orig or n_now
        if n_now == orig - len(heads):
            continue            # checkpoint written after model.prune_heads(): already the small shapes
        if n_now != orig:
            raise ValueError(f"layer {layer}: {n_now} heads in the tensors, config says {orig} original heads with "
                             f"{len(heads)} pruned")
        keep = [h for h in range(n_now) if h not in heads]
        if not keep:
            keep = [0]          # "at least keep one head", inference_model_patcher.py:72-74
        idx = torch.cat([torch.arange(h * head_size, (h + 1) * head_size) for h in keep]).to(qw.device)
        for n in _QKV:
            sd[p + f"attention.{n}.weight"] = sd[p + f"attention.{n}.weight"].index_select(0, idx).contiguous()
            if p + f"attention.{n}.bias" in sd:
                sd[p + f"attention.{n}.bias"] = sd[p + f"attention.{n}.bias"].index_select(0, idx).contiguous()
        sd[p + "output.dense.weight"] = sd[p + "output.dense.weight"].index_select(1, idx).contiguous()


def drop_zero_heads_(sd: Dict[str, torch.Tensor], head_size: int) -> Dict[int, list]:
    """nn_pruning block pruning zeroes whole heads (attention_block_rows = head size): heads whose q, k AND v
    blocks are all zero contribute a constant; heads whose out-proj columns are all zero contribute nothing.
    Only the latter are removed (exactly function preserving).  Returns {layer: [removed heads]}."""
    removed = {}
    l = 0
    while f"vit.encoder.layer.{l}.attention.output.dense.weight" in sd:
        p = f"vit.encoder.layer.{l}.attention."
        wo = sd[p + "output.dense.weight"]
        n = wo.shape[1] // head_size
        dead = [h for h in range(n) if not wo[:, h * head_size:(h + 1) * head_size].any()]
        if dead and len(dead) < n:
            prune_heads_(sd, {l: dead}, head_size)
            removed[l] = dead
        l += 1
    return removed


def drop_zero_ffn_(sd: Dict[str, torch.Tensor]) -> None:
    """In place: optimize_model(model, 'dense') on the state dict."""
    l = 0
    while f"vit.encoder.layer.{l}.intermediate.dense.weight" in sd:
        k1, k2 = f"vit.encoder.layer.{l}.intermediate.dense.", f"vit.encoder.layer.{l}.output.dense."
        w1, b1, w2 = sd[k1 + "weight"], sd[k1 + "bias"], sd[k2 + "weight"]
        dead = (w1.abs().sum(1) == 0) | (w2.abs().sum(0) == 0)
        if dead.any():
            keep = (~dead).nonzero().squeeze(-1)
            if keep.numel() == 0:
                keep = torch.zeros(1, dtype=torch.long, device=w1.device)
                w1 = w1.clone()
                w1[0] = 0
                b1 = b1.clone()   # the kept unit must stay dead: zero its output column
                w2 = w2.clone()
                w2[:, 0] = 0
            sd[k1 + "weight"] = w1.index_select(0, keep).contiguous()
            sd[k1 + "bias"] = b1.index_select(0, keep).contiguous()
            sd[k2 + "weight"] = w2.index_select(1, keep).contiguous()
        l += 1


def read_state_dict(model_dir: str) -> Dict[str, torch.Tensor]:
    st = os.path.join(model_dir, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        return load_file(st)
    pt = os.path.join(model_dir, "pytorch_model.bin")
    if os.path.exists(pt):
        return torch.load(pt, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no model.safetensors or pytorch_model.bin in {model_dir}")


def load_checkpoint(model_dir: str, optimize: bool = True) -> Tuple[Dict[str, torch.Tensor], dict]:
    """-> (HF-named float state dict with per-layer shapes finalised, kwargs for config_from_state_dict)."""
    from .modeling_vit import normalise_keys
    with open(os.path.join(model_dir, "config.json")) as f:
        cfg = json.load(f)
    sd = normalise_keys(read_state_dict(model_dir))
    sd = {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}
    hidden, heads = int(cfg["hidden_size"]), int(cfg["num_attention_heads"])
    head_size = hidden // heads
    pruned = {int(k): v for k, v in (cfg.get("pruned_heads") or {}).items()}
    if pruned:
        prune_heads_(sd, pruned, head_size, n_orig=heads)
    if optimize:
        drop_zero_heads_(sd, head_size)
        drop_zero_ffn_(sd)
    size = cfg.get("image_size", 224)
    patch = cfg.get("patch_size", 16)
    kw = dict(layer_norm_eps=float(cfg.get("layer_norm_eps", 1e-12)), hidden_act=cfg.get("hidden_act", "gelu"),
              head_size=head_size, image_size=size[0] if isinstance(size, (list, tuple)) else int(size),
              patch_size=patch[0] if isinstance(patch, (list, tuple)) else int(patch))
    return sd, kw


def save_checkpoint(model_dir: str, sd: Dict[str, torch.Tensor], *, hidden_size: int, num_attention_heads: int,
                    intermediate_size: int, num_hidden_layers: int = 12, pruned_heads=None, layer_norm_eps=1e-12,
                    hidden_act="gelu", image_size=224, patch_size=16, num_labels=1000, safetensors: bool = True) -> None:
    """Write the directory layout HF ``save_pretrained`` produces (enough for this loader and for HF 4.x)."""
    os.makedirs(model_dir, exist_ok=True)
    cfg = dict(architectures=["ViTForImageClassification"], model_type="vit", hidden_size=hidden_size,
               num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
               intermediate_size=intermediate_size, layer_norm_eps=layer_norm_eps, hidden_act=hidden_act,
               image_size=image_size, patch_size=patch_size, num_channels=3, qkv_bias=True,
               pruned_heads={str(k): list(v) for k, v in (pruned_heads or {}).items()},
               id2label={str(i): f"LABEL_{i}" for i in range(num_labels)})
    with open(os.path.join(model_dir, "config.json"), "w") as f:
        json.dump(cfg, f)
    sd = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
    if safetensors:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(model_dir, "model.safetensors"))
    else:
        torch.save(sd, os.path.join(model_dir, "pytorch_model.bin"))
