"""Oracle: HF ViT/DeiT inference forward restated in plain torch (test infrastructure).

Third-party algorithm: ``transformers`` (PyPI), pinned ``==4.7.0`` by the
reference (``deit_pruning/requirements.txt:21``), 5.5.0 installed here.  The
restatement below follows ``SITE/models/vit/modeling_vit.py`` (SITE = the
installed transformers) function by function and is pinned against the live
HF module by ``tests/test_oracle.py`` and the fixtures in ``tests/golden``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .spec import ViTSpec


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # HF ACT2FN["gelu"], selected at SITE/models/vit/modeling_vit.py:291-294
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    # modeling/torch_layers/activation.py:4-7 == modeling/layers/activation.py:13-15
    return x * 0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def patch_embed(sd: Dict[str, torch.Tensor], x: torch.Tensor, patch: int) -> torch.Tensor:
    """SITE/models/vit/modeling_vit.py:151-167: Conv2d(3,D,k,k,stride k) then flatten(2).transpose(1,2).

    Written as the im2col GEMM it is: K index order (c, i, j)."""
    w = sd["vit.embeddings.patch_embeddings.projection.weight"]
    b = sd["vit.embeddings.patch_embeddings.projection.bias"]
    B, C, H, W = x.shape
    gh, gw = H // patch, W // patch
    cols = x.reshape(B, C, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * patch * patch)
    return cols @ w.reshape(w.shape[0], -1).t() + b


def embeddings(sd: Dict[str, torch.Tensor], x: torch.Tensor, spec: ViTSpec) -> torch.Tensor:
    """SITE/models/vit/modeling_vit.py:100-128 (and modeling_deit.py:108-128 for 198 tokens)."""
    h = patch_embed(sd, x, spec.patch)
    B = h.shape[0]
    toks = [sd["vit.embeddings.cls_token"].expand(B, -1, -1)]
    if spec.tokens == spec.patches + 2:
        toks.append(sd["vit.embeddings.distillation_token"].expand(B, -1, -1))
    h = torch.cat(toks + [h], dim=1)
    return h + sd["vit.embeddings.position_embeddings"]


def attention(q, k, v, head_size: int, head_mask: Optional[torch.Tensor] = None, return_heads: bool = False):
    """eager_attention_forward, SITE/models/vit/modeling_vit.py:171-196.

    head_mask [heads]: `attention_probs = attention_probs * head_mask` of the pinned transformers 4.7.0
    (ViTSelfAttention.forward; the argument was dropped from later releases) -- what are_16_heads' `mask_heads`
    (are_16_heads/run_classifier.py:247-250) sets to 0 for masked heads.  return_heads: also return the per-head context
    [B, heads, S, head_size], the `context_layer_val` of are_16_heads/classifier_eval.py:183-191."""
    B, S, A = q.shape
    nh = A // head_size

    def split(t):
        return t.view(B, S, nh, head_size).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    scores = torch.matmul(q, k.transpose(-1, -2)) * (head_size ** -0.5)
    probs = torch.softmax(scores, dim=-1)
    if head_mask is not None:
        probs = probs * head_mask.to(probs.dtype).view(1, nh, 1, 1)
    ctx = torch.matmul(probs, v)
    out = ctx.transpose(1, 2).reshape(B, S, A)
    return (out, ctx) if return_heads else out


def encoder_layer(sd: Dict[str, torch.Tensor], l: int, x: torch.Tensor, spec: ViTSpec,
                  head_mask: Optional[torch.Tensor] = None, ctx_out: Optional[list] = None) -> torch.Tensor:
    """ViTLayer.forward, SITE/models/vit/modeling_vit.py:328-346 (pre-LN)."""
    p = f"vit.encoder.layer.{l}."
    D = spec.hidden
    act = gelu_erf if spec.gelu == "erf" else gelu_tanh
    y = F.layer_norm(x, (D,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], spec.eps)
    q = F.linear(y, sd[p + "attention.attention.query.weight"], sd[p + "attention.attention.query.bias"])
    k = F.linear(y, sd[p + "attention.attention.key.weight"], sd[p + "attention.attention.key.bias"])
    v = F.linear(y, sd[p + "attention.attention.value.weight"], sd[p + "attention.attention.value.bias"])
    ctx, heads_ctx = attention(q, k, v, spec.head_size, head_mask, return_heads=True)
    if ctx_out is not None:
        ctx_out.append(heads_ctx)
    x = x + F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    y = F.layer_norm(x, (D,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], spec.eps)
    h = act(F.linear(y, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    return x + F.linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])


@torch.no_grad()
def vit_forward(sd: Dict[str, torch.Tensor], spec: ViTSpec, pixel_values: torch.Tensor,
                return_hidden: bool = False, head_mask: Optional[torch.Tensor] = None, ctx_out: Optional[list] = None):
    """ViTForImageClassification.forward, SITE/models/vit/modeling_vit.py:620-653 -> logits [B, num_labels].

    head_mask [layers, >= max heads] (row l applies to the heads of layer l); ctx_out: a list that receives every layer's
    per-head context [B, heads_l, S, head_size]."""
    x = embeddings(sd, pixel_values.to(torch.float32), spec)
    hidden = [x]
    for l in range(spec.layers):
        hm = head_mask[l, :spec.heads[l]] if head_mask is not None else None
        x = encoder_layer(sd, l, x, spec, hm, ctx_out)
        hidden.append(x)
    x = F.layer_norm(x, (spec.hidden,), sd["vit.layernorm.weight"], sd["vit.layernorm.bias"], spec.eps)
    logits = F.linear(x[:, 0, :], sd["classifier.weight"], sd["classifier.bias"])
    if return_hidden:
        return logits, hidden
    return logits


# --------------------------------------------------------------------------------------
# Seeded synthetic weights and inputs (SURVEY.md section 8d "Value distributions")
# --------------------------------------------------------------------------------------

def build_hf_model(spec: ViTSpec, seed: int = 0, stress: bool = False):
    """Random-init HF ``ViTForImageClassification`` for an UNPRUNED spec (uniform heads / inter).

    ``stress=True`` additionally draws biases and LN affine from N(0, 0.1) because HF's
    ``_init_weights`` leaves them at 0 / 1 (SITE/models/vit/modeling_vit.py:385-398),
    which would not exercise the bias / affine code paths."""
    from transformers import ViTConfig, ViTForImageClassification

    assert len(set(spec.heads)) == 1 and len(set(spec.inter)) == 1, "build_hf_model wants an unpruned spec"
    assert spec.tokens == spec.patches + 1
    cfg = ViTConfig(hidden_size=spec.hidden, num_hidden_layers=spec.layers, num_attention_heads=spec.heads[0],
                    intermediate_size=spec.inter[0], num_labels=spec.num_labels, image_size=spec.image,
                    patch_size=spec.patch, layer_norm_eps=spec.eps,
                    hidden_act="gelu" if spec.gelu == "erf" else "gelu_new",
                    attn_implementation="eager")
    torch.manual_seed(seed)
    model = ViTForImageClassification(cfg).eval()
    if stress:
        g = torch.Generator().manual_seed(seed + 1000)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("bias") or "layernorm" in name:
                    noise = torch.randn(p.shape, generator=g) * 0.1
                    if name.endswith("weight"):      # LN gamma around 1
                        p.copy_(1.0 + noise)
                    else:
                        p.copy_(noise)
                if name.endswith("cls_token") or name.endswith("position_embeddings"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model


def synthetic_images(batch: int, seed: int = 1, size: int = 224, channels_last: bool = False) -> torch.Tensor:
    """randn inputs, mirroring utils.py:162,888 and tools.py:204."""
    g = torch.Generator().manual_seed(seed)
    if channels_last:
        return torch.randn(batch, size, size, 3, generator=g)
    return torch.randn(batch, 3, size, size, generator=g)


def spec_from_state_dict(sd: Dict[str, torch.Tensor], eps: float = 1e-12, gelu: str = "erf",
                         head_size: int = 64, patch: int = 16, image: int = 224) -> ViTSpec:
    """Read per-layer shapes off HF-named weights (Appendix B of SURVEY.md)."""
    D = sd["vit.embeddings.cls_token"].shape[-1]
    L = 0
    while f"vit.encoder.layer.{L}.attention.attention.query.weight" in sd:
        L += 1
    heads = [sd[f"vit.encoder.layer.{l}.attention.attention.query.weight"].shape[0] // head_size for l in range(L)]
    inter = [sd[f"vit.encoder.layer.{l}.intermediate.dense.weight"].shape[0] for l in range(L)]
    return ViTSpec(hidden=D, layers=L, heads=heads, inter=inter, head_size=head_size,
                   tokens=sd["vit.embeddings.position_embeddings"].shape[1], eps=eps, gelu=gelu,
                   num_labels=sd["classifier.weight"].shape[0], image=image, patch=patch)


def state_dict_of(model) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone().float() for k, v in model.state_dict().items()}


def compare_logits(got: torch.Tensor, want: torch.Tensor) -> Dict[str, float]:
    """max-abs logit difference and top-1 agreement.  A row whose argmax differs still agrees when the reference itself holds the
    two classes within twice that row's max-abs difference: random-init models have near-tied top logits, and an implementation
    whose logits are within e of the reference cannot be asked to break a tie narrower than 2e the same way (with split-K
    reduce-adds landing in any order at small batch, such a tie flipped in about one run out of four)."""
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    err = (got - want).abs()
    pick = got.argmax(-1)
    margin = want.max(-1).values - want.gather(-1, pick[..., None])[..., 0]      # 0 where the argmax matches
    agree = margin <= 2 * err.amax(-1)
    return {
        "max_abs": float(err.max()),
        "top1_agree": float(agree.float().mean()),
        "top1_exact": float((pick == want.argmax(-1)).float().mean()),
    }
