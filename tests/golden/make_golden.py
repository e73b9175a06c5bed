"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports, unchanged,
  * /root/reference/modeling/torch_layers + utils.get_attention/get_ffn(is_tf=False)   (utils.py:322-365)
  * /root/reference/deit_pruning/vendor/nn_pruning_v1 : nn_pruning.inference_model_patcher.optimize_model
  * the installed transformers ViTForImageClassification (the third-party forward the reference calls,
    deit_pruning/src/utils.py:194-195)
and stores small input/output vectors.  The fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import ViTSpec  # noqa: E402
from oracle import pruning as opr  # noqa: E402
from oracle import vit as ovit  # noqa: E402


def np_sd(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def gen_torch_layers():
    """Op-level models exactly as tools.py export_onnx_attention/ffn builds them (tools.py:735-787)."""
    sys.path.insert(0, REF)
    import importlib.util
    # utils.py imports heavy optional deps lazily inside functions; load only what we need
    from modeling.torch_layers.attention import Attention
    from modeling.torch_layers.ffn import FeedForward
    from modeling.torch_layers.norm import LayerNorm
    from modeling.torch_layers.residual import Residual

    cases = [
        ("attn_h192_a3_n197", dict(h=192, a=3, h_k=None, n=197)),
        ("attn_h128_a1_hk64_n128", dict(h=128, a=1, h_k=64, n=128)),   # head size decoupled from h/a
        ("attn_h384_a6_n40", dict(h=384, a=6, h_k=None, n=40)),
    ]
    for name, c in cases:
        torch.manual_seed(11)
        m = LayerNorm([c["n"], c["h"]], Residual(Attention(c["h"], c["a"], c["h_k"])))  # utils.py:338
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        m.eval()
        x = torch.randn(2, c["n"], c["h"])
        with torch.no_grad():
            y = m(x)
        hs = m.sub_layer.sub_layer.head_size
        np.savez_compressed(os.path.join(HERE, f"torch_layers_{name}.npz"), x=x.numpy(), y=y.numpy(),
                            num_heads=c["a"], head_size=hs, **{"sd." + k: v for k, v in np_sd(m.state_dict()).items()})
    for name, c in [("ffn_h192_i768_n197", dict(h=192, i=768, n=197)), ("ffn_h64_i230_n33", dict(h=64, i=230, n=33))]:
        torch.manual_seed(12)
        m = LayerNorm([c["n"], c["h"]], Residual(FeedForward(c["h"], c["i"])))          # utils.py:364
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        m.eval()
        x = torch.randn(2, c["n"], c["h"])
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(HERE, f"torch_layers_{name}.npz"), x=x.numpy(), y=y.numpy(),
                            **{"sd." + k: v for k, v in np_sd(m.state_dict()).items()})
    sys.path.remove(REF)


def gen_hf():
    """HF forward on seeded random-init weights: logits + per-layer hidden-state checksums."""
    for name, spec, seed, stress, bs in [
        ("tiny_s0", ViTSpec.deit("tiny"), 0, False, 2),          # BASELINE config 1 weights (seed 0), input seed 1
        ("tiny_s3_stress", ViTSpec.deit("tiny"), 3, True, 2),
        ("small_s0_stress", ViTSpec.deit("small"), 0, True, 1),
        ("tiny_tanh_eps5", ViTSpec.deit("tiny", gelu="tanh", eps=1e-5), 5, True, 1),
    ]:
        model = ovit.build_hf_model(spec, seed=seed, stress=stress)
        x = ovit.synthetic_images(bs, seed=1)
        with torch.no_grad():
            out = model(pixel_values=x, output_hidden_states=True)
        hs = torch.stack([h.double().abs().mean() for h in out.hidden_states]).numpy()
        np.savez_compressed(os.path.join(HERE, f"hf_{name}.npz"), logits=out.logits.numpy(), hidden_absmean=hs,
                            seed=seed, stress=int(stress), batch=bs, hidden=spec.hidden, heads=spec.heads[0],
                            gelu=spec.gelu, eps=spec.eps)


def gen_timm():
    """timm / facebookresearch-deit dialect: HF forward with eps 1e-6 (the same function, see oracle/timm_vit.py)."""
    spec = ViTSpec.deit("tiny", eps=1e-6)
    model = ovit.build_hf_model(spec, seed=5, stress=True)
    x = ovit.synthetic_images(2, seed=1)
    with torch.no_grad():
        logits = model(pixel_values=x).logits
    np.savez_compressed(os.path.join(HERE, "timm_tiny_s5.npz"), logits=logits.numpy(), seed=5, batch=2)


def gen_swin():
    """HF SwinForImageClassification (the transformers port of the model utils.get_swin builds) on seeded weights."""
    from oracle import swin as osw
    for name, kind, depths, seed, bs in [("swin_tiny_s7", "tiny", None, 7, 2), ("swin_tiny_d1131_s8", "tiny", [1, 1, 3, 1], 8, 1)]:
        model = osw.build_hf_swin(kind, seed=seed, stress=True, depths=depths)
        x = ovit.synthetic_images(bs, seed=1)
        with torch.no_grad():
            logits = model(pixel_values=x).logits
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), logits=logits.numpy(), seed=seed, batch=bs,
                            depths=np.array(model.config.depths))


def gen_pruned():
    """Vendored optimize_model + head pruning applied to the live HF module."""
    sys.path.insert(0, os.path.join(REF, "deit_pruning/vendor/nn_pruning_v1"))
    from nn_pruning.inference_model_patcher import optimize_model
    from transformers.pytorch_utils import prune_linear_layer

    spec = ViTSpec.deit("tiny")
    cases = {
        # BASELINE config 4: 1 head / layer, 230 FFN rows / layer
        "tiny_h1_d230": ([[0]] * 12, [230] * 12),
        # are16heads tiny-18 head set (draw.py:104-106) with uneven FFN widths
        "tiny_head18_uneven": (opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3),
                               [768, 231, 230, 64, 700, 8, 333, 512, 1, 96, 768, 407]),
    }
    for name, (heads_kept, inter_kept) in cases.items():
        model = ovit.build_hf_model(spec, seed=4, stress=True)
        sd = ovit.state_dict_of(model)
        full, pruned, to_prune = opr.synthesize_pruned(sd, heads_kept, inter_kept, seed=7)
        model.load_state_dict(full)
        x = ovit.synthetic_images(2, seed=1)
        with torch.no_grad():
            logits_full = model(pixel_values=x).logits
        opt = optimize_model(model, "dense")                        # deit_pruning/src/eval_main.py:91-93
        # HF-4.x prune_heads semantics on the live module (SURVEY.md section 8c recipe)
        for l, dead in to_prune.items():
            att = opt.vit.encoder.layer[l].attention
            kept = [h for h in range(3) if h not in dead]
            idx = torch.cat([torch.arange(h * 64, (h + 1) * 64) for h in kept])
            att.attention.query = prune_linear_layer(att.attention.query, idx)
            att.attention.key = prune_linear_layer(att.attention.key, idx)
            att.attention.value = prune_linear_layer(att.attention.value, idx)
            att.output.dense = prune_linear_layer(att.output.dense, idx, dim=1)
            att.attention.num_attention_heads = len(kept)
            att.attention.all_head_size = 64 * len(kept)
        with torch.no_grad():
            logits_opt = opt(pixel_values=x).logits
        shapes_ref = np.array([[opt.vit.encoder.layer[l].attention.attention.query.weight.shape[0],
                                opt.vit.encoder.layer[l].intermediate.dense.weight.shape[0]] for l in range(12)])
        np.savez_compressed(os.path.join(HERE, f"pruned_{name}.npz"), logits_full=logits_full.numpy(),
                            logits_opt=logits_opt.numpy(), shapes=shapes_ref,
                            heads_kept=np.array([len(h) for h in heads_kept]), inter_kept=np.array(inter_kept))
        print(name, "full-vs-opt max abs", float((logits_full - logits_opt).abs().max()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_torch_layers()
    gen_hf()
    gen_pruned()
    gen_timm()
    gen_swin()
    print("fixtures written to", HERE)
