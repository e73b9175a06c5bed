"""CPU tests: the oracle restatements against the committed golden fixtures (generated from the
reference itself by tests/golden/make_golden.py) and against the live HF module."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ViTSpec
from oracle import pruning as opr
from oracle import t2t as ot2t
from oracle import tf_vit as otf
from oracle import torch_layers as otl
from oracle import vit as ovit


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["tiny_s0", "tiny_s3_stress", "small_s0_stress", "tiny_tanh_eps5"])
def test_vit_restatement_matches_hf_golden(golden_dir, name):
    f = _load(golden_dir, f"hf_{name}.npz")
    kind = {192: "tiny", 384: "small", 768: "base"}[int(f["hidden"])]
    spec = ViTSpec.deit(kind, gelu=str(f["gelu"]), eps=float(f["eps"]))
    model = ovit.build_hf_model(spec, seed=int(f["seed"]), stress=bool(f["stress"]))
    x = ovit.synthetic_images(int(f["batch"]), seed=1)
    sd = ovit.state_dict_of(model)
    logits, hidden = ovit.vit_forward(sd, spec, x, return_hidden=True)
    # restatement vs fixture produced by the HF forward (fp32 rounding only)
    assert np.abs(logits.numpy() - f["logits"]).max() < 5e-5
    hs = torch.stack([h.double().abs().mean() for h in hidden]).numpy()
    np.testing.assert_allclose(hs, f["hidden_absmean"], rtol=1e-5)
    # live HF vs fixture: the seeded build is reproducible in this image
    with torch.no_grad():
        live = model(pixel_values=x).logits
    assert np.abs(live.numpy() - f["logits"]).max() < 5e-5
    assert (logits.argmax(-1).numpy() == f["logits"].argmax(-1)).all()


def test_spec_from_state_dict_and_flops():
    spec = ViTSpec.deit("base")
    assert abs(spec.matmul_flops() / 1e9 - 35.128) < 0.01
    assert abs(ViTSpec.deit("tiny").matmul_flops() / 1e9 - 2.507) < 0.01
    assert abs(ViTSpec.deit("small").matmul_flops() / 1e9 - 9.198) < 0.01
    pr = ViTSpec.deit("tiny", heads=[1] * 12, inter=[230] * 12)
    assert abs(pr.matmul_flops() / 1e9 - 0.827) < 0.01
    m = ovit.build_hf_model(ViTSpec.deit("tiny"), seed=0)
    s2 = ovit.spec_from_state_dict(ovit.state_dict_of(m))
    assert s2.heads == [3] * 12 and s2.inter == [768] * 12 and s2.hidden == 192 and s2.tokens == 197


@pytest.mark.parametrize("name", ["tiny_h1_d230", "tiny_head18_uneven"])
def test_pruning_restatement_matches_vendored_optimize_model(golden_dir, name):
    f = _load(golden_dir, f"pruned_{name}.npz")
    spec = ViTSpec.deit("tiny")
    model = ovit.build_hf_model(spec, seed=4, stress=True)
    sd = ovit.state_dict_of(model)
    if name == "tiny_h1_d230":
        heads_kept = [[0]] * 12
    else:
        heads_kept = opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3)
        assert [len(h) for h in heads_kept] == [1, 1, 1, 1, 2, 1, 2, 2, 2, 2, 1, 2]
    inter_kept = [int(v) for v in f["inter_kept"]]
    full, pruned, _ = opr.synthesize_pruned(sd, heads_kept, inter_kept, seed=7)
    pspec = ovit.spec_from_state_dict(pruned)
    shapes = np.array([[h * 64, i] for h, i in zip(pspec.heads, pspec.inter)])
    np.testing.assert_array_equal(shapes, f["shapes"])              # same shapes as optimize_model + prune_heads
    x = ovit.synthetic_images(2, seed=1)
    got = ovit.vit_forward(pruned, pspec, x)
    assert np.abs(got.numpy() - f["logits_opt"]).max() < 5e-5
    got_full = ovit.vit_forward(full, spec, x)
    assert np.abs(got_full.numpy() - f["logits_full"]).max() < 5e-5


def test_pruning_dsl_parsers():
    thr = opr.parse_layerwise_thresholds("-".join(["h_0.50_d_0.3"] * 12))
    assert len(thr) == 12 and thr[0] == {"head": 0.5, "dense": 0.3}
    heads, inter = opr.parse_prune_encoding("all_head2_ffn0.5", 12, 768)
    assert heads == [2] * 12 and inter == [384] * 12
    heads, inter = opr.parse_prune_encoding("layerwise_h2-d1.0_h3-d0.5_h1-d0.5", 3, 768)
    assert heads == [2, 3, 1] and inter == [768, 384, 384]
    m = ovit.build_hf_model(ViTSpec.deit("tiny"), seed=0)
    tp = opr.heads_to_prune_from_thresholds(ovit.state_dict_of(m), thr, 3)
    assert all(len(v) == 2 for v in tp.values())                   # int(0.5*3)=1 head kept per layer


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "torch_layers_*.npz"))))
def test_torch_layers_restatement_matches_reference(path):
    f = np.load(path)
    sd = {k[3:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("sd.")}
    x = torch.from_numpy(f["x"])
    if "num_heads" in f.files:
        y = otl.get_attention_fwd(sd, x, int(f["num_heads"]), int(f["head_size"]))
    else:
        y = otl.get_ffn_fwd(sd, x)
    assert np.abs(y.numpy() - f["y"]).max() < 2e-5


def test_tf_dialect_and_t2t_shapes():
    # parity unpinned (no TensorFlow): structural checks only
    sd, heads, inter = otf.init_tf_vit(dim=192, depth=2, seed=0)
    y = otf.tf_vit_forward(sd, torch.randn(1, 3, 224, 224), heads)
    assert y.shape == (1, 1000) and torch.isfinite(y).all()
    # tf_Unfold ordering: depth is (kh, kw, c); compare with torch.nn.Unfold's (c, kh, kw)
    x = torch.randn(2, 10, 10, 3)
    u = ot2t.unfold_nhwc(x, 3, 2, 1)
    ref = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), 3, padding=1, stride=2)      # [B, c*kh*kw, L]
    ref = ref.view(2, 3, 9, -1).permute(0, 3, 2, 1).reshape(2, -1, 27)
    assert torch.equal(u, ref)
    sd = ot2t.init_t2t_vit(hidden=64, depth=1, num_heads=1, seed=0)
    logits, tok = ot2t.t2t_vit_forward(sd, torch.randn(1, 224, 224, 3), 1, 1, return_tokens=True)
    assert tok.shape == (1, 196, 64) and logits.shape == (1, 1000) and torch.isfinite(logits).all()
    w = sd["t2t.performer1.w"]
    assert torch.allclose(w @ w.t(), 32.0 * torch.eye(32), atol=1e-3)
    tab = ot2t.sinusoid_table(197, 384)
    assert tab.shape == (197, 384) and float(tab[0, 0]) == 0.0 and float(tab[0, 1]) == 1.0


def test_timm_restatement_matches_hf(golden_dir):
    """oracle.timm_vit (restated timm VisionTransformer) == the HF forward with eps 1e-6 on converted weights."""
    from oracle import timm_vit as otimm
    spec = ViTSpec.deit("tiny", eps=1e-6)
    hf = ovit.build_hf_model(spec, seed=5, stress=True)
    tsd = otimm.hf_to_timm(ovit.state_dict_of(hf))
    assert tsd["blocks.0.attn.qkv.weight"].shape == (576, 192)
    x = ovit.synthetic_images(2, seed=1)
    got = otimm.timm_vit_forward(tsd, x, num_heads=3)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    assert (got - want).abs().max() < 1e-5
    f = np.load(os.path.join(golden_dir, "timm_tiny_s5.npz"))
    assert np.abs(got.numpy() - f["logits"]).max() < 1e-5


def test_timm_adapter_roundtrip():
    """dialects.timm_vit_to_canonical is the inverse of the HF -> timm renaming (host logic, no GPU)."""
    from edgevisiontransformer_b200.dialects import timm_vit_to_canonical
    from oracle import timm_vit as otimm
    sd = ovit.state_dict_of(ovit.build_hf_model(ViTSpec.deit("tiny", eps=1e-6), seed=5, stress=True))
    csd, kw = timm_vit_to_canonical({"model": otimm.hf_to_timm(sd)})
    assert kw["layer_norm_eps"] == 1e-6 and kw["hidden_act"] == "gelu"
    assert set(csd) == set(sd)
    for k in sd:
        assert torch.equal(csd[k], sd[k]), k
    bad = otimm.hf_to_timm(sd)
    bad["dist_token"] = torch.zeros(1, 1, 192)
    with pytest.raises(ValueError):
        timm_vit_to_canonical(bad)


@pytest.mark.parametrize("name,depths,seed,bs", [("swin_tiny_s7", None, 7, 2), ("swin_tiny_d1131_s8", [1, 1, 3, 1], 8, 1)])
def test_swin_restatement_matches_hf(golden_dir, name, depths, seed, bs):
    """oracle.swin == installed HF SwinForImageClassification on the same seeded weights, and == the committed fixture."""
    from oracle import swin as osw
    hf = osw.build_hf_swin("tiny", seed=seed, stress=True, depths=depths)
    sd = ovit.state_dict_of(hf)
    x = ovit.synthetic_images(bs, seed=1)
    got = osw.swin_forward(sd, x, hf.config.depths, hf.config.num_heads)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    assert (got - want).abs().max() < 1e-4
    f = np.load(os.path.join(golden_dir, name + ".npz"))
    assert np.abs(got.numpy() - f["logits"]).max() < 1e-4


def test_swin_host_tables_and_key_adapter():
    """Host logic of modeling_swin (no GPU): window-order tables == roll + window_partition, the shift mask and the padded
    attention table == the reference construction, microsoft_to_hf inverts the renaming."""
    from edgevisiontransformer_b200 import modeling_swin as ms
    from oracle import swin as osw
    for H, sh in ((14, 0), (14, 3), (28, 3), (7, 0)):
        x = torch.arange(H * H).float().view(1, H, H, 1)
        y = torch.roll(x, (-sh, -sh), (1, 2)) if sh else x
        assert torch.equal(osw.window_partition(y, 7).view(-1).long(), ms.window_order(H, H, 7, sh))
    assert torch.equal(ms.shift_mask(28, 28, 7, 3), osw.shift_mask(28, 28, 7, 3))
    assert torch.equal(ms.relative_position_index(7), osw.relative_position_index(7))
    tab = torch.randn(169, 3)
    t = ms.attention_table(tab, 3, 7, ms.shift_mask(14, 14, 7, 3))
    assert t.shape == (4, 3, 64, 56) and torch.isinf(t[..., 49:]).all() and (t[:, :, 49:, :49] == 0).all()
    want = tab[osw.relative_position_index(7).view(-1)].view(49, 49, 3).permute(2, 0, 1)[None] + osw.shift_mask(14, 14, 7, 3)[:, None]
    assert torch.allclose(t[:, :, :49, :49], want * ms.LOG2E)
    sd = ovit.state_dict_of(osw.build_hf_swin("tiny", seed=1, depths=[1, 1, 1, 1]))
    back = ms.microsoft_to_hf({"model": osw.hf_to_microsoft(sd)})
    keys = {k for k in sd if not k.endswith("relative_position_index")}
    assert set(back) == keys
    for k in keys:
        assert torch.equal(back[k], sd[k]), k


def test_head_mask_restatement_matches_hf_with_scaled_value_heads():
    """transformers 4.7.0 multiplies a head's attention probabilities by head_mask[l, h]; the installed release no longer
    takes the argument, so the restatement is pinned through an identity: probs * m @ v == probs @ (m * v), i.e. the live HF
    module with the value projection of head h scaled by m (weights and bias) must give the masked oracle's logits.  The
    per-head contexts the oracle returns are pinned the same way (the unmasked ones against a forward hook on the module)."""
    spec = ViTSpec.deit("tiny", layers=3, heads=[3] * 3, inter=[768] * 3)
    model = ovit.build_hf_model(spec, seed=4, stress=True)
    sd = ovit.state_dict_of(model)
    x = ovit.synthetic_images(2, seed=5)
    mask = torch.tensor([[1.0, 0.0, 1.0], [0.5, 1.0, 0.0], [1.0, 1.0, 1.0]])
    ctxs = []
    got = ovit.vit_forward(sd, spec, x, head_mask=mask, ctx_out=ctxs)
    assert [tuple(c.shape) for c in ctxs] == [(2, 3, 197, 64)] * 3
    with torch.no_grad():
        for l in range(3):
            v = model.vit.encoder.layer[l].attention.attention.value
            scale = mask[l].repeat_interleave(64)
            v.weight.mul_(scale[:, None])
            v.bias.mul_(scale)
        want = model(pixel_values=x).logits
    assert (got - want).abs().max() < 5e-5
    # a masked head (m = 0) is equivalent to the pruned model of the pinned pruning restatement
    base = ovit.vit_forward(sd, spec, x)
    assert (got - base).abs().max() > 1e-3      # the mask does something


def test_deit_198_token_restatement_matches_hf():
    """HF `DeiTForImageClassification` (cls + distillation token, one classifier on the cls row,
    SITE/models/deit/modeling_deit.py:595-659): the oracle's 198-token branch against the live module."""
    from transformers import DeiTConfig, DeiTForImageClassification
    cfg = DeiTConfig(hidden_size=192, num_hidden_layers=2, num_attention_heads=3, intermediate_size=768, num_labels=10,
                     attn_implementation="eager")
    torch.manual_seed(9)
    model = DeiTForImageClassification(cfg).eval()
    with torch.no_grad():
        model.deit.embeddings.cls_token.normal_(0, 0.02)
        model.deit.embeddings.distillation_token.normal_(0, 0.02)
        model.deit.embeddings.position_embeddings.normal_(0, 0.02)
    sd = {k.replace("deit.", "vit.", 1) if k.startswith("deit.") else k: v.detach().float() for k, v in model.state_dict().items()}
    spec = ViTSpec(hidden=192, layers=2, heads=[3, 3], inter=[768, 768], tokens=198, num_labels=10)
    x = ovit.synthetic_images(2, seed=3)
    got = ovit.vit_forward(sd, spec, x)
    with torch.no_grad():
        want = model(pixel_values=x).logits
    assert (got - want).abs().max() < 5e-5


def test_two_independent_t2t_front_end_restatements_agree():
    """oracle/t2t.py (torch: unfold / einsum) against oracle/t2t_np.py (NumPy float64, explicit index arithmetic, written
    separately from the reference's text): a shared misreading of tf_Unfold's depth order, of the k | q | v split or of the
    performer algebra would have to be made twice.  Parity of both stays unpinned (no TensorFlow in the image)."""
    from oracle import t2t_np
    sd = ot2t.init_t2t_vit(hidden=64, depth=1, num_heads=1, mlp_ratio=2.0, seed=5, stress=True)
    x = ovit.synthetic_images(1, seed=6, channels_last=True)
    a = ot2t.t2t_module(sd, x).double().numpy()
    b = t2t_np.t2t_tokens(sd, x)
    assert a.shape == b.shape == (1, 196, 64)
    assert np.abs(a - b).max() < 2e-4 * max(1.0, np.abs(b).max())
    # the soft split alone is bit-level: a gather
    u = ot2t.unfold_nhwc(x, 7, 4, 2).double().numpy()
    assert np.array_equal(u, t2t_np.soft_split(x.double().numpy(), 7, 4, 2))
    y = torch.randn(1, 28, 28, 8, generator=torch.Generator().manual_seed(1))
    assert np.array_equal(ot2t.unfold_nhwc(y, 3, 2, 1).double().numpy(), t2t_np.soft_split(y.double().numpy(), 3, 2, 1))


def test_two_tf_dialect_restatements_agree_and_einops_patterns_hold():
    """oracle/tf_vit.py (torch, hand-written permutes) against oracle/tf_vit_np.py (NumPy float64; every re-layout through
    einops.rearrange with the reference's own pattern strings, modeling/models/vit.py:33-34, modeling/layers/attention.py:19-20):
    the patch pixel order (p1 p2 c), the fused (qkv h d) column order and the head merge are pinned to the library the reference
    calls; the LN-in-skip dataflow would have to be misread twice.  Uneven heads / FFN widths = ViT_Pruned 'layerwise'."""
    from oracle import tf_vit_np
    for kw, seed in ((dict(dim=192, depth=2), 0), (dict(dim=128, depth=3, heads=[1, 2, 1], inter=[230, 64, 407], mlp_dim=96), 3)):
        sd, heads, _ = otf.init_tf_vit(seed=seed, stress=True, **kw)
        x = ovit.synthetic_images(2, seed=seed + 1)
        a = otf.tf_vit_forward(sd, x, heads).double().numpy()
        b = tf_vit_np.tf_vit_forward_np(sd, x, heads)
        assert a.shape == b.shape == (2, 1000)
        assert np.abs(a - b).max() < 2e-4 * max(1.0, np.abs(b).max()), np.abs(a - b).max()
        assert (a.argmax(-1) == b.argmax(-1)).all()
    # the hand-written patch flattening of oracle/tf_vit.py is bit-identical to the einops pattern (a pure gather)
    from einops import rearrange
    img = torch.randn(2, 3, 32, 48)
    mine = img.reshape(2, 3, 2, 16, 3, 16).permute(0, 2, 4, 3, 5, 1).reshape(2, 6, 768)
    assert torch.equal(mine, rearrange(img, tf_vit_np.PATCHES, p1=16, p2=16))


def test_distilled_timm_restatement_matches_hf_with_teacher():
    """facebookresearch/deit ``DistilledVisionTransformer`` (dist_token, head_dist, eval output = mean of the two heads) restated in
    oracle.timm_vit == HF ``DeiTForImageClassificationWithTeacher`` with eps 1e-6 on converted weights; the product's timm
    adapter maps it back onto the two-head layout that ``normalise_keys`` folds into one classifier."""
    from edgevisiontransformer_b200.dialects import timm_vit_to_canonical
    from edgevisiontransformer_b200.modeling_vit import config_from_state_dict, normalise_keys
    from oracle import timm_vit as otimm
    hf = otimm.build_hf_distilled()
    sd = {("vit." + k[5:] if k.startswith("deit.") else k): v.detach() for k, v in hf.state_dict().items()}
    tsd = otimm.hf_to_timm(sd)
    assert tsd["pos_embed"].shape == (1, 198, 192) and "head_dist.weight" in tsd
    x = ovit.synthetic_images(2, seed=2)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    assert (otimm.timm_vit_forward(tsd, x, num_heads=3) - want).abs().max() < 1e-5
    csd, kw = timm_vit_to_canonical(tsd)
    folded = normalise_keys(csd)
    assert folded["classifier.weight"].shape == (1000, 384)
    cfg = config_from_state_dict(folded, **kw)
    assert cfg.head_rows == 2 and cfg.tokens == 198
