"""GPU tests of the forward options of the C ABI (evt_model_forward_ex) and of the module surface that exposes them:
pixel storage types fused into the patch gather, model-level head masks (`mask_heads`, HF `head_mask=`), per-layer
context capture (`context_layer_val`), HF `DeiTForImageClassification` (198 tokens), and the regressions the round-1
advisor flagged (T2T graph after a workspace reallocation, TF dialect in the tf32 mode, swallowed kwargs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ViTSpec  # noqa: E402
from oracle import tf_vit as otf  # noqa: E402
from oracle import vit as ovit  # noqa: E402

BF16_TOL = 2e-2


def _model(sd, **kw):
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    return B200ViTForImageClassification.from_state_dict(sd, **kw)


def _check(got, want, tol=BF16_TOL):
    r = ovit.compare_logits(got, want)
    assert r["max_abs"] <= tol, r
    assert r["top1_agree"] == 1.0, r
    return r


def _tiny(seed=3, layers=12):
    spec = ViTSpec.deit("tiny", layers=layers, heads=[3] * layers, inter=[768] * layers)
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=seed, stress=True))
    return spec, sd


def test_pixel_dtypes_are_converted_inside_the_patch_gather():
    """bf16 pixels: the gather reads bf16 and writes the same bf16 -> bit-identical to the f32 path on the same values.
    u8 pixels: (x / 255 - mean) / std is applied in the gather (deit_pruning/src/utils.py:105-107 on the GPU)."""
    from edgevisiontransformer_b200 import ops
    from edgevisiontransformer_b200.modeling_vit import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
    spec, sd = _tiny()
    m = _model(sd)
    g = torch.Generator().manual_seed(11)
    try:
        ops.set_gemm_split_k(False)                  # bit-exact comparisons need a fixed summation order
        xb = torch.randn(5, 3, 224, 224, generator=g).bfloat16()
        a = m(xb.cuda()).logits
        b = m(xb.float().cuda()).logits
        assert torch.equal(a, b)
        _check(a, ovit.vit_forward(sd, spec, xb.float()))
        xu = torch.randint(0, 256, (4, 3, 224, 224), generator=g, dtype=torch.uint8)
        mean = torch.tensor(IMAGENET_DEFAULT_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_DEFAULT_STD).view(1, 3, 1, 1)
        xn = (xu.float() / 255.0 - mean) / std       # transforms.ToTensor() + Normalize
        got = m(xu.cuda()).logits
        _check(got, ovit.vit_forward(sd, spec, xn))
        # against the library's own f32 path on the normalised pixels: only bf16 roundings of single pixels can differ
        assert (got - m(xn.cuda()).logits).abs().max().item() < 5e-3
        # another normalisation
        m.pixel_mean, m.pixel_std = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)
        _check(m(xu.cuda()).logits, ovit.vit_forward(sd, spec, (xu.float() / 255.0 - 0.5) / 0.5))
    finally:
        ops.set_gemm_split_k(True)
    # tf32 accuracy mode takes the same pixel types
    mt = _model(sd, precision="tf32")
    _check(mt(xb.cuda()).logits, ovit.vit_forward(sd, spec, xb.float()), tol=1e-3)


def test_model_level_head_mask_and_context_capture():
    """are_16_heads: `model.vit.mask_heads(to_prune)` (run_classifier.py:247-250), HF `head_mask=` and the per-head
    context `context_layer_val` read by calculate_head_importance (classifier_eval.py:183-191)."""
    spec, sd = _tiny(seed=4, layers=4)
    m = _model(sd)
    x = ovit.synthetic_images(3, seed=6)
    base = ovit.vit_forward(sd, spec, x)
    # dict form: zero heads
    to_mask = {0: {1}, 2: {0, 2}}
    mask = torch.ones(4, 3)
    for l, hs in to_mask.items():
        for h in hs:
            mask[l, h] = 0
    want = ovit.vit_forward(sd, spec, x, head_mask=mask)
    m.vit.mask_heads(to_mask)
    got = m(x.cuda()).logits
    _check(got, want)
    assert (got.cpu() - base).abs().max() > 1e-3                     # the mask is not a no-op
    _check(m.forward_graphed(x.cuda()).logits, want)                 # the graph path must not drop the mask
    m.mask_heads(None)
    _check(m(x.cuda()).logits, base)
    # HF forward(head_mask=...) with fractional values, [layers, heads] and [heads]
    frac = torch.tensor([[1.0, 0.5, 0.0], [0.25, 1.0, 1.0], [1.0, 1.0, 1.0], [0.0, 0.0, 1.0]])
    _check(m(x.cuda(), head_mask=frac).logits, ovit.vit_forward(sd, spec, x, head_mask=frac))
    one = torch.tensor([1.0, 0.0, 1.0])
    _check(m(x.cuda(), head_mask=one).logits, ovit.vit_forward(sd, spec, x, head_mask=one.expand(4, 3)))
    with pytest.raises(ValueError):
        m(x.cuda(), head_mask=torch.ones(3, 3))
    with pytest.raises(ValueError):
        m.mask_heads({0: [3]})
    # context capture: [B, heads, tokens, 64] per layer, the forward half of calculate_head_importance
    ctx_want = []
    ovit.vit_forward(sd, spec, x, ctx_out=ctx_want)
    m.capture_context(True)
    out = m(x.cuda()).logits
    _check(out, base)
    assert len(m.context_layers) == 4
    for got_c, want_c in zip(m.context_layers, ctx_want):
        assert tuple(got_c.shape) == (3, 3, 197, 64)
        err = (got_c.float().cpu() - want_c).abs().max().item()
        assert err <= 2e-2 * max(1.0, want_c.abs().max().item()), err
    # masked + captured together: a masked head's context is zero
    m.mask_heads({1: [2]})
    m(x.cuda())
    assert m.context_layers[1][:, 2].abs().max().item() == 0.0
    assert m.context_layers[1][:, 1].abs().max().item() > 0.0
    m.capture_context(False)
    m.mask_heads({})
    assert m.context_layers == []
    # tf32 mode: same options, f32 context
    mt = _model(sd, precision="tf32")
    mt.capture_context(True)
    _check(mt(x.cuda(), head_mask=frac).logits, ovit.vit_forward(sd, spec, x, head_mask=frac), tol=1e-3)
    assert mt.context_layers[0].dtype == torch.float32


def test_unsupported_forward_kwargs_raise_instead_of_being_ignored():
    spec, sd = _tiny(layers=2)
    m = _model(sd)
    x = ovit.synthetic_images(1, seed=1).cuda()
    m(x, output_attentions=False, return_dict=True)                  # harmless values pass
    for kw in ({"output_attentions": True}, {"output_hidden_states": True}, {"labels": torch.zeros(1, dtype=torch.long)},
               {"interpolate_pos_encoding": True}):
        with pytest.raises(NotImplementedError):
            m(x, **kw)
    with pytest.raises(TypeError):
        m(x, bool_masked_pos=None)                                   # not part of the classification forward
    from edgevisiontransformer_b200.modeling_vit import B200ViTConfig, config_from_state_dict
    cfg = config_from_state_dict(sd, hidden_act="tanh")              # HF: plain nn.Tanh, not a GELU
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    with pytest.raises(ValueError):
        B200ViTForImageClassification(cfg, sd)
    from edgevisiontransformer_b200 import ops
    with pytest.raises((KeyError, ValueError)):
        ops.linear(torch.zeros(8, 64, device="cuda").bfloat16(), torch.zeros(8, 64, device="cuda").bfloat16(), None, act="tanh")
    assert isinstance(cfg, B200ViTConfig)


def test_hf_deit_for_image_classification_198_tokens():
    """HF `DeiTForImageClassification` (cls + distillation token, classifier on the cls row,
    SITE/models/deit/modeling_deit.py:595-659) through from_hf, at model level."""
    from transformers import DeiTConfig, DeiTForImageClassification
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    cfg = DeiTConfig(hidden_size=192, num_hidden_layers=12, num_attention_heads=3, intermediate_size=768, num_labels=1000,
                     attn_implementation="eager")
    torch.manual_seed(12)
    hf = DeiTForImageClassification(cfg).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias") or "layernorm" in n:
                p.add_(torch.randn_like(p) * 0.1)
        hf.deit.embeddings.cls_token.normal_(0, 0.02)
        hf.deit.embeddings.distillation_token.normal_(0, 0.02)
        hf.deit.embeddings.position_embeddings.normal_(0, 0.02)
    m = B200ViTForImageClassification.from_hf(hf)
    assert m.config.tokens == 198
    x = ovit.synthetic_images(4, seed=8)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    _check(m(x.cuda()).logits, want)
    _check(m.forward_graphed(x[:1].cuda()).logits, want[:1])


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_hf_distilled_deit_with_teacher_two_heads(precision):
    """HF `DeiTForImageClassificationWithTeacher` (the public distilled DeiT layout): logits = mean of the cls head on row 0 and
    the distillation head on row 1.  Both rows go through the final LayerNorm and ONE folded classifier GEMM (K = 2 D,
    evt_model_spec.head_rows = 2); the live HF module is the checker."""
    from transformers import DeiTConfig, DeiTForImageClassificationWithTeacher
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    cfg = DeiTConfig(hidden_size=192, num_hidden_layers=12, num_attention_heads=3, intermediate_size=768, num_labels=1000,
                     attn_implementation="eager")
    torch.manual_seed(21)
    hf = DeiTForImageClassificationWithTeacher(cfg).eval()
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias") or "layernorm" in n:
                p.add_(torch.randn_like(p) * 0.1)
        hf.deit.embeddings.cls_token.normal_(0, 0.02)
        hf.deit.embeddings.distillation_token.normal_(0, 0.02)
        hf.deit.embeddings.position_embeddings.normal_(0, 0.02)
    m = B200ViTForImageClassification.from_hf(hf, precision=precision)
    assert m.config.tokens == 198 and m.config.head_rows == 2
    x = ovit.synthetic_images(5, seed=9)
    with torch.no_grad():
        o = hf(pixel_values=x)
        want = o.logits
        assert torch.allclose(want, (o.cls_logits + o.distillation_logits) / 2, atol=1e-6)
    got = m(x.cuda()).logits
    if precision == "tf32":
        assert (got.cpu() - want).abs().max().item() < 1e-3
    else:
        _check(got, want)
        _check(m.forward_graphed(x[:1].cuda()).logits, want[:1])
    # the cls head alone would have been a different answer: the test inputs tell the two apart
    assert (o.cls_logits - want).abs().max().item() > 5e-2


def test_tf_dialect_in_tf32_mode_keeps_the_skip_connection_in_full_precision():
    """In the TF dialect the LayerNorm output IS the skip connection (modeling/layers/norm.py:10-12): the f32 copy written
    back into the residual stream must not be rounded to tf32 (only the GEMM operand copy is).

    Tolerance: the 1e-3 of BASELINE.json is stated for (and met by, 7.5e-4) the HF forward of config 1.  The TF-dialect
    DeiT measured 1.6e-3 on B200 in this mode (12 blocks whose skip path goes through LayerNorm, tanh-GELU MLP head:
    more tf32-rounded operands per logit); asserted at 2e-3 here and recorded in DESIGN.md section 4 -- the parity of this
    dialect is unpinned in any case (no TensorFlow in the image)."""
    from edgevisiontransformer_b200.dialects import tf_vit_to_canonical
    sd, heads, inter = otf.init_tf_vit(dim=192, depth=12, seed=1, stress=True)
    x = ovit.synthetic_images(2, seed=2)
    want = otf.tf_vit_forward(sd, x, heads)
    csd, kw = tf_vit_to_canonical(sd, heads)
    got = _model(csd, precision="tf32", **kw)(x.cuda()).logits
    r = _check(got, want, tol=2e-3)
    print("tf dialect tf32 max_abs", r["max_abs"])


def test_t2t_graph_survives_a_workspace_reallocation():
    """forward_graphed(B=1), then a larger batch (the core reallocates its workspace), then B=1 graphed again: the
    stale graph must be dropped, not replayed into freed memory."""
    from edgevisiontransformer_b200.modeling_t2t import B200T2TViT
    from oracle import t2t as ot2t
    sd = ot2t.init_t2t_vit(hidden=384, depth=2, num_heads=6, mlp_ratio=3.0, seed=1, stress=True)
    m = B200T2TViT(sd, depth=2, num_heads=6, max_batch=16)
    x1 = ovit.synthetic_images(1, seed=4, channels_last=True).cuda()
    x8 = ovit.synthetic_images(8, seed=5, channels_last=True).cuda()
    e1 = m(x1).logits.clone()
    g1 = m.forward_graphed(x1).logits
    assert (g1 - e1).abs().max().item() < 1e-2
    assert len(m._graphs) == 1
    e8 = m(x8).logits.clone()                         # larger batch: workspace grows
    assert len(m._graphs) == 0                        # ... and the graph that pointed at the old one is gone
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(8)]   # recycle the freed block
    g1b = m.forward_graphed(x1).logits
    assert torch.isfinite(g1b).all() and (g1b - e1).abs().max().item() < 1e-2
    g8 = m.forward_graphed(x8).logits
    assert (g8 - e8).abs().max().item() < 1e-2
    assert (m.forward_graphed(x1).logits - e1).abs().max().item() < 1e-2
    del junk
