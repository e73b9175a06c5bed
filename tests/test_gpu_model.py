"""GPU parity of the whole forward (through the C ABI model entry points) against the CPU oracle.

Tolerance (BASELINE.json north_star): bf16 mode -> logits max-abs <= 2e-2 and 100 % top-1 agreement."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ViTSpec  # noqa: E402
from oracle import pruning as opr  # noqa: E402
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


@pytest.mark.parametrize("kind,seed,stress,bs", [("tiny", 0, False, 2), ("tiny", 3, True, 5), ("small", 0, True, 3)])
def test_deit_matches_oracle_and_golden(kind, seed, stress, bs, golden_dir):
    spec = ViTSpec.deit(kind)
    hf = ovit.build_hf_model(spec, seed=seed, stress=stress)
    sd = ovit.state_dict_of(hf)
    x = ovit.synthetic_images(bs, seed=1)
    want = ovit.vit_forward(sd, spec, x)
    m = _model(sd)
    got = m(x.cuda()).logits
    _check(got, want)
    name = {("tiny", 0): "hf_tiny_s0.npz", ("tiny", 3): "hf_tiny_s3_stress.npz", ("small", 0): "hf_small_s0_stress.npz"}[(kind, seed)]
    f = np.load(os.path.join(golden_dir, name))
    nb = int(f["batch"])
    _check(got[:nb], torch.from_numpy(f["logits"]))       # committed fixture from the HF forward itself


def test_tf32_mode_config1(golden_dir):
    """BASELINE config 1: DeiT-Tiny, seed-0 weights, batch 1, input seed 1 -> max-abs 1e-3 in the tf32 mode."""
    spec = ViTSpec.deit("tiny")
    for seed, stress in ((0, False), (3, True)):
        sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=seed, stress=stress))
        x = ovit.synthetic_images(2, seed=1)
        want = ovit.vit_forward(sd, spec, x)
        m = _model(sd, precision="tf32")
        got = m(x.cuda()).logits
        _check(got, want, tol=1e-3)
        _check(m(x[:1].cuda()).logits, want[:1], tol=1e-3)
    f = np.load(os.path.join(golden_dir, "hf_tiny_s3_stress.npz"))
    _check(got, torch.from_numpy(f["logits"]), tol=1e-3)


def test_from_hf_and_module_surface():
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    spec = ViTSpec.deit("tiny")
    hf = ovit.build_hf_model(spec, seed=2, stress=True)
    m = B200ViTForImageClassification.from_hf(hf)
    x = ovit.synthetic_images(3, seed=5)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    out = m(pixel_values=x.cuda())                  # kw call as deit_pruning/src/trainer.py:58
    _check(out.logits, want)
    assert next(m.parameters()).device.type == "cuda"        # deit_pruning/src/utils.py:177
    assert m.num_parameters() == sum(p.numel() for p in hf.parameters())
    assert m.eval() is m and m.to("cuda") is m
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 192, 192, device="cuda"))
    with pytest.raises(RuntimeError):
        m(x)                                        # CPU input: no fallback
    g = m.forward_graphed(x.cuda()).logits
    # small batch: split-K reduce-adds land in any order; a last-bit f32 difference can flip bf16 roundings downstream
    assert (g - out.logits).abs().max().item() < 1e-2
    from edgevisiontransformer_b200 import ops
    try:
        ops.set_gemm_split_k(False)                 # without K splitting: graph replay == eager, bit for bit
        e = m(pixel_values=x.cuda()).logits
        m._graphs.clear()
        assert torch.equal(m.forward_graphed(x.cuda()).logits, e)
    finally:
        ops.set_gemm_split_k(True)
        m._graphs.clear()
    assert m.launches_per_forward() == 3 + 7 * 12 + 2


@pytest.mark.parametrize("case", ["h1_d230", "head18_uneven"])
def test_pruned_models(case, golden_dir):
    spec = ViTSpec.deit("tiny")
    hf = ovit.build_hf_model(spec, seed=4, stress=True)
    sd = ovit.state_dict_of(hf)
    f = np.load(os.path.join(golden_dir, f"pruned_tiny_{case}.npz"))
    heads_kept = [[0]] * 12 if case == "h1_d230" else opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3)
    full, pruned, _ = opr.synthesize_pruned(sd, heads_kept, [int(v) for v in f["inter_kept"]], seed=7)
    m = _model(pruned)
    assert m.config.heads == [len(h) for h in heads_kept]
    assert m.config.intermediate == [int(v) for v in f["inter_kept"]]
    x = ovit.synthetic_images(2, seed=1)
    got = m(x.cuda()).logits
    _check(got, torch.from_numpy(f["logits_opt"]))      # fixture from vendored optimize_model + prune_heads
    # the full-size checkpoint with zero rows/cols gives the same function
    got_full = _model(full)(x.cuda()).logits
    _check(got_full, torch.from_numpy(f["logits_full"]))


def test_tf_dialect_matches_restatement():
    from edgevisiontransformer_b200.dialects import tf_vit_to_canonical
    sd, heads, inter = otf.init_tf_vit(dim=192, depth=12, seed=1, stress=True)
    x = ovit.synthetic_images(3, seed=2)
    want = otf.tf_vit_forward(sd, x, heads)
    csd, kw = tf_vit_to_canonical(sd, heads)
    got = _model(csd, **kw)(x.cuda()).logits
    _check(got, want)
    # ViT_Pruned 'layerwise' encoding (modeling/models/vit.py:58-97)
    h2, i2 = opr.parse_prune_encoding("layerwise_" + "_".join(["h2-d0.5", "h1-d0.3", "h3-d1.0"] * 4), 12, 768)
    sd, heads, inter = otf.init_tf_vit(dim=192, depth=12, heads=h2, inter=i2, seed=3, stress=True)
    want = otf.tf_vit_forward(sd, x, heads)
    csd, kw = tf_vit_to_canonical(sd, heads)
    m = _model(csd, **kw)
    assert m.config.heads == h2 and m.config.intermediate == i2
    _check(m(x.cuda()).logits, want)


def test_batch_independence_and_chunking():
    """Size-independent property: images are independent units -> any batch split gives identical logits."""
    spec = ViTSpec.deit("tiny")
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=0))
    from edgevisiontransformer_b200 import ops
    x = ovit.synthetic_images(37, seed=9).cuda()
    m = _model(sd, max_batch=16)
    try:
        ops.set_gemm_split_k(False)                 # bit-exact invariance holds whenever K is not split
        a = m(x).logits
        b = torch.cat([m(x[:5]).logits, m(x[5:]).logits])
        assert torch.equal(a, b)
        perm = torch.randperm(37, device="cuda")
        assert torch.equal(m(x[perm]).logits, a[perm])
    finally:
        ops.set_gemm_split_k(True)
    # default (latency-tuned) small-batch path: same logits up to the f32 summation order of the K splits (which can
    # flip bf16 roundings downstream: differences of a few 1e-3, inside the 2e-2 parity budget)
    c = torch.cat([m(x[:5]).logits, m(x[5:]).logits])
    assert (c - a).abs().max().item() < 1e-2


def test_t2t_front_end_and_model():
    """T2T-ViT against the torch restatement of modeling/models/t2t_vit.py (parity unpinned by the reference: no TF here)."""
    from edgevisiontransformer_b200 import ops
    from edgevisiontransformer_b200.modeling_t2t import B200T2TViT
    from oracle import t2t as ot2t
    sd = ot2t.init_t2t_vit(hidden=384, depth=3, num_heads=6, mlp_ratio=3.0, seed=0, stress=True)
    x = ovit.synthetic_images(2, seed=4, channels_last=True)
    want_logits, want_tok = ot2t.t2t_vit_forward(sd, x, 3, 6, return_tokens=True)
    m = B200T2TViT(sd, depth=3, num_heads=6)
    # soft split + LN alone (bit-level gather, LN in f32)
    u = ops.unfold_ln_nhwc(x.cuda(), 7, 4, 2)
    assert torch.equal(u[:, :147].cpu(), ot2t.unfold_nhwc(x, 7, 4, 2).reshape(-1, 147).bfloat16())
    # tokens entering `project`: compare after the project Dense on the CPU side
    pm = m.tokens(x.cuda())
    got_tok = pm[:, :576].float().cpu() @ sd["t2t.project.kernel"] + sd["t2t.project.bias"]
    err = (got_tok.view(2, 196, 384) - want_tok).abs().max().item()
    assert err < 5e-2, err
    r = ovit.compare_logits(m(x.cuda()).logits, want_logits)
    assert r["max_abs"] <= BF16_TOL and r["top1_agree"] == 1.0, r
    # the front-end inside the C++ runtime (evt_model_forward, spec.t2t) issues the same kernels as the op-level composition
    try:
        ops.set_gemm_split_k(False)
        assert torch.equal(m(x.cuda()).logits, m.core.forward_embedded(m.tokens(x.cuda())).logits)
    finally:
        ops.set_gemm_split_k(True)
    assert m.launches_per_forward() == 11 + 2 + 7 * 3 + 2     # front-end: 2 x (unfold+LN, kqv, 3 performer kernels) + last split
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 224, 224, device="cuda"))
    for _ in range(2):                                   # latency path: one CUDA graph over front-end + encoder
        r = ovit.compare_logits(m.forward_graphed(x.cuda()).logits, want_logits)
        assert r["max_abs"] <= BF16_TOL and r["top1_agree"] == 1.0, r
    # a larger batch makes the core reallocate its workspace: graphs captured over the old one must not be replayed
    small = m.forward_graphed(x.cuda()).logits.clone()
    big = ovit.synthetic_images(8, seed=5, channels_last=True).cuda()
    big_eager = m(big).logits.clone()
    again = m.forward_graphed(x.cuda()).logits
    r = ovit.compare_logits(again, small)
    assert r["max_abs"] <= BF16_TOL, r                   # same numbers as before the reallocation up to the split-K reduce order
    r = ovit.compare_logits(m(big).logits, big_eager)    # and the replay did not scribble over anything the eager path uses
    assert r["max_abs"] <= BF16_TOL, r


def test_stage_profile_tap():
    """evt_model_profile_begin/end: launches are attributed to the eight stages, results are unchanged by the tap."""
    spec = ViTSpec.deit("tiny")
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=0, stress=True))
    m = _model(sd)
    x = ovit.synthetic_images(4, seed=2).cuda()
    ref = m(x).logits.clone()
    m.profile_begin()
    got = m(x).logits
    got2 = m(x).logits
    st = m.profile_end()
    _check(got2, ref.cpu(), tol=5e-3)
    _check(got, ref.cpu(), tol=5e-3)             # split-K reduce order may differ run to run (include/evt.h)
    L = spec.layers
    assert {k: n for k, (_, n) in st.items()} == {"embed": 6, "layernorm": 4 * L, "qkv": 2 * L, "attention": 2 * L,
                                                  "out_proj": 2 * L, "fc1": 2 * L, "fc2": 2 * L, "head": 4}
    assert all(ms > 0 for ms, _ in st.values())
    assert sum(n for _, n in st.values()) == 2 * m.launches_per_forward()
    with pytest.raises(RuntimeError):
        m.profile_end()                           # not begun


def test_timm_dialect(golden_dir):
    """timm / facebookresearch-deit checkpoints (fused qkv, eps 1e-6): utils.py:52-62, tools.py:244-263."""
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    from oracle import timm_vit as otimm
    sd = ovit.state_dict_of(ovit.build_hf_model(ViTSpec.deit("tiny", eps=1e-6), seed=5, stress=True))
    tsd = otimm.hf_to_timm(sd)
    x = ovit.synthetic_images(2, seed=1)
    want = otimm.timm_vit_forward(tsd, x, num_heads=3)
    m = B200ViTForImageClassification.from_timm(tsd)
    assert m.config.layer_norm_eps == 1e-6 and m.config.heads == [3] * 12 and m.config.image_size == 224
    got = m(x.cuda()).logits
    _check(got, want)
    _check(got, torch.from_numpy(np.load(os.path.join(golden_dir, "timm_tiny_s5.npz"))["logits"]))
    _check(B200ViTForImageClassification.from_timm(tsd, precision="tf32")(x.cuda()).logits, want, tol=1e-3)


def test_timm_distilled_dialect():
    """facebookresearch/deit hub ``deit_*_distilled_patch16_224`` layout (dist_token + head_dist, logits = mean of the two heads)
    through from_timm; the checker is HF ``DeiTForImageClassificationWithTeacher`` (eps 1e-6) on the same weights."""
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    from oracle import timm_vit as otimm
    hf = otimm.build_hf_distilled(layers=12)
    sd = {("vit." + k[5:] if k.startswith("deit.") else k): v.detach() for k, v in hf.state_dict().items()}
    tsd = otimm.hf_to_timm(sd)
    x = ovit.synthetic_images(3, seed=4)
    with torch.no_grad():
        want = hf(pixel_values=x).logits
    m = B200ViTForImageClassification.from_timm({"model": tsd})
    assert m.config.tokens == 198 and m.config.head_rows == 2 and m.config.image_size == 224
    _check(m(x.cuda()).logits, want)


def test_full_size_config3_properties():
    """BASELINE config 3 at its full size (DeiT-Base, global batch 4096 as 4 chunks of 1024 -- the CTA-pair GEMMs at
    M = 201 728 rows): images are independent units, so the logits of any image must not depend on what it is batched
    with; a few images are also checked against the CPU oracle at the bf16 tolerance."""
    from edgevisiontransformer_b200 import ops
    spec = ViTSpec.deit("base")
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=0, stress=True))
    m = _model(sd, max_batch=1024, keep_params=False)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(4096, 3, 224, 224, device="cuda", generator=g)
    big = m(x).logits
    assert big.shape == (4096, 1000) and torch.isfinite(big).all()
    pick = torch.tensor([0, 1, 1023, 1024, 2500, 4095], device="cuda")
    try:
        ops.set_gemm_split_k(False)
        small = m(x[pick].contiguous()).logits          # same images, batch 6: small-M kernels (1-CTA GEMM, narrow tiles)
    finally:
        ops.set_gemm_split_k(True)
    r = ovit.compare_logits(big[pick], small)
    assert r["max_abs"] <= 5e-3 and r["top1_agree"] == 1.0, r
    want = ovit.vit_forward(sd, spec, x[pick[:3]].cpu())
    _check(big[pick[:3]], want)
    # a second pass over the same inputs reproduces the first bit for bit (no split K at this size, fixed schedules)
    assert torch.equal(m(x).logits, big)
