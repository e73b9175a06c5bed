"""BASELINE.json configs 2, 4 and 5 (and HF DeiT-198) at their STATED sizes, in the pattern of
test_full_size_config3_properties: the large batch goes through the large-M kernels (CTA-pair GEMMs, persistent attention),
picked images are compared with the CPU oracle at the bf16 tolerance, and -- images being independent units -- the same
images pushed through the small-batch kernels must agree with their large-batch logits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ViTSpec  # noqa: E402
from oracle import pruning as opr  # noqa: E402
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


def _big_vs_small_vs_oracle(m, sd, spec, batch, pick, seed, n_oracle=3):
    from edgevisiontransformer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(batch, 3, 224, 224, device="cuda", generator=g)
    big = m(x).logits
    assert big.shape == (batch, spec.num_labels) and torch.isfinite(big).all()
    idx = torch.tensor(pick, device="cuda")
    try:
        ops.set_gemm_split_k(False)
        small = m(x[idx].contiguous()).logits             # same images through the small-M kernels
    finally:
        ops.set_gemm_split_k(True)
    r = ovit.compare_logits(big[idx], small)
    assert r["max_abs"] <= 5e-3 and r["top1_agree"] == 1.0, r
    want = ovit.vit_forward(sd, spec, x[idx[:n_oracle]].cpu())
    _check(big[idx[:n_oracle]], want)
    assert torch.equal(m(x).logits, big)                  # fixed schedules: a second pass reproduces the first bit for bit
    return big


def test_config2_deit_small_batch_256():
    """BASELINE config 2: DeiT-Small patch16-224 bf16, batch 256 on one B200."""
    spec = ViTSpec.deit("small")
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=1, stress=True))
    m = _model(sd, max_batch=256, keep_params=False)
    _big_vs_small_vs_oracle(m, sd, spec, 256, [0, 1, 127, 128, 255], seed=21)


@pytest.mark.parametrize("case", ["h1_d230", "head18_uneven"])
def test_config4_pruned_tiny_batch_1024(case, golden_dir):
    """BASELINE config 4: nn_pruning / are16heads DeiT-Tiny with per-layer head counts and FFN widths, batch 1024: the
    CTA-pair GEMMs with N = 230 / K = 230 (leading dimension 232) and one- and two-head attention at M = 201 728 rows."""
    spec0 = ViTSpec.deit("tiny")
    sd0 = ovit.state_dict_of(ovit.build_hf_model(spec0, seed=4, stress=True))
    if case == "h1_d230":
        heads_kept, inter_kept = [[0]] * 12, [230] * 12
    else:
        heads_kept = opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3)
        inter_kept = [230, 231, 200, 256, 1, 8, 407, 230, 300, 150, 768, 64]
    _, pruned, _ = opr.synthesize_pruned(sd0, heads_kept, inter_kept, seed=7)
    spec = ovit.spec_from_state_dict(pruned)
    assert spec.heads == [len(h) for h in heads_kept] and spec.inter == list(inter_kept)
    m = _model(pruned, max_batch=1024, keep_params=False)
    assert m.config.heads == spec.heads and m.config.intermediate == spec.inter
    _big_vs_small_vs_oracle(m, pruned, spec, 1024, [0, 1, 511, 512, 1023], seed=22)
    if case == "h1_d230":                                  # the committed fixture of the vendored optimize_model run
        f = np.load(golden_dir + "/pruned_tiny_h1_d230.npz")
        _check(m(ovit.synthetic_images(2, seed=1).cuda()).logits, torch.from_numpy(f["logits_opt"]))


def test_config5_t2t_vit_14_batch_256():
    """BASELINE config 5: get_t2t_vit_14 (depth 14, 6 heads, hidden 384, mlp ratio 3; modeling/models/t2t_vit.py:147-148)
    at a large batch, against the restatement (parity unpinned: the reference's TF code cannot run here) on picked images."""
    from edgevisiontransformer_b200 import ops
    from edgevisiontransformer_b200.modeling_t2t import get_t2t_vit_14
    from oracle import t2t as ot2t
    sd = ot2t.init_t2t_vit(hidden=384, depth=14, num_heads=6, mlp_ratio=3.0, seed=2, stress=True)
    m = get_t2t_vit_14(sd, max_batch=256)
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(256, 224, 224, 3, device="cuda", generator=g)
    big = m(x).logits
    assert big.shape == (256, 1000) and torch.isfinite(big).all()
    idx = torch.tensor([0, 1, 128, 255], device="cuda")
    try:
        ops.set_gemm_split_k(False)
        small = m(x[idx].contiguous()).logits
    finally:
        ops.set_gemm_split_k(True)
    r = ovit.compare_logits(big[idx], small)
    # The large batch runs each residual projection with the following LayerNorm in its epilogue (csrc/gemm_rowln.cu), the small
    # one runs the LayerNorm kernel: same f32 statistics in a different summation order.  In this dialect the skip connection
    # carries the normalised rows, so last-bit differences feed the residual stream of 14 layers: measured 1.2e-2 on these
    # stress weights (5e-3 and less for the HF dialect above), inside the bf16 contract either way.
    assert r["max_abs"] <= BF16_TOL and r["top1_agree"] == 1.0, r
    want = ot2t.t2t_vit_forward(sd, x[idx[:2]].cpu(), 14, 6)
    r = _check(big[idx[:2]], want)
    print("t2t_vit_14 bs256 max_abs", r["max_abs"])


def test_hf_deit_198_tokens_batch_512():
    """198-token DeiT (cls + distillation token) through the large-M path: S = 198 is a ragged second query tile with 70
    rows and a 208-key score tile with 10 masked columns."""
    spec = ViTSpec.deit("tiny", tokens=198)
    sd = ovit.state_dict_of(ovit.build_hf_model(ViTSpec.deit("tiny"), seed=6, stress=True))
    g = torch.Generator().manual_seed(3)
    sd["vit.embeddings.distillation_token"] = torch.randn(1, 1, 192, generator=g) * 0.02
    sd["vit.embeddings.position_embeddings"] = torch.randn(1, 198, 192, generator=g) * 0.02
    m = _model(sd, max_batch=512, keep_params=False)
    assert m.config.tokens == 198
    _big_vs_small_vs_oracle(m, sd, spec, 512, [0, 255, 256, 511], seed=24)
