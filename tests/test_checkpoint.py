"""CPU tests of the checkpoint loader (host logic; no GPU): pruned_heads + zero-FFN elision reproduce the oracle's
pruned shapes and function."""
import numpy as np
import torch

from edgevisiontransformer_b200 import checkpoint as ck
from edgevisiontransformer_b200.modeling_vit import config_from_state_dict, normalise_keys
from oracle import ViTSpec
from oracle import pruning as opr
from oracle import vit as ovit


def _tiny(seed=4):
    return ovit.state_dict_of(ovit.build_hf_model(ViTSpec.deit("tiny"), seed=seed, stress=True))


def test_full_size_checkpoint_with_pruned_heads_and_zero_ffn(tmp_path):
    sd = _tiny()
    heads_kept = opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3)
    inter = [768, 231, 230, 64, 700, 8, 333, 512, 1, 96, 768, 407]
    full, pruned, to_prune = opr.synthesize_pruned(sd, heads_kept, inter, seed=7)
    ck.save_checkpoint(str(tmp_path), full, hidden_size=192, num_attention_heads=3, intermediate_size=768,
                       pruned_heads=to_prune)
    got, kw = ck.load_checkpoint(str(tmp_path))
    cfg = config_from_state_dict(got, **kw)
    assert cfg.heads == [len(h) for h in heads_kept]
    assert cfg.intermediate == inter
    for k, v in pruned.items():
        assert torch.equal(got[k], v), k
    x = ovit.synthetic_images(1, seed=1)
    a = ovit.vit_forward(got, ovit.spec_from_state_dict(got), x)
    b = ovit.vit_forward(full, ViTSpec.deit("tiny"), x)
    assert (a - b).abs().max() < 1e-5


def test_checkpoint_written_after_prune_heads(tmp_path):
    sd = _tiny()
    to_prune = {0: [1], 5: [0, 2]}
    small = opr.prune_heads(sd, to_prune)
    ck.save_checkpoint(str(tmp_path), small, hidden_size=192, num_attention_heads=3, intermediate_size=768,
                       pruned_heads=to_prune, safetensors=False)
    got, kw = ck.load_checkpoint(str(tmp_path))
    cfg = config_from_state_dict(got, **kw)
    assert cfg.heads == [2, 3, 3, 3, 3, 1, 3, 3, 3, 3, 3, 3] and kw["head_size"] == 64
    assert torch.equal(got["vit.encoder.layer.5.attention.attention.key.weight"],
                       small["vit.encoder.layer.5.attention.attention.key.weight"])


def test_zero_heads_without_config_entry_and_all_zero_ffn():
    sd = _tiny()
    full, pruned, _ = opr.synthesize_pruned(sd, [[0]] * 12, [230] * 12, seed=7)   # heads zeroed, not listed anywhere
    work = {k: v.clone() for k, v in full.items()}
    ck.drop_zero_heads_(work, 64)
    ck.drop_zero_ffn_(work)
    cfg = config_from_state_dict(work)
    assert cfg.heads == [1] * 12 and cfg.intermediate == [230] * 12
    for k, v in pruned.items():
        assert torch.equal(work[k], v), k
    # an FFN pruned to nothing keeps one (dead) unit, like SparseDimensionsLinear.get_sparsity
    z = {k: v.clone() for k, v in sd.items()}
    z["vit.encoder.layer.3.intermediate.dense.weight"].zero_()
    ck.drop_zero_ffn_(z)
    assert z["vit.encoder.layer.3.intermediate.dense.weight"].shape == (1, 192)
    assert z["vit.encoder.layer.3.output.dense.weight"].abs().sum() == 0


def test_key_normalisation():
    sd = {"module.deit.embeddings.cls_token": torch.zeros(1, 1, 8), "module.classifier.weight": torch.zeros(2, 8)}
    out = normalise_keys(sd)
    assert set(out) == {"vit.embeddings.cls_token", "classifier.weight"}


def test_distilled_deit_heads_are_folded():
    """DeiTForImageClassificationWithTeacher averages two heads (cls row, distillation row): normalise_keys folds them into
    ONE classifier over the concatenated rows -- (W_c x_c + b_c + W_d x_d + b_d) / 2 == [W_c | W_d] / 2 . [x_c ; x_d] + (b_c + b_d) / 2 --
    and refuses half a pair (answering from the cls head alone would be silently wrong)."""
    import pytest
    g = torch.Generator().manual_seed(0)
    wc, wd, bc, bd = (torch.randn(5, 8, generator=g), torch.randn(5, 8, generator=g), torch.randn(5, generator=g), torch.randn(5, generator=g))
    sd = {"deit.embeddings.cls_token": torch.zeros(1, 1, 8), "cls_classifier.weight": wc, "cls_classifier.bias": bc,
          "module.distillation_classifier.weight": wd, "distillation_classifier.bias": bd}
    out = normalise_keys(sd)
    assert set(out) == {"vit.embeddings.cls_token", "classifier.weight", "classifier.bias"}
    xc, xd = torch.randn(3, 8, generator=g), torch.randn(3, 8, generator=g)
    want = ((xc @ wc.T + bc) + (xd @ wd.T + bd)) / 2
    got = torch.cat((xc, xd), 1) @ out["classifier.weight"].T + out["classifier.bias"]
    assert torch.allclose(got, want, atol=1e-6)
    for extra in ("cls_classifier.weight", "distillation_classifier.weight", "module.distillation_classifier.bias"):
        half = {"deit.embeddings.cls_token": torch.zeros(1, 1, 8), extra: torch.zeros(2, 8)}
        with pytest.raises(ValueError, match="distilled DeiT"):
            normalise_keys(half)
