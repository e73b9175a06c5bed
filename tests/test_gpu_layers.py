"""GPU parity of the modeling/torch_layers drop-ins against fixtures produced by the reference's own modules
(tests/golden/torch_layers_*.npz) and of the checkpoint-directory path."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ViTSpec  # noqa: E402
from oracle import pruning as opr  # noqa: E402
from oracle import vit as ovit  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "torch_layers_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[13:-4] for p in GOLD])
def test_torch_layers_match_reference_fixture(path):
    from edgevisiontransformer_b200 import torch_layers as tl
    f = np.load(path)
    sd = {k[3:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("sd.")}
    x = torch.from_numpy(f["x"])
    n, h = sd["layer_norm.weight"].shape
    if "num_heads" in f.files:
        m = tl.get_attention(h=h, a=int(f["num_heads"]), h_k=int(f["head_size"]), n=n)
    else:
        m = tl.get_ffn(h=h, i=sd["sub_layer.sub_layer.linear1.weight"].shape[0], n=n)
    m.load_state_dict(sd)                      # same parameter names as the reference modules
    m = m.cuda().eval()
    y = m(x.cuda())
    err = (y.cpu() - torch.from_numpy(f["y"])).abs().max().item()
    assert err < 3e-2, err                     # bf16 operands, f32 accumulate, output is LayerNorm'ed (unit scale)


def test_torch_layers_api_and_errors():
    from edgevisiontransformer_b200 import torch_layers as tl
    with pytest.raises(ValueError):
        tl.Attention(100, 3)                   # modeling/torch_layers/attention.py:8-9
    att = tl.Attention(192, 3).cuda()
    x = torch.randn(2, 50, 192, device="cuda")
    y = att(x)
    q = torch.nn.functional.linear(x, att.to_query.weight, att.to_query.bias)
    k = torch.nn.functional.linear(x, att.to_key.weight, att.to_key.bias)
    v = torch.nn.functional.linear(x, att.to_value.weight, att.to_value.bias)
    ref = torch.nn.functional.linear(ovit.attention(q, k, v, 64), att.to_out.weight, att.to_out.bias)
    assert (y - ref).abs().max().item() < 2e-2
    with torch.no_grad():
        att.to_out.bias.add_(1.0)              # in-place update must invalidate the packed copy
    assert (att(x) - (ref + 1.0)).abs().max().item() < 2e-2
    with pytest.raises(RuntimeError):
        att(x.cpu())
    ff = tl.FeedForward(192, 230).cuda()        # FFN width that is not a multiple of 8
    z = ff(x)
    refz = torch.nn.functional.linear(ovit.gelu_tanh(torch.nn.functional.linear(x, ff.linear1.weight, ff.linear1.bias)),
                                      ff.linear2.weight, ff.linear2.bias)
    assert (z - refz).abs().max().item() < 2e-2
    pre = tl.LayerNorm(192, tl.FeedForward(192, 768), is_pre=True).cuda()
    assert pre(x).shape == x.shape
    assert torch.allclose(tl.gelu(x), ovit.gelu_tanh(x), atol=1e-6)


def test_from_pretrained_pruned_checkpoint(tmp_path):
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    from edgevisiontransformer_b200 import checkpoint as ck
    sd = ovit.state_dict_of(ovit.build_hf_model(ViTSpec.deit("tiny"), seed=4, stress=True))
    heads_kept = opr.kept_heads_from_pruned_str(opr.DEIT_TINY_HEAD18, 12, 3)
    inter = [768, 231, 230, 64, 700, 8, 333, 512, 1, 96, 768, 407]
    full, pruned, to_prune = opr.synthesize_pruned(sd, heads_kept, inter, seed=7)
    ck.save_checkpoint(str(tmp_path), full, hidden_size=192, num_attention_heads=3, intermediate_size=768, pruned_heads=to_prune)
    m = B200ViTForImageClassification.from_pretrained(str(tmp_path))
    assert m.config.heads == [len(h) for h in heads_kept] and m.config.intermediate == inter
    x = ovit.synthetic_images(3, seed=1)
    want = ovit.vit_forward(pruned, ovit.spec_from_state_dict(pruned), x)
    r = ovit.compare_logits(m(x.cuda()).logits, want)
    assert r["max_abs"] <= 2e-2 and r["top1_agree"] == 1.0, r


def test_eval_loop_matches_reference_evaluate_semantics():
    """evaluate(): same counters as deit_pruning/src/utils.py:151-228 computed with the CPU oracle."""
    from edgevisiontransformer_b200 import B200ViTForImageClassification
    from edgevisiontransformer_b200.eval_loop import PipelinedClassifier, evaluate
    spec = ViTSpec.deit("tiny")
    sd = ovit.state_dict_of(ovit.build_hf_model(spec, seed=0))
    m = B200ViTForImageClassification.from_state_dict(sd)
    x = ovit.synthetic_images(23, seed=3)
    want = ovit.vit_forward(sd, spec, x)
    got = m(x.cuda()).logits.cpu()
    assert ovit.compare_logits(got, want)["max_abs"] <= 2e-2
    # near-tied random-init logits may legitimately flip under bf16 (and the small-batch kernels split K differently per
    # batch size): images without a clear margin get a label that is wrong for every implementation
    top2 = want.topk(2, dim=-1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.05
    labels = torch.where(clear, want.argmax(-1), want.argmin(-1))
    labels[::4] = (labels[::4] + 1) % 1000                    # every fourth label wrong on purpose
    wrong = torch.zeros(23, dtype=torch.bool)
    wrong[::4] = True
    expected = int((clear & ~wrong).sum())
    assert expected >= 10
    batches = [(x[i:i + 10], labels[i:i + 10]) for i in range(0, 23, 10)]
    res = evaluate(batches, m, eval_batch_size=10)
    assert abs(res["eval_accuracy"] - expected / 23.0) < 1e-9
    ref_loss = np.mean([want[i:i + 10].mean().item() for i in range(0, 23, 10)])
    assert abs(res["eval_loss"] - ref_loss) < 2e-3 and res["inference_time"] > 0
    runner = PipelinedClassifier(m, chunk=8)
    lg = runner.logits(x.pin_memory())
    assert not lg.is_cuda and ovit.compare_logits(lg, want)["max_abs"] <= 2e-2
    from edgevisiontransformer_b200 import ops
    try:
        ops.set_gemm_split_k(False)       # two separate runs agree bit for bit only without small-batch K splitting
        lg = runner.logits(x.pin_memory())
        assert torch.equal(runner.predict(x.pin_memory()), lg.argmax(-1))
    finally:
        ops.set_gemm_split_k(True)
    # submit / result: two batches in flight (the second one's copies start under the first one's forwards, staging buffers
    # and pinned result buffers alternate): every handle returns the logits of ITS batch, in any collection order
    xs = [ovit.synthetic_images(23, seed=20 + i).pin_memory() for i in range(4)]
    wants = [runner.logits(xi).clone() for xi in xs]
    h0 = runner.submit(xs[0])
    h1 = runner.submit(xs[1])
    r1, r0 = h1.result().clone(), h0.result().clone()
    h2 = runner.submit(xs[2])
    h3 = runner.submit(xs[3])
    r2, r3 = h2.result().clone(), h3.result().clone()
    for i, (got_i, want_i) in enumerate(zip((r0, r1, r2, r3), wants)):
        # the same numbers up to the chunk schedule (uniform while another batch is in flight) and the split-K reduce order ...
        assert ovit.compare_logits(got_i, want_i)["max_abs"] <= 2e-2
        # ... and nobody else's: the batches differ from each other by far more than that
        assert all(ovit.compare_logits(got_i, wants[j])["max_abs"] > 5e-2 for j in range(4) if j != i)
