"""The B200 backend of benchmark/ (`python -m edgevisiontransformer_b200.benchmark`), the sibling of the reference's
`trt_benchmark` (utils.py:860-900, tools.py:993-1009): argument surface on the CPU, line format / --topk / op-level models
on the GPU."""
import json
import re

import pytest
import torch

from edgevisiontransformer_b200.benchmark import b200

LINE = re.compile(r"^Avg latency:\s+(\d+\.\d{3}) ms, Std:\s+(\d+\.\d{3}) ms$")    # tools.py:1009


def test_cli_surface_and_no_cpu_path(capsys):
    """Same knobs as trt_benchmark_cmd (num_runs 50, warmup_runs 20, topk) and no silent CPU fallback."""
    with pytest.raises(SystemExit):
        b200.main([])                                  # --model is required
    if not torch.cuda.is_available():
        assert b200.main(["--model", "deit_tiny"]) == 2
        assert "no CPU path" in capsys.readouterr().out
    assert set(b200.DEIT) == {"deit_tiny", "deit_small", "deit_base"}
    assert set(b200.T2T) == {"t2t_vit_7", "t2t_vit_10", "t2t_vit_12", "t2t_vit_14"}
    assert b200.T2T["t2t_vit_14"] == (384, 14, 6, 3.0)               # modeling/models/t2t_vit.py:147-148
    sd = b200._random_t2t_weights(256, 2, 4, 2.0)
    assert sd["t2t.project.kernel"].shape == (576, 256) and sd["layers.1.ffn.fc1.kernel"].shape == (256, 512)
    assert sd["pos_embedding"].shape == (197, 256) and sd["t2t.performer1.w"].shape == (32, 64)


@pytest.mark.gpu
@pytest.mark.parametrize("argv", [
    ["--model", "deit_tiny", "--batch", "1", "--precision", "tf32", "--graph", "--num_runs", "12", "--warmup_runs", "3"],
    ["--model", "deit_tiny", "--batch", "8", "--num_runs", "12", "--warmup_runs", "3", "--topk", "5"],
    ["--model", "attention", "--h", "768", "--a", "12", "--n", "197", "--batch", "4", "--num_runs", "10", "--warmup_runs", "2"],
    ["--model", "ffn", "--h", "192", "--i", "230", "--n", "128", "--batch", "4", "--num_runs", "10", "--warmup_runs", "2"],
    ["--model", "t2t_vit_7", "--batch", "2", "--num_runs", "6", "--warmup_runs", "2"],
    ["--model", "swin_tiny", "--batch", "2", "--num_runs", "6", "--warmup_runs", "2", "--graph"],
])
def test_backend_line_format(argv, capsys):
    assert b200.main(argv) == 0
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 3
    m = LINE.match(out[1])
    assert m, out[1]
    rec = json.loads(out[2])
    assert rec["model"] == argv[1] and rec["batch"] == int(argv[argv.index("--batch") + 1])
    assert abs(rec["avg_ms"] - float(m.group(1))) < 1e-3 and rec["avg_ms"] > 0 and rec["std_ms"] >= 0
    assert rec["images_per_sec"] == pytest.approx(rec["batch"] / (rec["avg_ms"] / 1e3))
    assert rec["num_runs"] == int(argv[argv.index("--num_runs") + 1])


@pytest.mark.gpu
def test_topk_keeps_the_fastest_runs_and_checkpoint_dir(tmp_path):
    """--topk averages the k fastest runs (benchmark/tensorrt/onnx_trt_test.py:103-105); a pruned checkpoint directory is a
    model name (deit_pruning/src/eval_main.py:87)."""
    m, shape, _ = b200.build_model("deit_tiny", max_batch=4)
    avg_all, _, times = b200.b200_benchmark(m, (4, *shape), num_runs=20, warmup_runs=3)
    avg_top, std_top, _ = b200.b200_benchmark(m, (4, *shape), num_runs=20, warmup_runs=3, topk=5)
    assert len(times) == 20 and avg_top <= avg_all * 1.5 and std_top >= 0
    from edgevisiontransformer_b200 import checkpoint as ck
    from edgevisiontransformer_b200.modeling_vit import normalise_keys
    sd = normalise_keys({k: v.detach().clone() for k, v in b200._random_hf("deit_tiny").state_dict().items()})
    ck.prune_heads_(sd, {l: [1, 2] for l in range(12)}, 64, n_orig=3)
    ck.save_checkpoint(str(tmp_path), sd, hidden_size=192, num_hidden_layers=12, num_attention_heads=3, intermediate_size=768,
                       pruned_heads={l: [1, 2] for l in range(12)})
    m2, shape2, desc = b200.build_model(str(tmp_path), max_batch=2)
    assert m2.config.heads == [1] * 12 and "heads=[1, 1" in desc
    avg, _, _ = b200.b200_benchmark(m2, (2, *shape2), num_runs=5, warmup_runs=2)
    assert avg > 0
