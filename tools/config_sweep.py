"""Throughput and stage shares of the five BASELINE.json configs on one B200 (a measurement aid; bench.py stays the
headline).  One JSON line per config: images/s with device-resident inputs (CUDA events around `steps` forwards after
warm-up), model TFLOP/s from SURVEY.md section 8d's algorithmic FLOPs, and -- for the ViT/DeiT runtime -- the per-stage
split from the library's event tap (evt_model_profile_begin/end).

    python tools/config_sweep.py [--steps 5] [--only small,base]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import B200ViTForImageClassification  # noqa: E402
from edgevisiontransformer_b200.benchmark.b200 import _random_hf, _random_t2t_weights  # noqa: E402

GF = {"swin_tiny_bs1024": 4.5, "tiny_bs1": 2.507, "small_bs256": 9.198, "base_bs4096": 35.128, "pruned_tiny_bs1024": 0.827, "t2t14_bs1024": 9.567}


def pruned_tiny_state_dict():
    """nn_pruning 'h_0.50_d_0.3' shapes on DeiT-Tiny: 1 head and 230 FFN units per layer;
    random-init weights sliced to those shapes (values do not matter for timing)."""
    sd = {k: v.detach().clone() for k, v in _random_hf("deit_tiny").state_dict().items()}
    for l in range(12):
        p = f"vit.encoder.layer.{l}."
        for n in ("query", "key", "value"):
            sd[p + f"attention.attention.{n}.weight"] = sd[p + f"attention.attention.{n}.weight"][:64].contiguous()
            sd[p + f"attention.attention.{n}.bias"] = sd[p + f"attention.attention.{n}.bias"][:64].contiguous()
        sd[p + "attention.output.dense.weight"] = sd[p + "attention.output.dense.weight"][:, :64].contiguous()
        sd[p + "intermediate.dense.weight"] = sd[p + "intermediate.dense.weight"][:230].contiguous()
        sd[p + "intermediate.dense.bias"] = sd[p + "intermediate.dense.bias"][:230].contiguous()
        sd[p + "output.dense.weight"] = sd[p + "output.dense.weight"][:, :230].contiguous()
    return sd


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    dev = torch.device("cuda")
    cases = [
        ("tiny_bs1", "deit_tiny", 1, 1, "tf32"),
        ("small_bs256", "deit_small", 256, 256, "bf16"),
        ("base_bs4096", "deit_base", 4096, 1024, "bf16"),
        ("pruned_tiny_bs1024", "pruned", 1024, 1024, "bf16"),
        ("t2t14_bs1024", "t2t_vit_14", 1024, 1024, "bf16"),
        ("swin_tiny_bs1024", "swin_tiny", 1024, 1024, "bf16"),      # SURVEY.md section 8f rank 4 (not a BASELINE config)
    ]
    for name, kind, batch, chunk, prec in cases:
        if only and not any(o in name for o in only):
            continue
        if kind == "swin_tiny":
            from edgevisiontransformer_b200.benchmark.b200 import build_model
            model, _, _ = build_model("swin_tiny", max_batch=chunk)
            x = torch.randn(batch, 3, 224, 224, device=dev)
            tap = None
        elif kind == "t2t_vit_14":
            from edgevisiontransformer_b200.modeling_t2t import B200T2TViT
            model = B200T2TViT(_random_t2t_weights(384, 14, 6, 3.0), depth=14, num_heads=6, device=dev, max_batch=chunk)
            x = torch.randn(batch, 224, 224, 3, device=dev)
            tap = None
        else:
            if kind == "pruned":
                model = B200ViTForImageClassification.from_state_dict(pruned_tiny_state_dict(), device=dev, max_batch=chunk,
                                                                      keep_params=False)
            else:
                model = B200ViTForImageClassification.from_hf(_random_hf(kind), device=dev, max_batch=chunk, precision=prec,
                                                              keep_params=False)
            x = torch.randn(batch, 3, 224, 224, device=dev)
            tap = model
        run = (lambda: model.forward_graphed(x)) if batch == 1 else (lambda: model(x))
        ms = timed(run, args.steps if batch > 1 else 200, warmup=3 if batch > 1 else 30)
        line = {"config": name, "precision": prec, "batch": batch, "chunk": chunk, "ms_per_step": ms,
                "img_per_s": batch / ms * 1e3, "model_tflops": batch / ms * 1e3 * GF[name] / 1e3}
        if tap is not None and batch > 1:
            tap.profile_begin()
            for _ in range(args.steps):
                model(x)
            st = tap.profile_end()
            tot = sum(v for v, _ in st.values())
            line["stages"] = {k: {"share": round(v / tot, 4), "us_per_launch": round(v / max(n, 1) * 1e3, 1)} for k, (v, n) in st.items()}
        print(json.dumps(line), flush=True)
        del model, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
