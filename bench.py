#!/usr/bin/env python
"""Headline benchmark: DeiT-Base patch16-224 bf16 inference forward, images/s, batch-sharded over N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the reference's CPU forward on the host cores)

One "step" = one forward of the hot path over the rank's shard of the global batch (BASELINE.json config 3:
global batch 4096 sharded 4096/2048/1024/512 per GPU; strong scaling, no collective on the data path).
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (hidden, heads, inter, default global batch)
    "deit_base": (768, 12, 3072, 4096),
    "deit_small": (384, 6, 1536, 256),
    "deit_tiny": (192, 3, 768, 1024),
}
GFLOP_PER_IMG = {"deit_base": 35.128, "deit_small": 9.198, "deit_tiny": 2.507}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def build_hf(workload: str, seed: int = 0):
    """Random-init HF ViTForImageClassification of the named architecture (no checkpoints offline)."""
    from transformers import ViTConfig, ViTForImageClassification
    d, h, i, _ = WORKLOADS[workload]
    cfg = ViTConfig(hidden_size=d, num_hidden_layers=12, num_attention_heads=h, intermediate_size=i, num_labels=1000,
                    image_size=224, patch_size=16, attn_implementation="eager")
    torch.manual_seed(seed)
    return ViTForImageClassification(cfg).eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_forward_rate(workload: str, batch: int, budget_s: float, threads: int):
    """images/s of the reference's own forward (HF ViT, fp32, eager attention) on the host cores."""
    torch.set_num_threads(threads)
    model = build_hf(workload)
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        model(pixel_values=x)                                   # warm-up
        t0, n = time.perf_counter(), 0
        while True:
            model(pixel_values=x)
            n += 1
            el = time.perf_counter() - t0
            if el > budget_s or n >= 50:
                break
    return batch * n / el, n, el


def run_reference(args):
    """--impl reference: the reference path (HF ViTForImageClassification called as deit_pruning/src/utils.py:194-195
    does) on the host CPU, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = args.ref_batch
    model = build_hf(args.workload)
    x = torch.randn(sample, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            model(pixel_values=x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model(pixel_values=x).logits
        el = time.perf_counter() - t0
    ips = sample * args.steps / el
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": ips, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} patch16-224, HF ViTForImageClassification forward on host CPU",
                   "sample": f"{sample} images per step"},
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": threads, "kind": "reference",
                         "sample": f"{args.steps} steps x {sample} synthetic images, fp32, eager attention, torch {torch.__version__}"},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="deit_base", choices=sorted(WORKLOADS))
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=1024, help="images per forward call inside a step")
    ap.add_argument("--ref-batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU baseline)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = max(world, 1)
    gbatch = args.global_batch or WORKLOADS[args.workload][3]
    per = gbatch // n_gpus
    assert per * n_gpus == gbatch, "global batch must divide by the GPU count"
    chunk = min(args.chunk, per)

    from edgevisiontransformer_b200 import B200ViTForImageClassification, ops
    hf = build_hf(args.workload)
    model = B200ViTForImageClassification.from_hf(hf, device=dev, max_batch=chunk, keep_params=False)
    del hf
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(per, 3, 224, 224, device=dev, generator=g)           # resident in HBM, >> L2 (126 MB)

    def step():
        return model(x).logits

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ timed region (device-resident inputs)
    barrier()
    ops.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
    launches = ops.launch_count()
    ms = e0.elapsed_time(e1)
    barrier()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = gbatch / (ms_step / 1e3)
    assert torch.isfinite(out).all()

    # ------------------------------------------------------------------ end to end through the public API, host buffers
    from edgevisiontransformer_b200.eval_loop import PipelinedClassifier
    runner = PipelinedClassifier(model, chunk=chunk)
    host = torch.empty((per, 3, 224, 224), dtype=torch.float32).pin_memory()
    host.copy_(x)
    for _ in range(2):
        runner.logits(host)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        lg = runner.logits(host)                 # H2D (pinned, chunked, overlapped) -> forward -> D2H logits
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = gbatch / (float(t.item()) / args.steps / 1e3)
    assert lg.shape == (per, 1000) and not lg.is_cuda

    # ------------------------------------------------------------------ roofline of the dominant kernel (FC1 GEMM)
    # Second pass over the SAME K steps with the library's event tap on: a CUDA event on the launching stream after
    # every launch (evt_model_profile_begin/end), so each stage's time is measured inside the step, GPU at its
    # sustained (power-capped) clocks.  `value` above comes from the untapped pass.
    pk = peaks()
    roof = None
    stage_line = None
    if rank == 0:
        d, _, inter, _ = WORKLOADS[args.workload]
        M = chunk * 197
        model.profile_begin()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        stages = model.profile_end()
        tapped_ms = e0.elapsed_time(e1)
        tot = sum(ms for ms, _ in stages.values())
        stage_line = {k: {"ms_per_launch": ms / max(n, 1), "launches": n, "share": ms / tot} for k, (ms, n) in stages.items()}
        stage_line["_tapped_step_ms"] = tapped_ms / args.steps
        fc1_ms, fc1_n = stages["fc1"]
        kms = fc1_ms / max(fc1_n, 1)
        flops = 2.0 * M * inter * d                       # algorithmic FLOPs of one FC1 launch (chunk x 197 rows)
        ach = flops / (kms / 1e3) / 1e12
        # the same kernel timed alone (burst clocks) for comparison with the burst peak
        a = torch.randn(M, d, device=dev).bfloat16()
        w = (torch.randn(inter, d, device=dev) * 0.02).bfloat16()
        b = torch.zeros(inter, device=dev)
        o = torch.empty(M, inter, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.linear(a, w, b, act="gelu_erf", out=o)
        reps = 10
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.linear(a, w, b, act="gelu_erf", out=o)
        e1.record()
        torch.cuda.synchronize()
        alone_ms = e0.elapsed_time(e1) / reps
        alone = flops / (alone_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_pair_kernel<256,bf16,gelu_erf> (FC1: M=%d N=%d K=%d)" % (M, inter, d),
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "peak_source": pk["src"] + " (sustained: kernel timed inside the step, %d launches)" % fc1_n,
                "ms_per_launch": kms, "share_of_step": fc1_ms / tot,
                "alone": {"achieved": alone, "peak": pk["tf_burst"], "frac": alone / pk["tf_burst"], "ms_per_launch": alone_ms,
                          "peak_source": pk["src"] + " (burst: kernel timed alone)"},
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, from the committed
                # ncu --set full capture profiles/r01_layer_ncu_full.md: 0.323 GB read + 1.187 GB written (algorithmic: A 0.310 + out 1.239 + W 0.005 GB)
                "traffic": 1.510e9 if (M, inter, d) == (201728, 3072, 768) else None,
                "model_frac_sustained": value / n_gpus * GFLOP_PER_IMG[args.workload] / 1e3 / pk["tf_sustained"]}
        del a, w, b, o

    # ------------------------------------------------------------------ batch-1 latency (BASELINE metric: bs1 p50)
    lat = None
    if rank == 0:
        x1 = x[:1].contiguous()
        for _ in range(30):
            model.forward_graphed(x1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(200):                      # sync-bracketed wall clock per run, as utils.py:866-872 times a model
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model.forward_graphed(x1)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        lat = {"batch": 1, "p50_ms": ts[len(ts) // 2], "p90_ms": ts[int(len(ts) * 0.9)], "runs": len(ts),
               "how": "CUDA-graph replay of the forward incl. the device-side input copy, host wall clock around each run"}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ips, n, el = cpu_forward_rate(args.workload, 16, 15.0, threads)
        cpu = {"value": ips, "unit": "img/s", "cores": threads, "kind": "reference",
               "sample": f"{n} forwards of 16 synthetic images in {el:.1f}s, HF ViTForImageClassification fp32 eager on host CPU"}

    if rank == 0:
        line = {
            "metric": "images_per_sec", "value": value, "unit": "img/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload} patch16-224 bf16 forward, random-init weights", "global_batch": gbatch,
                       "per_gpu_batch": per, "chunk": chunk, "parallelism": f"batch-sharded x{n_gpus}, no collective",
                       "l2": "inputs larger than L2 (%.0f MB of pixels per step)" % (per * 3 * 224 * 224 * 4 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": per * 3 * 224 * 224 * 4,
                    "d2h_bytes_per_step": per * 1000 * 4},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roof,
            "stages": stage_line,
            "latency": lat,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
