"""Kernel micro-benchmarks on the GPU box (CUDA events on the launching stream, warm-up, L2-exceeding operands).
Development tool; bench.py holds the judged numbers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm(M, N, K, act=None, out_dtype=torch.bfloat16, residual=False, tag=""):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda", dtype=out_dtype)
    res = torch.randn(M, N, device="cuda") if residual else None
    if residual:
        out = res
    ms = timeit(lambda: ops.linear(a, w, b, act=act, residual=res, out=out, out_dtype=out_dtype))
    tf = 2.0 * M * N * K / ms / 1e9
    bytes_ = M * K * 2 + N * K * 2 + M * N * (2 if out_dtype == torch.bfloat16 else 4) * (2 if residual else 1)
    print(f"gemm {tag:8s} M={M} N={N} K={K} act={act} out={'bf16' if out_dtype == torch.bfloat16 else 'f32'} res={residual}: "
          f"{ms:.3f} ms  {tf:.0f} TFLOP/s  ({bytes_ / ms / 1e6:.0f} GB/s algorithmic)", flush=True)
    return ms


def attn(B, S, heads):
    qkv = torch.randn(B * S, 3 * heads * 64, device="cuda").bfloat16()
    ms = timeit(lambda: ops.attention(qkv, B, S, heads))
    fl = 4.0 * S * S * 64 * heads * B
    print(f"attention B={B} S={S} heads={heads}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (algorithmic)  "
          f"{(qkv.numel() * 2 + B * S * heads * 64 * 2) / ms / 1e6:.0f} GB/s", flush=True)
    return ms


def ln(rows, D):
    x = torch.randn(rows, D, device="cuda")
    g = torch.ones(D, device="cuda")
    b = torch.zeros(D, device="cuda")
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-12))
    print(f"layernorm rows={rows} D={D}: {ms:.3f} ms  {rows * D * 6 / ms / 1e6:.0f} GB/s", flush=True)
    return ms


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "base"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    D, H, I = {"base": (768, 12, 3072), "small": (384, 6, 1536), "tiny": (192, 3, 768)}[which]
    M = B * 197
    t = {}
    t["qkv"] = gemm(M, 3 * D, D, tag="qkv")
    t["attn"] = attn(B, 197, H)
    t["proj"] = gemm(M, D, D, out_dtype=torch.float32, residual=True, tag="proj")
    t["fc1"] = gemm(M, I, D, act="gelu_erf", tag="fc1")
    t["fc1_noact"] = gemm(M, I, D, act=None, tag="fc1noact")
    t["fc2"] = gemm(M, D, I, out_dtype=torch.float32, residual=True, tag="fc2")
    t["ln"] = ln(M, D)
    layer = t["qkv"] + t["attn"] + t["proj"] + t["fc1"] + t["fc2"] + 2 * t["ln"]
    print(f"layer total {layer:.3f} ms -> 12 layers {12 * layer:.2f} ms -> {B / (12 * layer) * 1e3:.0f} img/s (encoder only)")
    for k, v in t.items():
        print(f"  share {k}: {v / layer * 100:.1f}%")


if __name__ == "__main__":
    main()
