"""ncu target: the row-resident residual GEMM + LayerNorm kernel (csrc/gemm_rowln.cu) on the four shapes the model runtime gives
it (pruned DeiT-Tiny batch 1024: out-proj K=64, FC2 K=230; DeiT-Small batch 256: out-proj K=384, FC2 K=1536), next to the two
kernels it replaces.  No timing here -- numbers printed under a profiler are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops  # noqa: E402

SHAPES = [(197 * 1024, 192, 64), (197 * 1024, 192, 230), (197 * 256, 384, 384), (197 * 256, 384, 1536)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for M, N, K in SHAPES:
    ld = (K + 7) // 8 * 8
    a = (torch.randn(M, ld, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, ld, device="cuda") * 0.05).bfloat16()
    bias = torch.zeros(N, device="cuda")
    g, b = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    xn = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(reps):
        ops.linear_residual_layernorm(a, w, bias, res, g, b, 1e-12, k=K, xn=xn)
    for _ in range(reps):
        ops.linear(a, w, bias, residual=res, out=res, out_dtype=torch.float32, k=K)
        ops.layernorm(res, g, b, 1e-12)
    torch.cuda.synchronize()
    print("ok", M, N, K, float(xn.float().abs().mean()))
