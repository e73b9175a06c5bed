"""Times the FC1-shaped GEMM with the single-CTA and the CTA-pair kernel (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops
from tools.gemm_bench import timeit

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100864
for (N, K, act, odt, res) in [(3072, 768, "gelu_erf", torch.bfloat16, False), (2304, 768, None, torch.bfloat16, False),
                              (768, 3072, None, torch.float32, True), (768, 768, None, torch.float32, True)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda") * 0.1
    out = torch.zeros(M, N, device="cuda", dtype=odt)
    for mode in (0, 1):
        ops.set_gemm_pair_mode(mode)
        ms = timeit(lambda: ops.linear(a, w, b, act=act, residual=out if res else None, out=out, out_dtype=odt))
        print(f"M={M} N={N} K={K} act={act} pair={mode}: {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
ops.set_gemm_pair_mode(-1)

# residual GEMM + LayerNorm: unfused (reduce-add GEMM, then the LayerNorm kernel) against the fused kernel
for (N, K) in [(768, 768), (768, 3072)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda") * 0.1
    g = torch.ones(N, device="cuda")
    be = torch.zeros(N, device="cuda")
    res = torch.zeros(M, N, device="cuda")
    xn = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def unfused():
        ops.linear(a, w, b, residual=res, out=res, out_dtype=torch.float32)
        ops.layernorm(res, g, be, 1e-12)

    ms_u = timeit(unfused)
    ms_f = timeit(lambda: ops.linear_residual_layernorm(a, w, b, res, g, be, 1e-12, xn=xn))
    print(f"M={M} N={N} K={K} residual+LN: unfused {ms_u:.3f} ms, fused {ms_f:.3f} ms", flush=True)
