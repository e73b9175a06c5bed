"""Small pass over every hot kernel for `compute-sanitizer --tool memcheck` (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops

torch.manual_seed(0)
for (M, N, K) in [(197, 576, 192), (300, 768, 3072), (197, 230, 192), (256 * 3 + 57, 768, 192)]:
    a = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16()
    w = (torch.randn(N, (K + 7) // 8 * 8, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda")
    for mode in (0, 1):
        ops.set_gemm_pair_mode(mode)
        o = torch.zeros(M, (N + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)
        ops.linear(a, w, b, act="gelu_erf", out=o, out_dtype=torch.bfloat16, k=K, n=N)
        if N % 4 == 0:
            r = torch.randn(M, N, device="cuda")
            ops.linear(a, w, b, residual=r, out=r, out_dtype=torch.float32, k=K)
    ops.set_gemm_pair_mode(-1)
    if N % 4 == 0:
        r = torch.randn(M, N, device="cuda")
        ops.linear(a, w, b, residual=r, out=r, out_dtype=torch.float32, k=K)     # auto: split K at small M
        ops.linear(a.float()[:, :K].contiguous(), w.float()[:, :K].contiguous(), b, out_dtype=torch.float32)   # tf32
    if N % 64 == 0:
        r = torch.randn(M, N, device="cuda")
        ops.linear_residual_layernorm(a, w, b, r, torch.ones(N, device="cuda"), torch.zeros(N, device="cuda"), 1e-12, k=K)
for (B, S, H) in [(2, 197, 3), (1, 256, 2), (3, 128, 2), (2, 16, 1)]:
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    ops.attention(qkv, B, S, H)
    ops.attention(qkv.float(), B, S, H)
for D in (192, 768, 230):
    x = torch.randn(333, D, device="cuda")
    ops.layernorm(x, torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"), 1e-12)
torch.cuda.synchronize()
print("sanitize target ok")
