"""Sustained cuBLAS (torch) throughput on the DeiT-Base GEMM shapes, for comparison with tools/sustained_probe.py
(library reference only -- not used by the product)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sustained_probe import sustained  # noqa: E402  (runs the libevt probe first when imported as a script)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M, D, I = B * 197, 768, 3072
x = torch.randn(M, D, device="cuda").bfloat16()
h = torch.randn(M, I, device="cuda").bfloat16()
wqkv = (torch.randn(3 * D, D, device="cuda") * 0.02).bfloat16()
w1 = (torch.randn(I, D, device="cuda") * 0.02).bfloat16()
w2 = (torch.randn(D, I, device="cuda") * 0.02).bfloat16()
bq = torch.zeros(3 * D, device="cuda").bfloat16()
b1 = torch.zeros(I, device="cuda").bfloat16()
b2 = torch.zeros(D, device="cuda").bfloat16()
oq = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
o1 = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
o2 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
sustained("cublas qkv", lambda: torch.addmm(bq, x, wqkv.t(), out=oq), flops=2.0 * M * 3 * D * D)
sustained("cublas fc1", lambda: torch.addmm(b1, x, w1.t(), out=o1), flops=2.0 * M * I * D)
sustained("cublas fc2", lambda: torch.addmm(b2, h, w2.t(), out=o2), flops=2.0 * M * D * I)
a = torch.randn(8192, 8192, device="cuda").bfloat16()
b = torch.randn(8192, 8192, device="cuda").bfloat16()
c = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
sustained("cublas 8k", lambda: torch.matmul(a, b, out=c), flops=2.0 * 8192 ** 3, secs=3.0)
