"""ncu target: every hot kernel of one DeiT encoder layer (default Base; `B D heads inter` for another size), a few launches each
(no timing here -- numbers printed under a profiler are never bench values)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D, H, I = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (768, 12, 3072)   # e.g. 256 384 6 1536 = DeiT-Small
S = 197
M = B * S
x = torch.randn(M, D, device="cuda")
g, b0 = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
wqkv = (torch.randn(3 * D, D, device="cuda") * 0.02).bfloat16()
wo = (torch.randn(D, D, device="cuda") * 0.02).bfloat16()
w1 = (torch.randn(I, D, device="cuda") * 0.02).bfloat16()
w2 = (torch.randn(D, I, device="cuda") * 0.02).bfloat16()
bq, bo, b1, b2 = (torch.zeros(n, device="cuda") for n in (3 * D, D, I, D))
for _ in range(3):
    xn = ops.layernorm(x, g, b0, 1e-12)
    qkv = ops.linear(xn, wqkv, bq)
    ctx = ops.attention(qkv, B, S, H)
    ops.linear(ctx, wo, bo, residual=x, out=x, out_dtype=torch.float32)
    xn = ops.layernorm(x, g, b0, 1e-12)
    h = ops.linear(xn, w1, b1, act="gelu_erf")
    ops.linear(h, w2, b2, residual=x, out=x, out_dtype=torch.float32)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
