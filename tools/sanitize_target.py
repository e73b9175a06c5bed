"""Small pass over every hot kernel for `compute-sanitizer --tool memcheck` (development tool; where the sanitizer is not
available, tests/test_gpu_guards.py checks the same shapes with sentinel-guarded output buffers)."""
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
# projection + LayerNorm with rows resident in tensor memory (gemm_rowln.cu): needs >= 74 row blocks of 256; row and K tails
for (M, N, K, copy_ln) in [(256 * 74 + 129, 192, 64, False), (256 * 74 + 1, 192, 230, True), (256 * 74 + 33, 384, 384, False),
                           (256 * 75, 384, 200, True)]:
    a = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16()
    w = (torch.randn(N, (K + 7) // 8 * 8, device="cuda") * 0.05).bfloat16()
    r = torch.randn(M, N, device="cuda")
    ops.linear_residual_layernorm(a, w, torch.randn(N, device="cuda"), r, torch.ones(N, device="cuda"), torch.zeros(N, device="cuda"),
                                  1e-5, k=K, copy_ln=copy_ln)
for (B, S, H) in [(2, 197, 3), (1, 256, 2), (3, 128, 2), (2, 16, 1)]:
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
    ops.attention(qkv, B, S, H)
    ops.attention(qkv.float(), B, S, H)
for D in (192, 768, 230):
    x = torch.randn(333, D, device="cuda")
    ops.layernorm(x, torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"), 1e-12)
for D in (64, 96):                                    # sub-warp LayerNorm, partial last warp
    x = torch.randn(101, D, device="cuda")
    ops.layernorm(x, torch.ones(D, device="cuda"), torch.zeros(D, device="cuda"), 1e-5)
# Swin kernels
import math
from edgevisiontransformer_b200.modeling_swin import attention_table, shift_mask
for (B, T_in, T_out, G, C) in [(3, 196, 196, 1, 96), (2, 784, 196, 4, 192), (1, 49, 49, 1, 96), (3, 196, 50, 1, 192), (2, 196, 49, 4, 768)]:
    x = torch.randn(B * T_in, C, device="cuda")
    idx = torch.stack([torch.randperm(T_in)[:T_out] for _ in range(G)], 1).to(torch.int32).contiguous().view(-1).cuda()
    g, b = torch.ones(G * C, device="cuda"), torch.zeros(G * C, device="cuda")
    ops.gather_layernorm(x, idx, g, b, 1e-5, B, T_in, T_out, G, copy=(G == 1))
    ops.gather_layernorm(x, idx, g, b, 1e-5, B, T_in, T_out, G, out_dtype=torch.float32)
for heads, n_win, n_tab in [(3, 9, 4), (24, 3, 1)]:
    qkv = torch.randn(n_win * 49, 3 * heads * 32, device="cuda").bfloat16()
    tab = attention_table(torch.randn(169, heads), heads, 7, shift_mask(14, 14, 7, 3) if n_tab == 4 else None).cuda()
    ops.window_attention(qkv, tab, n_win, heads)
ops.layernorm_mean_tokens(torch.randn(3 * 49, 768, device="cuda"), torch.ones(768, device="cuda"), torch.zeros(768, device="cuda"), 1e-5, 3, 49)
ops.im2col_patch(torch.randn(2, 3, 224, 224, device="cuda"), 4)
# T2T front-end kernels: partial tiles / blocks
for (B, T) in [(2, 50), (1, 257), (1, 784)]:
    kqv = (torch.randn(B * T, 192, device="cuda") * 0.5).bfloat16()
    q_, _ = torch.linalg.qr(torch.randn(64, 32))
    ops.performer(kqv, (q_.t() * math.sqrt(32)).contiguous().cuda(), B, T)
img = torch.randn(2, 224, 224, 3, device="cuda")
ops.unfold_ln_nhwc(img, 7, 4, 2, torch.ones(147, device="cuda"), torch.zeros(147, device="cuda"), ld=152)
ops.unfold_ln_nhwc(torch.randn(2, 56, 56, 64, device="cuda"), 3, 2, 1, torch.ones(576, device="cuda"), torch.zeros(576, device="cuda"))
ops.unfold_ln_nhwc(torch.randn(2, 10, 10, 5, device="cuda"), 3, 1, 1)          # run-time-generic shape
# whole model at batch 1 and 3 (token-row patch matrix, split-tile attention, pre-wait weight prefetch)
from transformers import ViTConfig, ViTForImageClassification
from edgevisiontransformer_b200 import B200ViTForImageClassification
hf = ViTForImageClassification(ViTConfig(hidden_size=192, num_hidden_layers=2, num_attention_heads=3, intermediate_size=768,
                                         num_labels=10)).eval()
m = B200ViTForImageClassification.from_hf(hf)
for bsz in (1, 3):
    m(torch.randn(bsz, 3, 224, 224, device="cuda"))
torch.cuda.synchronize()
print("sanitize target ok")
