"""Diagnostic probe for the GPU box: runs the basic kernels and prints WHERE results differ, so a single
gpurun round trip gives enough information to fix a layout / descriptor bug.  Not a test, not a benchmark."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops  # noqa: E402


def describe(name, got, ref, tol):
    d = (got.float() - ref.float()).abs()
    bad = d > tol
    print(f"[{name}] shape={tuple(got.shape)} max_err={d.max().item():.4g} mean_err={d.mean().item():.4g} "
          f"bad={int(bad.sum())}/{bad.numel()} ref_absmax={ref.abs().max().item():.4g} got_absmax={got.float().abs().max().item():.4g} "
          f"nan={int(torch.isnan(got.float()).sum())}")
    if bad.any() and got.dim() == 2:
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"    bad rows: n={rows.numel()} first={rows[:12].tolist()} last={rows[-4:].tolist()}")
        print(f"    bad cols: n={cols.numel()} first={cols[:12].tolist()} last={cols[-4:].tolist()}")
        r, c = int(rows[0]), int(cols[0])
        print(f"    got[{r},{c}:{c+8}]={got[r, c:c+8].float().tolist()}")
        print(f"    ref[{r},{c}:{c+8}]={ref[r, c:c+8].float().tolist()}")
    return not bad.any()


def gemm_case(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).cuda().bfloat16()
    ref = a.float() @ w.float().t()
    out = ops.linear(a, w, None, out_dtype=torch.float32)
    torch.cuda.synchronize()
    return describe(f"gemm {M}x{N}x{K}", out, ref, 5e-3)


def attn_case(B, S, heads, seed=0):
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B * S, 3 * heads * 64, generator=g).cuda().bfloat16()
    a = heads * 64
    q, k, v = (qkv[:, i * a:(i + 1) * a].float().view(B, S, heads, 64).transpose(1, 2) for i in range(3))
    sc = (q @ k.transpose(-1, -2)) * 0.125
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, a)
    ctx = ops.attention(qkv, B, S, heads)
    torch.cuda.synchronize()
    return describe(f"attn B{B} S{S} h{heads}", ctx, ref, 2e-2)


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    steps = [
        ("gemm 128x64x64", lambda: gemm_case(128, 64, 64)),
        ("gemm 128x128x64", lambda: gemm_case(128, 128, 64)),
        ("gemm 128x256x128", lambda: gemm_case(128, 256, 128)),
        ("gemm 256x192x192", lambda: gemm_case(256, 192, 192)),
        ("gemm 197x576x192", lambda: gemm_case(197, 576, 192)),
        ("gemm 1000x768x3072", lambda: gemm_case(1000, 768, 3072)),
        ("gemm 40000x3072x768", lambda: gemm_case(40000, 3072, 768)),
        ("attn 1x128x1", lambda: attn_case(1, 128, 1)),
        ("attn 1x197x1", lambda: attn_case(1, 197, 1)),
        ("attn 2x197x3", lambda: attn_case(2, 197, 3)),
        ("attn 64x197x12", lambda: attn_case(64, 197, 12)),
    ]
    ok = True
    for name, fn in steps:
        try:
            r = fn()
            ok = ok and bool(r)
        except Exception:
            ok = False
            print(f"[{name}] EXCEPTION")
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception:
                print("context is poisoned; stopping")
                break
    print("PROBE", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
