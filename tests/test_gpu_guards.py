"""Out-of-bounds canaries for the kernels added last (compute-sanitizer is not available on the GPU pool): every output is
a window inside a larger sentinel-filled buffer; the bytes before and after the window must come back untouched, for shapes
with partial warps / tiles / blocks."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096          # elements on either side
SENT = -12345.0


class Guarded:
    def __init__(self, numel, dtype):
        self.buf = torch.full((numel + 2 * GUARD,), SENT, dtype=dtype, device="cuda")
        self.out = self.buf[GUARD:GUARD + numel]

    def check(self, what):
        torch.cuda.synchronize()
        assert bool((self.buf[:GUARD] == SENT).all()) and bool((self.buf[GUARD + self.out.numel():] == SENT).all()), what
        assert bool((self.out != SENT).all()), what + ": output not fully written"


def _lib():
    from edgevisiontransformer_b200 import _lib
    return _lib, _lib.load()


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("rows,D", [(101, 96), (3, 64), (197, 192), (5, 768), (33, 576)])
def test_layernorm_stays_in_bounds(rows, D):
    L, lib = _lib()
    x = torch.randn(rows, D, device="cuda")
    g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    for dt, code in ((torch.bfloat16, L.EVT_BF16), (torch.float32, L.EVT_F32)):
        y = Guarded(rows * D, dt)
        L.check(lib.evt_layernorm_fwd(x.data_ptr(), D, g.data_ptr(), b.data_ptr(), y.out.data_ptr(), code, D, None, rows, D, 1e-5, _st()))
        y.check(f"layernorm {rows}x{D} {dt}")


@pytest.mark.parametrize("M,N,K,copy_ln", [(256 * 74 + 129, 192, 64, False), (256 * 74 + 1, 192, 230, True), (256 * 74 + 33, 384, 384, False),
                                         (256 * 75 + 255, 384, 200, True)])
def test_projection_layernorm_rows_in_tmem_stays_in_bounds(M, N, K, copy_ln):
    """csrc/gemm_rowln.cu with row tails of 129 / 1 / 33 / 255 rows in the last 256-row block: the residual stream (updated in
    place through TMA stores from the ring tiles) and the normalised rows must not reach past row M; rows below M all written."""
    L, lib = _lib()
    ld = (K + 7) // 8 * 8
    a = torch.randn(M, ld, device="cuda").bfloat16()
    w = (torch.randn(N, ld, device="cuda") * 0.05).bfloat16()
    bias, g, b = torch.randn(N, device="cuda"), torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    res, xn = Guarded(M * N, torch.float32), Guarded(M * N, torch.bfloat16)
    res.out.copy_(torch.randn(M * N, device="cuda"))
    L.check(lib.evt_gemm_residual_layernorm_ex(a.data_ptr(), ld, w.data_ptr(), ld, bias.data_ptr(), res.out.data_ptr(), N, g.data_ptr(),
                                               b.data_ptr(), 1e-5, int(copy_ln), xn.out.data_ptr(), N, M, N, K, _st()))
    res.check(f"gemm+ln residual {M}x{N} K={K}")
    xn.check(f"gemm+ln normalised rows {M}x{N} K={K}")
    assert torch.isfinite(res.out).all() and torch.isfinite(xn.out.float()).all()


@pytest.mark.parametrize("B,T_in,T_out,G,C", [(3, 196, 196, 1, 96), (1, 49, 49, 1, 96), (3, 196, 50, 1, 192), (2, 784, 196, 4, 192),
                                               (2, 196, 49, 4, 768), (1, 49, 49, 1, 384)])
def test_gather_layernorm_stays_in_bounds(B, T_in, T_out, G, C):
    L, lib = _lib()
    x = torch.randn(B * T_in, C, device="cuda")
    idx = torch.stack([torch.randperm(T_in)[:T_out] for _ in range(G)], 1).to(torch.int32).contiguous().view(-1).cuda()
    g, b = torch.ones(G * C, device="cuda"), torch.zeros(G * C, device="cuda")
    n = B * T_out * G * C
    y, cp = Guarded(n, torch.bfloat16), Guarded(n, torch.float32)
    L.check(lib.evt_gather_layernorm(x.data_ptr(), idx.data_ptr(), g.data_ptr(), b.data_ptr(), y.out.data_ptr(), L.EVT_BF16,
                                     cp.out.data_ptr(), B, T_in, T_out, G, C, 1e-5, _st()))
    y.check("gather_ln y")
    cp.check("gather_ln copy")


@pytest.mark.parametrize("heads,n_win,n_tab", [(3, 9, 4), (24, 3, 1), (6, 1, 1)])
def test_window_attention_stays_in_bounds(heads, n_win, n_tab):
    from edgevisiontransformer_b200.modeling_swin import attention_table, shift_mask
    L, lib = _lib()
    qkv = torch.randn(n_win * 49, 3 * heads * 32, device="cuda").bfloat16()
    tab = attention_table(torch.randn(169, heads), heads, 7, shift_mask(14, 14, 7, 3) if n_tab == 4 else None).cuda()
    ctx = Guarded(n_win * 49 * heads * 32, torch.bfloat16)
    L.check(lib.evt_window_attention_fwd(qkv.data_ptr(), qkv.stride(0), ctx.out.data_ptr(), heads * 32, tab.data_ptr(), n_tab, n_win,
                                         49, heads, 32, 32 ** -0.5, _st()))
    ctx.check("window_attention")


@pytest.mark.parametrize("B,T", [(2, 50), (1, 257), (1, 784), (3, 16)])
def test_performer_stays_in_bounds(B, T):
    import ctypes as C
    L, lib = _lib()
    kqv = (torch.randn(B * T, 192, device="cuda") * 0.5).bfloat16()
    q_, _ = torch.linalg.qr(torch.randn(64, 32))
    w = (q_.t() * math.sqrt(32)).contiguous().cuda()
    n = C.c_size_t()
    L.check(lib.evt_performer_workspace_bytes(B, T, C.byref(n)))
    ws = Guarded(n.value // 4, torch.float32)
    y, v = Guarded(B * T * 64, torch.bfloat16), Guarded(B * T * 64, torch.float32)
    L.check(lib.evt_performer_fwd(kqv.data_ptr(), 192, w.data_ptr(), y.out.data_ptr(), v.out.data_ptr(), ws.out.data_ptr(), B, T, 64, 32,
                                  1e-8, _st()))
    y.check("performer yattn")
    v.check("performer vout")
    torch.cuda.synchronize()
    assert bool((ws.buf[:GUARD] == SENT).all()) and bool((ws.buf[GUARD + ws.out.numel():] == SENT).all())


@pytest.mark.parametrize("shape,k,s,p,ld", [((2, 224, 224, 3), 7, 4, 2, 152), ((2, 56, 56, 64), 3, 2, 1, 576), ((2, 10, 10, 5), 3, 1, 1, 46)])
def test_unfold_ln_stays_in_bounds(shape, k, s, p, ld):
    L, lib = _lib()
    x = torch.randn(*shape, device="cuda")
    B, H, W, Cc = shape
    oh, ow = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    Ln = k * k * Cc
    g, b = torch.ones(Ln, device="cuda"), torch.zeros(Ln, device="cuda")
    out = Guarded(B * oh * ow * ld, torch.bfloat16)
    L.check(lib.evt_unfold_ln_nhwc(x.data_ptr(), L.EVT_F32, out.out.data_ptr(), ld, g.data_ptr(), b.data_ptr(), 1e-5, B, H, W, Cc, k, s, p,
                                   _st()))
    out.check("unfold_ln")


def test_layernorm_mean_tokens_and_im2col4_stay_in_bounds():
    L, lib = _lib()
    x = torch.randn(3 * 49, 768, device="cuda")
    g, b = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
    y = Guarded(3 * 768, torch.bfloat16)
    L.check(lib.evt_layernorm_mean_tokens(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.out.data_ptr(), 3, 49, 768, 1e-5, _st()))
    y.check("ln_mean_tokens")
    px = torch.randn(2, 3, 56, 56, device="cuda")
    cols = Guarded(2 * 14 * 14 * 48, torch.bfloat16)
    L.check(lib.evt_im2col_patch(px.data_ptr(), cols.out.data_ptr(), 2, 56, 56, 4, _st()))
    cols.check("im2col4")
