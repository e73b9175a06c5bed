"""Derives the polynomials used by the erf-GELU epilogue (csrc/gemm_common.cuh: gelu_erf_pair; degree 6 for f32
output, degree 5 for bf16 output).

erfc(a / sqrt 2) = 2^(a Q(a)) on a in [0, 4 sqrt 2]; Q is a minimax fit (Lawson iteration) of log2(erfc(a/sqrt2))/a
in the log domain, so erfc keeps uniform RELATIVE accuracy (the negative tail of GELU is x/2 * erfc).
Prints the coefficients (highest degree first) and the float32 error of the resulting GELU against float64 erf."""
import numpy as np
from scipy import special

AMAX = 4.0 * np.sqrt(2.0)
DEG = 6


def target(a):
    z = a / np.sqrt(2.0)
    return (np.log(special.erfcx(z)) - z * z) / np.log(2.0) / a


def fit(deg=DEG, n=6000, iters=300):
    k = np.arange(n)
    a = np.sort((np.cos(np.pi * (k + 0.5) / n) * 0.5 + 0.5) * AMAX)
    a = a[a > 1e-6]
    A = np.vander(a, deg + 1, increasing=True) * a[:, None]
    b = target(a) * a
    w = np.ones_like(a)
    for _ in range(iters):
        sw = np.sqrt(w)
        c, *_ = np.linalg.lstsq(A * sw[:, None], b * sw, rcond=None)
        r = np.abs(A @ c - b)
        w = w * (r / r.max() + 1e-3)
        w /= w.sum()
    return c.astype(np.float32), float(r.max())


def gelu_f32(x, q):
    x = x.astype(np.float32)
    a = np.minimum(np.abs(x), np.float32(AMAX))
    acc = np.full_like(a, q[-1])
    for c in q[-2::-1]:
        acc = acc * a + c
    e = np.exp2((acc * a - np.float32(1.0)).astype(np.float32)).astype(np.float32)   # erfc / 2
    return np.maximum(x, np.float32(0.0)) - np.abs(x) * e


if __name__ == "__main__":
    x = np.linspace(-12, 12, 2_000_001)
    ref = 0.5 * x * (1.0 + special.erf(x / np.sqrt(2.0)))
    for deg in (6, 5):
        q, log_err = fit(deg)
        got = gelu_f32(x, q).astype(np.float64)
        err = np.abs(got - ref)
        rel = err / np.maximum(np.abs(ref), 1e-30)
        print(f"degree {deg} coefficients (a^{deg} .. a^0):", ", ".join(f"{v:.9e}f" for v in q[::-1]))
        print(f"  max |log2 erfc| error {log_err:.3e}; gelu max abs error {err.max():.3e} at x={x[err.argmax()]:.3f}; "
              f"max relative error for |x|<5: {rel[np.abs(x) < 5].max():.3e}")
