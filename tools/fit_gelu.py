"""Fit the single-branch erf-GELU used by the GEMM epilogue:
     e = exp2(|v| * P(|v|)),  P(a) ~= log2(erfc(a / sqrt2)) / a  on [0, A_MAX]   (|v| clamped to A_MAX)
     gelu(v) = v > 0 ? v * (1 - 0.5 e) : 0.5 * v * e
   and report the max abs error against the exact erf-GELU, evaluated in float32 like the kernel."""
import numpy as np
from scipy.special import erfc, erf

A_MAX = 5.75
DEG = int(__import__("sys").argv[1]) if len(__import__("sys").argv) > 1 else 8
a = np.cos(np.pi * (np.arange(20000) + 0.5) / 20000) * 0.5 * A_MAX + 0.5 * A_MAX  # Chebyshev nodes on [0, A_MAX]
a = np.sort(a)
t = a / np.sqrt(2.0)
from scipy.special import log_ndtr
# log(erfc(t)) computed stably: erfc(t) = 2 * ndtr(-a)
log_erfc = np.log(2.0) + log_ndtr(-a)
P = (log_erfc / np.log(2.0)) / a
w = erfc(t) * a * np.log(2.0) * np.maximum(a, 0.05)   # d gelu = 0.5 |v| * erfc * ln2 * |v| dP
x = 2 * a / A_MAX - 1
best = None
wi = w.copy()
for it in range(60):  # Lawson iteration towards minimax of the weighted error
    c = np.polynomial.chebyshev.chebfit(x, P, DEG, w=wi)
    err = np.abs(np.polynomial.chebyshev.chebval(x, c) - P) * w
    wi = wi * (0.3 + err / err.max())
    wi /= wi.max()
poly = np.polynomial.chebyshev.cheb2poly(c)  # in x
# convert to polynomial in a: x = 2a/A - 1
pa = np.polynomial.polynomial.Polynomial(poly)(np.polynomial.polynomial.Polynomial([-1.0, 2.0 / A_MAX])).coef
coef32 = pa.astype(np.float32)
print("// P(a) coefficients, ascending powers of a (float32):")
print(", ".join("%.9ef" % v for v in coef32))


def gelu_fast(v):
    v = v.astype(np.float32)
    av = np.minimum(np.abs(v), np.float32(A_MAX))
    p = np.float32(coef32[-1])
    for cc in coef32[-2::-1]:
        p = (p * av + np.float32(cc)).astype(np.float32)
    e = np.exp2((p * av).astype(np.float32)).astype(np.float32)
    return np.where(v > 0, v * (np.float32(1) - np.float32(0.5) * e), np.float32(0.5) * v * e).astype(np.float32)


v = np.linspace(-12, 12, 2000001)
exact = 0.5 * v * (1 + erf(v / np.sqrt(2)))
got = gelu_fast(v).astype(np.float64)
d = np.abs(got - exact)
print("max abs err %.3e at v=%.4f ; max rel err (|v|>1e-3) %.3e" % (d.max(), v[d.argmax()], (d / np.maximum(np.abs(exact), 1e-30))[np.abs(v) > 1e-3].max()))
ref32 = (0.5 * v.astype(np.float32) * (1 + erf((v / np.sqrt(2)).astype(np.float32)).astype(np.float32))).astype(np.float64)
print("float32 reference formula's own error: %.3e" % np.abs(ref32 - exact).max())
