"""Derive the polynomial coefficients used by the deterministic fp32 math contract.

Run once, by hand; the coefficients are then frozen as hex-float literals in
oracle/ps_oracle.c and pose_splatter_b200/csrc/ps_contract.cuh.  Least-squares
fit on Chebyshev nodes (close to minimax), evaluated in emulated fp32 Horner.
"""
import numpy as np

def cheb_nodes(a, b, n):
    k = np.arange(n)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (k + 0.5) / n)

def fit(f, a, b, deg, n=4000):
    x = cheb_nodes(a, b, n)
    V = np.vander(x, deg + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V, f(x), rcond=None)
    return c

def horner32(c, x):
    x = x.astype(np.float32)
    acc = np.full_like(x, np.float32(c[-1]))
    for k in range(len(c) - 2, -1, -1):
        # emulate fmaf: exact product in f64, one rounding to f32 (double rounding ignored here)
        acc = (acc.astype(np.float64) * x.astype(np.float64) + np.float64(np.float32(c[k]))).astype(np.float32)
    return acc

if __name__ == "__main__":
    xs = np.linspace(-0.5, 0.5, 2000001)
    for deg in (4, 5, 6):
        c = fit(np.exp2, -0.5, 0.5, deg)
        c[0] = 1.0
        y = horner32(c, xs)
        rel = np.abs(y.astype(np.float64) / np.exp2(xs.astype(np.float32).astype(np.float64)) - 1).max()
        print("exp2 deg", deg, "max rel err", rel)
        print("  ", [float(np.float32(v)).hex() for v in c])
