"""Warp emulation (32 lanes, shfl.bfly + PRMT + rotate / bit-select) of transpose32 in csrc/ps_raster.cu against the
definition "lane l gives row l, gets column l" -- all 1024 single-bit matrices and random ones; also the round-1 form.
Run: python tests/analysis/emulate_transpose32.py"""
import numpy as np
rng = np.random.default_rng(1)
M32 = 0xffffffff
def byte_perm(a, b, s):
    by = [(a >> (8*i)) & 0xff for i in range(4)] + [(b >> (8*i)) & 0xff for i in range(4)]
    r = 0
    for i in range(4):
        r |= by[(s >> (4*i)) & 7] << (8*i)
    return r
def rotl(y, sh):
    sh &= 31
    return ((y << sh) | (y >> (32 - sh))) & M32 if sh else y
def transpose_new(xs):
    xs = list(xs)
    def shfl(xs, J): return [xs[l ^ J] for l in range(32)]
    ys = shfl(xs, 16); xs = [byte_perm(xs[l], ys[l], 0x3276 if l & 16 else 0x5410) for l in range(32)]
    ys = shfl(xs, 8);  xs = [byte_perm(xs[l], ys[l], 0x3715 if l & 8 else 0x6240) for l in range(32)]
    for J, M in ((4, 0x0f0f0f0f), (2, 0x33333333), (1, 0x55555555)):
        ys = shfl(xs, J)
        out = []
        for l in range(32):
            up = bool(l & J)
            t = rotl(ys[l], 32 - J if up else J)
            m = (~M & M32) if up else M
            out.append((xs[l] & m) | (t & ~m & M32))
        xs = out
    return xs
def transpose_ref(xs):
    return [sum(((xs[e] >> p) & 1) << e for e in range(32)) for p in range(32)]
for _ in range(2000):
    xs = [int(v) for v in rng.integers(0, 2**32, 32, dtype=np.uint64)]
    assert transpose_new(xs) == transpose_ref(xs)
for e in range(32):
    for p in range(32):
        xs = [0]*32; xs[e] = 1 << p
        assert transpose_new(xs) == transpose_ref(xs)
print("ok")
def transpose_old(xs):
    xs = list(xs)
    for J, M in ((16, 0x0000ffff), (8, 0x00ff00ff), (4, 0x0f0f0f0f), (2, 0x33333333), (1, 0x55555555)):
        ys = [xs[l ^ J] for l in range(32)]
        xs = [((((ys[l] >> J) & M) | (xs[l] & ~M & M32)) if (l & J) else ((xs[l] & M) | ((ys[l] & M) << J))) & M32 for l in range(32)]
    return xs
for _ in range(500):
    xs = [int(v) for v in rng.integers(0, 2**32, 32, dtype=np.uint64)]
    assert transpose_old(xs) == transpose_ref(xs) == transpose_new(xs)
print("old == new == ref")
