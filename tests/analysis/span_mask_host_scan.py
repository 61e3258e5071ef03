"""Host analysis of raster_fwd6's candidate masks (span_mask in csrc/ps_raster.cu): are the pixels the per-pixel test accepts
(sigma >= 0 && sigma <= thr + slack, ps_sigma3d) always inside the per-row spans?  The function text is extracted from the .cu,
built for the host with sqrt.approx replaced by sqrtf * (1 -/+ 2 ulp) and checked by brute force over the 32 pixels of random
blocks.  Result of the run kept in DESIGN.md section 7: no drops for sigma up to 150 px, needles included; 0.2 % of the entries
of giant needles (150..1000 px long, 0.55 px wide) lose a pixel.
    python tests/analysis/span_mask_host_scan.py"""
import ctypes
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
HEAD = '''#include <algorithm>
#include <math.h>
#include <stdint.h>
using std::max; using std::min;
#define __device__
#define __forceinline__ inline
static inline float __fdividef(float a, float b) { return a / b; }
struct float4 { float x, y, z, w; };
#include "%s"
static float g_sqrt_scale = 1.0f;
static inline float sqrt_approx(float x) { return sqrtf(x) * g_sqrt_scale; }
'''
TAIL = '''
extern "C" void run(const float *sp, int n, int bx, int by, float sqrt_scale, uint32_t *mask, uint32_t *exact)
{
    g_sqrt_scale = sqrt_scale;
    const float bxf = bx + 0.5f, byf = by + 0.5f;
    for (int i = 0; i < n; ++i) {
        const float *s = sp + 6 * (size_t)i;  // gx gy hA B hC thr
        mask[i] = span_mask(s[0], s[1], s[2], s[3], s[4], s[5] * 1.0001f + 2.0f * PS_THR_SLACK, bxf, byf);
        uint32_t m = 0;
        for (int p = 0; p < 32; ++p) {
            float dx, dy;
            const float sg = ps_sigma3d(s[0], s[1], s[2], s[3], s[4], bxf + (p & 7), byf + (p >> 3), &dx, &dy);
            if (sg >= 0.0f && sg <= s[5] + PS_THR_SLACK) m |= 1u << p;
        }
        exact[i] = m;
    }
}
'''


def build(tmp):
    src = (ROOT / "pose_splatter_b200" / "csrc" / "ps_raster.cu").read_text()
    i0 = src.index("__device__ __forceinline__ uint32_t span_mask")
    fn = src[i0:src.index("\n}\n", i0) + 3]
    cpp = Path(tmp) / "span.cpp"
    cpp.write_text(HEAD % (ROOT / "pose_splatter_b200" / "csrc" / "ps_cull.cuh") + fn + TAIL)
    so = Path(tmp) / "span.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(cpp)], check=True)
    return ctypes.CDLL(str(so))


def main():
    fp, up = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint32)
    bits = lambda x: int(np.unpackbits(x.view(np.uint8)).sum())  # noqa: E731
    with tempfile.TemporaryDirectory() as tmp:
        lib = build(tmp)
        for label, smin, smax, needle in (("ordinary, sigma 0.08..20 px", -2.5, 3.0, 1.0), ("needles x 0.02", -2.5, 3.0, 0.02),
                                          ("large, sigma 20..150 px", 3.0, 5.0, 1.0), ("large needles x 0.003", 3.0, 5.0, 0.003),
                                          ("giant needles, 150..1000 px x 0.001", 5.0, 6.9, 0.001)):
            rng = np.random.default_rng(2)
            drops = tot = kept = ex = 0
            for _ in range(8):
                n = 200000
                bx, by = int(rng.integers(0, 144)) * 8, int(rng.integers(0, 256)) * 4
                sx, sy = np.exp(rng.uniform(smin, smax, n)) * needle, np.exp(rng.uniform(smin, smax, n))
                th = rng.uniform(0, np.pi, n)
                c, s = np.cos(th), np.sin(th)
                a, b, d = c * c * sx * sx + s * s * sy * sy + 0.3, c * s * (sx * sx - sy * sy), s * s * sx * sx + c * c * sy * sy + 0.3
                det = a * d - b * b
                thr = np.log(255 * np.exp(rng.uniform(np.log(1 / 255), 0, n)))
                reach = 6 + 3.5 * np.sqrt(np.maximum(a, d))
                gx, gy = bx + 4 + rng.uniform(-1, 1, n) * reach, by + 2 + rng.uniform(-1, 1, n) * reach
                sp = np.stack([gx, gy, 0.5 * d / det, -b / det, 0.5 * a / det, thr], 1).astype(np.float32)
                for scale in (1.0 - 2.4e-7, 1.0 + 2.4e-7):
                    m, e = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
                    lib.run(sp.ctypes.data_as(fp), n, bx, by, ctypes.c_float(scale), m.ctypes.data_as(up), e.ctypes.data_as(up))
                    drops += int(((e & ~m) != 0).sum())
                    tot += int((e != 0).sum())
                    kept += bits(m)
                    ex += bits(e)
            print(f"{label}: entries with a dropped pixel {drops} of {tot}; candidate / passing pixels {kept / max(ex, 1):.3f}")


if __name__ == "__main__":
    main()
