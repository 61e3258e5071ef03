"""Seeded random sweep of shapes and options through the full bit-exact comparison (tests/test_gpu_parity._compare):
odd image sizes (not multiples of 4 or 16, down to 1x1), N from 1 to a few thousand, uneven view->frame maps,
radius_clip, backgrounds, both modes."""
import numpy as np
import pytest
import torch

from test_gpu_parity import _compare, _mods

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(12)))
def test_fuzz_3d(seed):
    _, _, _, synth = _mods()
    rng = np.random.default_rng(100 + seed)
    W, H = int(rng.choice([1, 5, 17, 33, 100, 145, 288])), int(rng.choice([1, 3, 16, 31, 64, 129, 256]))
    N = int(rng.choice([1, 2, 31, 257, 1000, 2500]))
    F, V = int(rng.integers(1, 4)), int(rng.integers(1, 6))
    vm, Ks = synth.ring_cameras(6, ds=1152.0 / max(W, 16))
    cams = rng.integers(0, 6, V)
    p = torch.stack([synth.gaussians_3d(N, 7 * seed + f) for f in range(F)])
    p[:, :, 3:6] += float(rng.uniform(0.0, 1.5))
    vf = torch.from_numpy(rng.integers(0, F, V).astype(np.int32))
    bg = tuple(float(x) for x in rng.uniform(0, 1, 3))
    _compare("3d", p, vf, W, H, bg, vm[cams], Ks[cams], seed_w=seed, radius_clip=float(rng.choice([0.0, 2.0])))


@pytest.mark.parametrize("seed", list(range(10)))
def test_fuzz_2d(seed):
    rng = np.random.default_rng(200 + seed)
    g = torch.Generator().manual_seed(300 + seed)
    W, H = int(rng.choice([1, 7, 18, 47, 96, 131])), int(rng.choice([1, 4, 15, 33, 80, 97]))
    N = int(rng.choice([1, 3, 40, 300, 900]))
    F = int(rng.integers(1, 3))
    V = int(rng.integers(1, 4))
    lo, hi = (0.05, 0.8) if seed % 3 == 0 else (0.5, 4.0)
    p = torch.cat([torch.rand(F, N, 1, generator=g) * (W + 10) - 5, torch.rand(F, N, 1, generator=g) * (H + 10) - 5,
                   torch.log(torch.rand(F, N, 2, generator=g) * (hi - lo) + lo), torch.rand(F, N, 1, generator=g) * 12.0 - 6.0,
                   torch.rand(F, N, 3, generator=g) * 1.4 - 0.2, torch.randn(F, N, 1, generator=g) * 3.0], 2)
    vf = torch.from_numpy(rng.integers(0, F, V).astype(np.int32))
    bg = tuple(float(x) for x in rng.uniform(0, 1, 3))
    _compare("2d", p, vf, W, H, bg, seed_w=seed)
