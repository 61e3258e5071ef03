"""CPU-side pinning of the oracle (no GPU needed).

  * 2D: oracle/ps_oracle.c against fixtures produced by RUNNING THE REFERENCE CLASS
    (tests/golden/make_golden.py; src/gaussian_renderer.py:269-427 of the reference) -- parity pinned.
  * 3D: the reference's 3D arithmetic (gsplat) is absent from its tree, so the C oracle is cross-checked
    against an independent fp64 torch restatement with autograd (oracle/ref3d_torch.py) -- parity unpinned.
  * The product's arithmetic contract header (csrc/ps_contract.cuh, host build) against the oracle's own
    restatement: bit-exact projection records, tile rectangles and pair arithmetic.
  * Structural properties of the binning (sortedness, range consistency, stable ties).
"""
import ctypes

import numpy as np
import pytest
import torch

from helpers import GOLDEN, GRAD_TOL, RGB_TOL, bits, column_rel_err, golden_cotangents, hc_project, host_contract, \
    records_from_oracle
from oracle import oracle as ora
from oracle import ref2d_dense, ref3d_torch
from pose_splatter_b200 import synth

GOLDENS_2D = ["ref2d_random_96x80", "ref2d_adversarial_70x50", "ref2d_c1_192x171", "ref2d_dense_small_sigma_33x47"]


@pytest.mark.parametrize("name", GOLDENS_2D)
def test_oracle_2d_matches_reference_fixture(name):
    z = np.load(GOLDEN / f"{name}.npz")
    W, H = int(z["W"]), int(z["H"])
    w_rgb, w_a = golden_cotangents(int(z["seed_w"]), H, W)
    o = ora.render("2d", z["params"], W, H, z["bg"], w_rgb=w_rgb.numpy(), w_a=w_a.numpy())
    assert np.abs(o["rgb"] - z["rgb"]).max() <= RGB_TOL
    assert np.abs(o["alpha"] - z["alpha"]).max() <= RGB_TOL
    rel = column_rel_err(o["d_params"], z["grad"])
    assert rel.max() <= GRAD_TOL, rel


FULLSIZE_2D = ["ref2d_c3_workload_576x512_n16000", "ref2d_c3_spread_576x512_n16000", "ref2d_c3_grad_576x512_n1024",
               "ref2d_c5_workload_1152x1024_n4096"]


@pytest.mark.parametrize("name", FULLSIZE_2D)
def test_oracle_2d_matches_reference_fixture_full_size(name):
    """Outputs of the unmodified reference class at the sizes the metric is quoted on (make_golden_fullsize.py):
    c3 576x512 with N = 16000 (the bench workload itself and a whole-image spread), fwd+bwd at 576x512 N = 1024,
    and 1152x1024 (c5) with N = 4096."""
    z = np.load(GOLDEN / f"{name}.npz")
    W, H = int(z["W"]), int(z["H"])
    if "grad" in z.files:
        w_rgb, w_a = golden_cotangents(int(z["seed_w"]), H, W)
        o = ora.render("2d", z["params"], W, H, z["bg"], w_rgb=w_rgb.numpy(), w_a=w_a.numpy())
        rel = column_rel_err(o["d_params"], z["grad"])
        assert rel.max() <= GRAD_TOL, rel
    else:
        o = ora.render("2d", z["params"], W, H, z["bg"])
    assert np.abs(o["rgb"] - z["rgb"]).max() <= RGB_TOL
    assert np.abs(o["alpha"] - z["alpha"]).max() <= RGB_TOL


def test_oracle_2d_reference_known_answers():
    """The reference's own single / two Gaussian cases (tests/test_gaussian_renderer.py:58-125): its actual rows."""
    z = np.load(GOLDEN / "ref2d_known_answers.npz")
    bg = np.zeros(3, np.float32)
    one = np.array([[128.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0]], np.float32)
    o = ora.render("2d", one, 256, 256, bg)
    assert np.abs(o["rgb"][128] - z["single_rgb_row128"]).max() <= RGB_TOL
    assert np.abs(o["alpha"][128] - z["single_alpha_row128"]).max() <= RGB_TOL
    assert o["rgb"][128, 128, 0] > 0.5 and o["rgb"][128, 128, 1] < 0.1 and o["alpha"][0, 0] < 0.1
    two = np.array([[64.0, 128.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0], [192.0, 128.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 2.0]],
                   np.float32)
    o = ora.render("2d", two, 256, 256, bg)
    assert np.abs(o["rgb"][128] - z["two_rgb_row128"]).max() <= RGB_TOL
    assert np.abs(o["alpha"][128] - z["two_alpha_row128"]).max() <= RGB_TOL


def test_oracle_2d_matches_dense_fp64_restatement():
    """Binned oracle vs the untruncated dense sum in fp64 (the tau budget of DESIGN.md section 5)."""
    z = np.load(GOLDEN / "ref2d_adversarial_70x50.npz")
    W, H = int(z["W"]), int(z["H"])
    w_rgb, w_a = golden_cotangents(int(z["seed_w"]), H, W)
    rgb, alpha, grad = ref2d_dense.render_dense_with_grad(torch.from_numpy(z["params"]).double(), W, H,
                                                          torch.from_numpy(z["bg"]).double(), w_rgb, w_a)
    o = ora.render("2d", z["params"], W, H, z["bg"], w_rgb=w_rgb.numpy(), w_a=w_a.numpy())
    assert np.abs(o["rgb"] - rgb.numpy()).max() <= 2e-5
    assert np.abs(o["alpha"] - alpha.numpy()).max() <= 2e-5
    assert column_rel_err(o["d_params"], grad.numpy()).max() <= GRAD_TOL


def test_oracle_2d_empty_and_offscreen():
    bg = np.array([0.25, 0.5, 0.75], np.float32)
    o = ora.render("2d", np.zeros((0, 9), np.float32), 40, 30, bg)
    assert np.allclose(o["rgb"], bg) and float(np.abs(o["alpha"]).max()) == 0.0 and o["binned"]["M"] == 0
    off = np.array([[-100.0, -100.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 2.0]], np.float32)
    o = ora.render("2d", off, 256, 256, bg)  # reference tests/test_gaussian_renderer.py:89-105
    assert o["alpha"].max() < 0.01


@pytest.mark.parametrize("cam,N,W,H,bump", [(0, 300, 72, 64, 0.0), (4, 260, 48, 40, 1.0)])
def test_oracle_3d_matches_fp64_torch_restatement(cam, N, W, H, bump):
    vm, Ks = synth.ring_cameras(6, ds=1152.0 / W)
    p = synth.gaussians_3d(N, 20 + cam)
    p[:, 3:6] += bump  # bigger splats -> deep stacks, alpha clamp and the T <= 1e-4 stop
    if bump:
        p[:40, 13] = 9.0
    bg = torch.tensor([0.3, 0.6, 0.9], dtype=torch.float64)
    w_rgb, w_a = synth.cotangents(1, H, W, seed=4)
    pd = p.double().requires_grad_(True)
    rgb, alpha, ncon = ref3d_torch.render(pd, vm[cam].double(), Ks[cam].double(), W, H, bg)
    ((rgb * w_rgb[0].double()).sum() + (alpha * w_a[0].double()).sum()).backward()
    o = ora.render("3d", p.numpy(), W, H, bg.float().numpy(), vm[cam].numpy(), Ks[cam].numpy(), w_rgb[0].numpy(),
                   w_a[0].numpy())
    assert np.abs(o["rgb"] - rgb.detach().numpy()).max() <= RGB_TOL
    assert np.abs(o["alpha"] - alpha.detach().numpy()).max() <= RGB_TOL
    # fp32 vs fp64 can disagree on a threshold tie (alpha >= 1/255, T <= 1e-4) for isolated pixels
    assert (o["n_contrib"] != ncon.numpy()).mean() <= 2e-3
    rel = column_rel_err(o["d_params"], pd.grad.numpy())
    assert rel.max() <= GRAD_TOL, rel


def test_oracle_3d_adversarial_reference_test_input():
    """randn rows with an identity viewmat (reference tests/test_gaussian_renderer.py:207-229): no crash,
    finite outputs, Gaussians behind the camera culled."""
    g = torch.Generator().manual_seed(0)
    p = torch.randn(100, 14, generator=g).numpy()
    K = np.array([[256.0, 0, 128], [0, 256.0, 128], [0, 0, 1]], np.float32)
    o = ora.render("3d", p, 256, 256, np.zeros(3, np.float32), np.eye(4, dtype=np.float32), K)
    assert np.isfinite(o["rgb"]).all() and np.isfinite(o["alpha"]).all()
    assert o["rgb"].shape == (256, 256, 3) and o["alpha"].shape == (256, 256)
    behind = p[:, 2] < 0.01
    assert (o["tab"]["tiles"][behind] == 0).all()


# ------------------------------------------------------------------------------------------
# product arithmetic contract (host build of csrc/ps_contract.cuh) == oracle restatement, bit for bit
# ------------------------------------------------------------------------------------------
def test_contract_math_bit_exact():
    x = np.concatenate([np.linspace(-40, 40, 20001), np.random.default_rng(1).normal(size=20000) * 300]).astype(np.float32)
    want = ora.math_probe(x)
    L = host_contract()
    fp = ctypes.POINTER(ctypes.c_float)
    outs = [np.empty_like(x) for _ in range(5)]
    L.hc_math_probe(x.ctypes.data_as(fp), len(x), *[o.ctypes.data_as(fp) for o in outs])
    for got, name in zip(outs, ("exp", "log", "sigmoid", "sin", "cos")):
        assert np.array_equal(bits(got), bits(want[name])), name
    # accuracy of the deterministic polynomials against libm (they define alpha and the extents)
    xs = np.linspace(-20, 20, 4001).astype(np.float32)
    m = ora.math_probe(xs)
    rel_exp = np.abs(m["exp"] / np.exp(xs.astype(np.float64)) - 1)
    assert rel_exp.max() < 2e-6                      # x * log2(e) is rounded once in fp32 (like __expf)
    assert rel_exp[np.abs(xs) <= 6.0].max() < 5e-7   # the range the alpha test uses (sigma <= ln 255)
    assert np.max(np.abs(m["sin"] - np.sin(xs.astype(np.float64)))) < 3e-7
    assert np.max(np.abs(m["cos"] - np.cos(xs.astype(np.float64)))) < 3e-7


@pytest.mark.parametrize("cam", [0, 3])
def test_contract_projection_3d_bit_exact(cam):
    W, H = 288, 256
    vm, Ks = synth.ring_cameras(6, ds=4.0)
    p = synth.gaussians_3d(4000, 50 + cam).numpy()
    tab = ora.project("3d", p, W, H, vm[cam].numpy(), Ks[cam].numpy())
    r, tile, low = hc_project("3d", p, W, H, vm[cam].numpy(), Ks[cam].numpy())
    want, _ = records_from_oracle("3d", tab)
    assert np.array_equal(bits(r), bits(want))
    assert np.array_equal(tile, tab["tile_rect"])
    assert np.array_equal(low, tab["low"])
    assert tab["tiles"].sum() > 0


def test_contract_projection_3d_adversarial_bit_exact():
    g = torch.Generator().manual_seed(0)
    p = torch.randn(500, 14, generator=g).numpy()
    K = np.array([[256.0, 0, 128], [0, 256.0, 128], [0, 0, 1]], np.float32)
    V = np.eye(4, dtype=np.float32)
    for clip in (0.0, 2.0):
        tab = ora.project("3d", p, 256, 256, V, K, radius_clip=clip)
        r, tile, low = hc_project("3d", p, 256, 256, V, K, clip=clip)
        want, _ = records_from_oracle("3d", tab)
        assert np.array_equal(bits(r), bits(want))
        assert np.array_equal(tile, tab["tile_rect"])


@pytest.mark.parametrize("name", GOLDENS_2D)
def test_contract_projection_2d_bit_exact(name):
    z = np.load(GOLDEN / f"{name}.npz")
    W, H = int(z["W"]), int(z["H"])
    tab = ora.project("2d", z["params"], W, H)
    r, tile, low = hc_project("2d", z["params"], W, H)
    want, _ = records_from_oracle("2d", tab)
    assert np.array_equal(bits(r), bits(want))
    assert np.array_equal(tile, tab["tile_rect"])
    assert np.array_equal(low, np.arange(len(low), dtype=np.uint32))


# ------------------------------------------------------------------------------------------
# binning definition: size-independent properties
# ------------------------------------------------------------------------------------------
def _binned_c2(cam=1, n=5000):
    vm, Ks = synth.ring_cameras(6, ds=4.0)
    p = synth.gaussians_3d(n, 9).numpy()
    tab = ora.project("3d", p, 288, 256, vm[cam].numpy(), Ks[cam].numpy())
    return tab, ora.bin_view(tab, 288, 256)


def test_binning_sorted_stable_and_ranges_consistent():
    tab, b = _binned_c2()
    keys, vals, off = b["keys"], b["vals"], b["offsets"]
    assert b["M"] == int(tab["tiles"].sum()) == len(keys)
    assert np.all(np.diff(keys) >= 0), "keys must be sorted"
    ties = np.flatnonzero(np.diff(keys) == 0)
    assert np.all(vals[ties] < vals[ties + 1]), "stable sort keeps gaussian order among equal keys"
    tw, th = ora.tile_grid(288, 256)
    assert len(off) == tw * th + 1 and off[0] == 0 and off[-1] == b["M"] and np.all(np.diff(off) >= 0)
    tile_of = (keys >> 32) & ((1 << b["tile_bits"]) - 1)
    for t in np.unique(tile_of):
        assert np.all(tile_of[off[t]:off[t + 1]] == t)
    # low word = depth bits of the listed Gaussian
    assert np.array_equal((keys & 0xffffffff).astype(np.uint32), tab["low"][vals])
    # emission order is gaussian-major, tile row-major
    uk, uv = b["unsorted"]
    assert np.all(np.diff(uv) >= 0)
    assert sorted(zip(uk.tolist(), uv.tolist())) == list(zip(keys.tolist(), vals.tolist()))


def test_tile_bits_rule():
    for (W, H), want in {(192, 171): 8, (288, 256): 9, (576, 512): 11, (1152, 1024): 13, (16, 16): 1}.items():
        assert ora.tile_bits(W, H) == want  # floor(log2(n_tiles)) + 1 (gsplat isect_tiles)


def test_gradients_are_linear_in_the_cotangents():
    """Size-independent property of the backward: d_params(a*w1 + b*w2) = a*d_params(w1) + b*d_params(w2)."""
    W, H = 72, 64
    vm, Ks = synth.ring_cameras(6, ds=16.0)
    p = synth.gaussians_3d(400, 3).numpy()
    bg = np.ones(3, np.float32)
    (w1, a1), (w2, a2) = (synth.cotangents(1, H, W, seed=s) for s in (1, 2))

    def grad(wr, wa):
        return ora.render("3d", p, W, H, bg, vm[0].numpy(), Ks[0].numpy(), wr, wa)["d_params"]

    g1, g2 = grad(w1[0].numpy(), a1[0].numpy()), grad(w2[0].numpy(), a2[0].numpy())
    g12 = grad((2 * w1[0] - 0.5 * w2[0]).numpy(), (2 * a1[0] - 0.5 * a2[0]).numpy())
    assert column_rel_err(g12, 2 * g1 - 0.5 * g2).max() < 1e-4


def _activated_rows(N, seed):
    """[N,14] rows holding ACTIVATED values (scales, raw quaternion, colours in [0,1.2], opacity) like the legacy
    PoseSplatter.splat call hands to gsplat (src/model.py:300-317,342-361)."""
    p = synth.gaussians_3d(N, seed)
    p[:, 3:6] = torch.exp(p[:, 3:6] + 0.7)
    p[:, 10:13] = p[:, 10:13] * 1.2
    p[:, 13] = torch.sigmoid(p[:, 13])
    return p


def test_oracle_3d_activated_inputs_match_fp64_torch_restatement():
    """The gsplat.rendering shim path: no adapter activations, gradients w.r.t. the activated values."""
    W, H, cam = 64, 56, 2
    vm, Ks = synth.ring_cameras(6, ds=1152.0 / W)
    p = _activated_rows(280, 5)
    bg = torch.zeros(3, dtype=torch.float64)
    w_rgb, w_a = synth.cotangents(1, H, W, seed=6)
    pd = p.double().requires_grad_(True)
    rgb, alpha, ncon = ref3d_torch.render(pd, vm[cam].double(), Ks[cam].double(), W, H, bg, radius_clip=2.0, activated=True)
    ((rgb * w_rgb[0].double()).sum() + (alpha * w_a[0].double()).sum()).backward()
    o = ora.render("3d", p.numpy(), W, H, np.zeros(3, np.float32), vm[cam].numpy(), Ks[cam].numpy(), w_rgb[0].numpy(),
                   w_a[0].numpy(), radius_clip=2.0, activated=True)
    assert np.abs(o["rgb"] - rgb.detach().numpy()).max() <= RGB_TOL
    assert np.abs(o["alpha"] - alpha.detach().numpy()).max() <= RGB_TOL
    rel = column_rel_err(o["d_params"], pd.grad.numpy())
    assert rel.max() <= GRAD_TOL, rel


def test_contract_projection_3d_activated_bit_exact():
    W, H = 144, 128
    vm, Ks = synth.ring_cameras(6, ds=8.0)
    p = _activated_rows(1500, 9).numpy()
    tab = ora.project("3d", p, W, H, vm[1].numpy(), Ks[1].numpy(), radius_clip=2.0, activated=True)
    r, tile, low = hc_project("3d", p, W, H, vm[1].numpy(), Ks[1].numpy(), clip=2.0, activated=True)
    want, _ = records_from_oracle("3d", tab)
    assert np.array_equal(bits(r), bits(want))
    assert np.array_equal(tile, tab["tile_rect"]) and np.array_equal(low, tab["low"])
