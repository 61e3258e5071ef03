"""GPU parity tests: the CUDA path, called through the C ABI (pose_splatter_b200._capi / batched),
against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): sort keys, tile ranges and per-pixel contributor counts
bit-exact; rendered RGB / alpha within 1e-4 max-abs; gradients within 1e-3 relative
(per parameter column, normalised by the column's max-abs gradient).
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, GRAD_TOL, RGB_TOL, bits, column_rel_err, golden_cotangents, records_from_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _mods():
    from oracle import oracle as ora
    from pose_splatter_b200 import _capi, batched, synth
    return ora, _capi, batched, synth


def _run_product(mode, params, view_frame, W, H, bg, viewmats=None, Ks=None, w_rgb=None, w_a=None, **opts):
    """forward (+ backward) through the C ABI with all debug taps kept."""
    _, _capi, batched, _ = _mods()
    p = params.to(DEV).contiguous()
    vf = view_frame.to(DEV).int()
    vm = None if viewmats is None else viewmats.to(DEV).contiguous()
    Kd = None if Ks is None else Ks.to(DEV).contiguous()
    bgd = torch.as_tensor(bg, dtype=torch.float32, device=DEV)
    rgb, alpha, counts, saved = batched.forward_raw(mode, p, vf, vm, Kd, bgd, W, H,
                                                    _capi.FLAG_SAVE_FOR_BACKWARD | _capi.FLAG_KEEP_BINNING, True, opts)
    out = dict(rgb=rgb.cpu().numpy(), alpha=alpha.cpu().numpy(), n_contrib=counts.cpu().numpy(), saved=saved)
    if w_rgb is not None:
        d = batched.backward_raw(saved, p, vf, vm, Kd, bgd, w_rgb.to(DEV).contiguous(), w_a.to(DEV).contiguous())
        out["d_params"] = d.cpu().numpy()
    return out


def _compare(mode, params, view_frame, W, H, bg, viewmats=None, Ks=None, seed_w=5, check_grad=True, **opts):
    ora, _capi, batched, synth = _mods()
    V = len(view_frame)
    w_rgb, w_a = synth.cotangents(V, H, W, seed=seed_w)
    got = _run_product(mode, params, view_frame, W, H, bg, viewmats, Ks, w_rgb, w_a, **opts)
    want = ora.render_views(mode, params.numpy(), view_frame.numpy(), W, H, np.asarray(bg, np.float32),
                            None if viewmats is None else viewmats.numpy(), None if Ks is None else Ks.numpy(),
                            w_rgb.numpy(), w_a.numpy(), **opts)
    sv = got["saved"]
    info = sv.info()
    N = params.shape[1]
    # --- projection records: bit-exact
    rec = torch.cat([sv.tap("rec0"), sv.tap("rec1"), sv.tap("rec2")], 1).cpu().numpy().reshape(V, N, 12)
    touched = sv.tap("tiles_touched").cpu().numpy().reshape(V, N)
    for v in range(V):
        r_want, _ = records_from_oracle(mode, want["views"][v]["tab"], table=True)
        assert np.array_equal(touched[v], want["views"][v]["tab"]["tiles"]), f"tiles_touched differ (view {v})"
        assert np.array_equal(bits(rec[v]), bits(r_want)), f"splat records differ (view {v})"
    if mode == "3d":
        depth = sv.tap("depth").cpu().numpy().view(np.uint32).reshape(V, N)
        for v in range(V):
            assert np.array_equal(depth[v], want["views"][v]["tab"]["low"]), f"depth words differ (view {v})"
    # --- binning: bit-exact
    assert int(info.n_isect) == len(want["keys"])
    assert np.array_equal(sv.tap("isect_keys").cpu().numpy(), want["keys"]), "sorted keys differ"
    assert np.array_equal(sv.tap("flatten_ids").cpu().numpy(), want["vals"]), "sorted values differ"
    assert np.array_equal(sv.tap("tile_offsets").cpu().numpy(), want["offsets"]), "tile ranges differ"
    # --- raster: counts bit-exact, images to tolerance
    assert np.array_equal(got["n_contrib"], want["n_contrib"]), "per-pixel contributor counts differ"
    last_want = np.stack([want["views"][v]["last"] + want["offsets"][v * (len(want["offsets"]) - 1) // V]
                          for v in range(V)])
    assert np.array_equal(sv.tap("last_ids").cpu().numpy(), last_want), "last ids differ"
    assert np.abs(got["rgb"] - want["rgb"]).max() <= RGB_TOL
    assert np.abs(got["alpha"] - want["alpha"]).max() <= RGB_TOL
    if check_grad:
        rel = column_rel_err(got["d_params"], want["d_params"])
        assert rel.max() <= GRAD_TOL, f"gradient column errors {rel}"
    return got, want


def test_math_contract_bit_exact_on_device():
    ora, _capi, _, _ = _mods()
    x = np.concatenate([np.linspace(-30, 30, 100001), np.random.default_rng(0).normal(size=50000) * 200]).astype(np.float32)
    xd = torch.from_numpy(x).to(DEV)
    y = torch.empty(5, len(x), dtype=torch.float32, device=DEV)
    dev = torch.device(DEV, torch.cuda.current_device())
    _capi.check(_capi.load().ps_math_probe(_capi.context(dev), _capi.ptr(xd), len(x), _capi.ptr(y), _capi.stream_ptr(dev)), "probe")
    torch.cuda.synchronize()
    want = ora.math_probe(x)
    y = y.cpu().numpy()
    for k, name in enumerate(("exp", "log", "sigmoid", "sin", "cos")):
        assert np.array_equal(bits(y[k]), bits(want[name])), name


@pytest.mark.parametrize("cam,N,W,H", [(0, 500, 72, 64), (3, 3000, 288, 256), (1, 777, 100, 59)])
def test_3d_single_view(cam, N, W, H):
    _, _, _, synth = _mods()
    vm, Ks = synth.ring_cameras(6, ds=1152.0 / W)
    p = synth.gaussians_3d(N, 10 + cam)[None]
    _compare("3d", p, torch.zeros(1, dtype=torch.int32), W, H, (1.0, 1.0, 1.0), vm[cam:cam + 1], Ks[cam:cam + 1])


def test_3d_dense_overlap_early_stop_and_clamp():
    _, _, _, synth = _mods()
    W, H = 96, 80
    vm, Ks = synth.ring_cameras(6, ds=12.0)
    p = synth.gaussians_3d(1500, 77)
    p[:, 3:6] += 1.0          # bigger splats: deep stacks, T falls below 1e-4
    p[:200, 13] = 9.0         # opacity ~ 1 -> alpha clamp 0.999
    got, want = _compare("3d", p[None], torch.zeros(1, dtype=torch.int32), W, H, (0.2, 0.5, 0.9), vm[2:3], Ks[2:3])
    assert (want["alpha"] > 1 - 2e-4).any(), "case must exercise the early stop"


def test_3d_batched_views_two_frames():
    _, _, _, synth = _mods()
    d = synth.make_views("c2", n_frames=2, n_cams=6, seed=3, n=1200)
    _compare("3d", d["params"], d["view_frame"], d["width"], d["height"], (1.0, 1.0, 1.0), d["viewmats"], d["Ks"])


def test_3d_adversarial_randn_eye_viewmat():
    """The one 3D render the reference tests do (tests/test_gaussian_renderer.py:207-229): randn rows,
    eye(4) viewmat, fx=fy=256: half behind the camera, some at z~0 with huge radii."""
    g = torch.Generator().manual_seed(0)
    p = torch.randn(1, 100, 14, generator=g)
    K = torch.tensor([[[256.0, 0, 128], [0, 256.0, 128], [0, 0, 1]]])
    _compare("3d", p, torch.zeros(1, dtype=torch.int32), 256, 256, (0.0, 0.0, 0.0), torch.eye(4)[None], K)


def test_3d_radius_clip_legacy_splat_options():
    _, _, _, synth = _mods()
    vm, Ks = synth.ring_cameras(6, ds=8.0)
    p = synth.gaussians_3d(800, 5)[None]
    _compare("3d", p, torch.zeros(1, dtype=torch.int32), 144, 128, (1.0, 1.0, 1.0), vm[:1], Ks[:1], radius_clip=2.0)


@pytest.mark.parametrize("name", ["ref2d_random_96x80", "ref2d_adversarial_70x50", "ref2d_c1_192x171",
                                  "ref2d_dense_small_sigma_33x47"])
def test_2d_against_reference_goldens(name):
    """CUDA 2D path vs fixtures produced by running the reference class itself (make_golden.py)."""
    z = np.load(GOLDEN / f"{name}.npz")
    W, H = int(z["W"]), int(z["H"])
    w_rgb, w_a = golden_cotangents(int(z["seed_w"]), H, W)
    got = _run_product("2d", torch.from_numpy(z["params"])[None], torch.zeros(1, dtype=torch.int32), W, H, z["bg"],
                       w_rgb=w_rgb[None], w_a=w_a[None])
    assert np.abs(got["rgb"][0] - z["rgb"]).max() <= RGB_TOL
    assert np.abs(got["alpha"][0] - z["alpha"]).max() <= RGB_TOL
    rel = column_rel_err(got["d_params"][0], z["grad"])
    assert rel.max() <= GRAD_TOL, rel


@pytest.mark.parametrize("name", ["ref2d_c3_workload_576x512_n16000", "ref2d_c3_spread_576x512_n16000",
                                  "ref2d_c3_grad_576x512_n1024", "ref2d_c5_workload_1152x1024_n4096"])
def test_2d_against_reference_goldens_full_size(name):
    """CUDA 2D path vs outputs of the unmodified reference class at the metric's sizes (make_golden_fullsize.py):
    576x512 N = 16000 (bench workload view; whole-image spread), fwd+bwd at 576x512 N = 1024, 1152x1024 N = 4096."""
    z = np.load(GOLDEN / f"{name}.npz")
    W, H = int(z["W"]), int(z["H"])
    has_grad = "grad" in z.files
    w_rgb = w_a = None
    if has_grad:
        w_rgb, w_a = golden_cotangents(int(z["seed_w"]), H, W)
        w_rgb, w_a = w_rgb[None], w_a[None]
    got = _run_product("2d", torch.from_numpy(z["params"])[None], torch.zeros(1, dtype=torch.int32), W, H, z["bg"],
                       w_rgb=w_rgb, w_a=w_a)
    assert np.abs(got["rgb"][0] - z["rgb"]).max() <= RGB_TOL
    assert np.abs(got["alpha"][0] - z["alpha"]).max() <= RGB_TOL
    if has_grad:
        rel = column_rel_err(got["d_params"][0], z["grad"])
        assert rel.max() <= GRAD_TOL, rel


@pytest.mark.parametrize("name", ["ref2d_random_96x80", "ref2d_adversarial_70x50", "ref2d_c1_192x171"])
def test_2d_against_oracle_bit_exact_binning(name):
    z = np.load(GOLDEN / f"{name}.npz")
    _compare("2d", torch.from_numpy(z["params"])[None], torch.zeros(1, dtype=torch.int32), int(z["W"]), int(z["H"]), z["bg"])


def test_2d_projected_views_batched():
    _, _, _, synth = _mods()
    d = synth.make_views("c3", n_frames=1, n_cams=3, seed=2, n=1500)
    _compare("2d", d["params"], d["view_frame"], d["width"], d["height"], (1.0, 1.0, 1.0))


def test_2d_shared_params_across_views_sums_gradients():
    """The reference model feeds one [N,9] to every camera: gradients of the views add up."""
    z = np.load(GOLDEN / "ref2d_random_96x80.npz")
    p = torch.from_numpy(z["params"])[None]
    _compare("2d", p, torch.zeros(3, dtype=torch.int32), int(z["W"]), int(z["H"]), z["bg"])


@pytest.mark.parametrize("mode,P", [("2d", 9), ("3d", 14)])
def test_empty_input_is_background(mode, P):
    _, _, batched, _ = _mods()
    bg = torch.tensor([0.25, 0.5, 0.75], device=DEV)
    p = torch.zeros(1, 0, P, device=DEV)
    rgb, alpha, counts, _ = batched.forward_raw(mode, p, torch.zeros(1, dtype=torch.int32, device=DEV),
                                                torch.eye(4, device=DEV)[None], torch.eye(3, device=DEV)[None], bg, 50, 37,
                                                0, True)
    assert torch.allclose(rgb, bg.view(1, 1, 1, 3).expand_as(rgb), atol=1e-5)
    assert float(alpha.abs().max()) == 0.0 and int(counts.max()) == 0


def test_autograd_function_matches_raw_backward():
    _, _, batched, synth = _mods()
    d = synth.make_views("c2", 1, 2, seed=4, n=600)
    p = d["params"].to(DEV).requires_grad_(True)
    bg = torch.ones(3, device=DEV)
    rgb, alpha = batched.render_views("3d", p, d["view_frame"].to(DEV), d["width"], d["height"], bg,
                                      d["viewmats"].to(DEV), d["Ks"].to(DEV))
    w_rgb, w_a = synth.cotangents(2, d["height"], d["width"], 9)
    ((rgb * w_rgb.to(DEV)).sum() + (alpha * w_a.to(DEV)).sum()).backward()
    got = _run_product("3d", d["params"], d["view_frame"], d["width"], d["height"], (1.0, 1.0, 1.0), d["viewmats"], d["Ks"],
                       w_rgb, w_a)
    rel = column_rel_err(p.grad.cpu().numpy(), got["d_params"])
    assert rel.max() < 1e-4  # same kernels; only atomic ordering differs


# ------------------------------------------------------------------------------------------
# BASELINE.json full sizes
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wl,cams", [("c2", 2), ("c3", 1), ("c1", 2), ("c5_3d", 1), ("c5_2d", 1)])
def test_full_size_workload_against_oracle(wl, cams):
    """Full N and full resolution of the benchmark configs (c2: 3D 288x256 N=16000; c3: 2D 576x512 N=16000;
    c1: 2D 192x171 N=4096; c5: 1152x1024 N=16000 in both modes): every bit-exact stage and the tolerances,
    against the oracle."""
    _, _, _, synth = _mods()
    d = synth.make_views(wl, n_frames=1, n_cams=cams, seed=21)
    _compare(d["mode"], d["params"], d["view_frame"], d["width"], d["height"], (1.0, 1.0, 1.0), d["viewmats"], d["Ks"])


def test_batch_invariance_full_size_c2():
    """Size-independent property: a view rendered inside a 24-view batch (4 frames x 6 cameras, N=16000) is
    bit-identical to the same view rendered alone, and its gradient contribution adds up linearly."""
    _, _capi, batched, synth = _mods()
    d = synth.make_views("c2", n_frames=4, n_cams=6, seed=5)
    W, H = d["width"], d["height"]
    V = len(d["view_frame"])
    p, vf = d["params"].to(DEV), d["view_frame"].to(DEV)
    vm, Ks = d["viewmats"].to(DEV), d["Ks"].to(DEV)
    bg = torch.ones(3, device=DEV)
    w_rgb, w_a = synth.cotangents(V, H, W, seed=8)
    w_rgb, w_a = w_rgb.to(DEV), w_a.to(DEV)
    rgb, alpha, counts, saved = batched.forward_raw("3d", p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD, True)
    g_all = batched.backward_raw(saved, p, vf, vm, Ks, bg, w_rgb, w_a)
    g_sum = torch.zeros_like(g_all)
    for v in (0, 7, 13, 23):
        f = int(vf[v])
        r1, a1, c1, s1 = batched.forward_raw("3d", p[f:f + 1], torch.zeros(1, dtype=torch.int32, device=DEV), vm[v:v + 1],
                                             Ks[v:v + 1], bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD, True)
        assert torch.equal(r1[0], rgb[v]) and torch.equal(a1[0], alpha[v]) and torch.equal(c1[0], counts[v])
    # gradients: the batch equals the sum over single views (atomic ordering differs -> tolerance)
    for v in range(V):
        f = int(vf[v])
        _, _, _, s1 = batched.forward_raw("3d", p[f:f + 1], torch.zeros(1, dtype=torch.int32, device=DEV), vm[v:v + 1],
                                          Ks[v:v + 1], bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
        g_sum[f] += batched.backward_raw(s1, p[f:f + 1], torch.zeros(1, dtype=torch.int32, device=DEV), vm[v:v + 1],
                                         Ks[v:v + 1], bg, w_rgb[v:v + 1].contiguous(), w_a[v:v + 1].contiguous())[0]
    rel = column_rel_err(g_all.cpu().numpy(), g_sum.cpu().numpy())
    assert rel.max() < 1e-4, rel


def test_3d_many_gaussians_global_rank_path():
    """N above the shared-memory capacity of the per-view depth ranking (16.4k) takes the global-scratch variant."""
    _, _, _, synth = _mods()
    vm, Ks = synth.ring_cameras(6, ds=12.0)
    p = synth.gaussians_3d(20000, 31)[None]
    _compare("3d", p, torch.zeros(1, dtype=torch.int32), 96, 80, (1.0, 1.0, 1.0), vm[4:5], Ks[4:5])


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_large_image_global_histogram_path(mode):
    """More than 8192 tiles per view: the per-tile histograms of projection / partition use global atomics."""
    _, _, _, synth = _mods()
    W, H = 2064, 1040  # 129 x 65 = 8385 tiles
    if mode == "3d":
        vm, Ks = synth.ring_cameras(6, ds=1152.0 / W)
        p = synth.gaussians_3d(300, 41)
        p[:, 3:6] += 1.5
        _compare("3d", p[None], torch.zeros(1, dtype=torch.int32), W, H, (0.1, 0.2, 0.3), vm[:1], Ks[:1], check_grad=True)
    else:
        g = torch.Generator().manual_seed(3)
        N = 200
        p = torch.cat([torch.rand(N, 1, generator=g) * W, torch.rand(N, 1, generator=g) * H,
                       torch.log(torch.rand(N, 2, generator=g) * 6 + 0.5), torch.rand(N, 1, generator=g) * 6.28,
                       torch.rand(N, 3, generator=g), torch.randn(N, 1, generator=g)], 1)
        _compare("2d", p[None], torch.zeros(1, dtype=torch.int32), W, H, (1.0, 1.0, 1.0))


def test_3d_activated_inputs_bit_exact():
    """PS_FLAG_ACTIVATED_INPUTS (the gsplat.rendering shim path): every bit-exact stage and the gradients."""
    _, _, _, synth = _mods()
    vm, Ks = synth.ring_cameras(6, ds=8.0)
    p = synth.gaussians_3d(900, 6)
    p[:, 3:6] = torch.exp(p[:, 3:6] + 0.5)
    p[:, 10:13] = p[:, 10:13] * 1.2
    p[:, 13] = torch.sigmoid(p[:, 13])
    _compare("3d", p[None], torch.zeros(2, dtype=torch.int32), 144, 128, (0.0, 0.0, 0.0), vm[:2], Ks[:2],
             radius_clip=2.0, activated=True)


@pytest.mark.parametrize("wl", ["c2", "c1"])
def test_rgba8_inference_output_matches_reference_quantisation(wl):
    """ps_forward_rgba8 == (255 * clip(cat(rgb, alpha), 0, 1)).astype(uint8) of the oracle's float image
    (scripts/utils/evaluate_model.py:101-113), bit for bit; tiles without splats included."""
    ora, _, batched, synth = _mods()
    d = synth.make_views(wl, n_frames=1, n_cams=2, seed=13, n=2500)
    W, H = d["width"], d["height"]
    bg = np.array([1.0, 1.0, 1.0], np.float32)
    got = batched.render_views_rgba8(d["mode"], d["params"].to(DEV), d["view_frame"].to(DEV), W, H,
                                     torch.from_numpy(bg).to(DEV), d["viewmats"].to(DEV), d["Ks"].to(DEV)).cpu().numpy()
    want = ora.render_views(d["mode"], d["params"].numpy(), d["view_frame"].numpy(), W, H, bg, d["viewmats"].numpy(),
                            d["Ks"].numpy())
    rgba = np.concatenate([want["rgb"], want["alpha"][..., None]], -1)
    want8 = (255 * rgba.clip(0, 1)).astype(np.uint8)
    assert got.shape == want8.shape and got.dtype == np.uint8
    assert np.array_equal(got, want8)


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_non_finite_rows_are_culled_like_the_oracle(mode):
    """NaN / inf / huge values in some rows: those Gaussians are dropped by the contract's finite checks on both
    sides; everything else stays bit-exact and nothing hangs."""
    _, _, _, synth = _mods()
    if mode == "3d":
        vm, Ks = synth.ring_cameras(6, ds=8.0)
        p = synth.gaussians_3d(600, 12)
        p[5, 0] = float("nan"); p[6, 4] = float("inf"); p[7, 7] = float("nan"); p[8, 13] = float("nan")
        p[9, 3:6] = 80.0       # exp overflow -> inf covariance
        p[10, 6:10] = 0.0      # zero quaternion
        p[11, 2] = 1e30
        _compare("3d", p[None], torch.zeros(1, dtype=torch.int32), 144, 128, (1.0, 1.0, 1.0), vm[:1], Ks[:1], check_grad=False)
    else:
        z = np.load(GOLDEN / "ref2d_random_96x80.npz")
        p = torch.from_numpy(z["params"]).clone()
        p[3, 0] = float("nan"); p[4, 2] = float("inf"); p[5, 4] = float("inf"); p[6, 8] = float("nan"); p[7, 3] = 90.0
        p[8, 0] = 1e30
        _compare("2d", p[None], torch.zeros(1, dtype=torch.int32), int(z["W"]), int(z["H"]), z["bg"], check_grad=False)


def test_many_views_few_gaussians():
    """2000 views of a tiny frame set: exercises view indexing beyond the benchmark's batch shapes."""
    _, _capi, batched, synth = _mods()
    vm, Ks = synth.ring_cameras(6, ds=32.0)
    F, V = 5, 2000
    p = torch.stack([synth.gaussians_3d(64, 100 + f) for f in range(F)])
    p[:, :, 3:6] += 2.0
    vf = (torch.arange(V) % F).int()
    cams = torch.arange(V) % 6
    rgb, alpha, cnt, sv = batched.forward_raw("3d", p.to(DEV), vf.to(DEV), vm[cams].to(DEV), Ks[cams].to(DEV),
                                              torch.ones(3, device=DEV), 36, 32, _capi.FLAG_SAVE_FOR_BACKWARD, True)
    # view v and view v + 30 (same frame, same camera) must be identical
    assert torch.equal(rgb[7], rgb[37]) and torch.equal(cnt[7], cnt[37]) and torch.equal(alpha[1999], alpha[1999 - 30])
    g = batched.backward_raw(sv, p.to(DEV), vf.to(DEV), vm[cams].to(DEV), Ks[cams].to(DEV), torch.ones(3, device=DEV),
                             torch.ones(V, 32, 36, 3, device=DEV), torch.ones(V, 32, 36, device=DEV))
    assert torch.isfinite(g).all() and float(g.abs().sum()) > 0


def test_render_views_vjp_equals_autograd():
    """forward + VJP without a graph == render_views + backward (same kernels, same bits)"""
    _, _, batched, synth = _mods()
    d = synth.make_views("c2", n_frames=2, n_cams=3, seed=8, n=900)
    W, H = d["width"], d["height"]
    V = len(d["view_frame"])
    w_rgb, w_a = synth.cotangents(V, H, W, seed=2)
    w_rgb, w_a = w_rgb.to(DEV), w_a.to(DEV)
    bg = torch.ones(3, device=DEV)
    p = d["params"].to(DEV).requires_grad_(True)
    rgb, alpha = batched.render_views("3d", p, d["view_frame"].to(DEV), W, H, bg, d["viewmats"].to(DEV), d["Ks"].to(DEV))
    ((rgb * w_rgb).sum() + (alpha * w_a).sum()).backward()
    rgb2, alpha2, g2 = batched.render_views_vjp("3d", d["params"].to(DEV), d["view_frame"].to(DEV), W, H, bg, w_rgb, w_a,
                                                d["viewmats"].to(DEV), d["Ks"].to(DEV))
    assert torch.equal(rgb, rgb2) and torch.equal(alpha, alpha2)
    # atomics make the accumulation order free: equal up to fp32 re-association
    rel = (p.grad - g2).abs().max() / p.grad.abs().max()
    assert rel.item() < 1e-5


def test_bad_inputs_are_rejected_or_guarded():
    """short viewmats / Ks raise on the host; a frame id outside [0, F) held on the device never reads out of
    bounds -- that view renders nothing and gets no gradient."""
    _, _capi, batched, synth = _mods()
    d = synth.make_views("c2", 1, 3, seed=2, n=500)
    W, H = d["width"], d["height"]
    p, bg = d["params"].to(DEV), torch.ones(3, device=DEV)
    with pytest.raises(ValueError, match="viewmats"):
        batched.render_views("3d", p, d["view_frame"].to(DEV), W, H, bg, d["viewmats"][:2].to(DEV), d["Ks"].to(DEV))
    with pytest.raises(ValueError, match="view_frame entries"):
        batched.render_views("3d", p, torch.tensor([0, 0, 5], dtype=torch.int32), W, H, bg, d["viewmats"].to(DEV), d["Ks"].to(DEV))
    vf = torch.tensor([0, 7, -3], dtype=torch.int32, device=DEV)  # device-side map: guarded inside the kernels
    pr = p.clone().requires_grad_(True)
    rgb, alpha = batched.render_views("3d", pr, vf, W, H, bg, d["viewmats"].to(DEV), d["Ks"].to(DEV))
    (rgb.sum() + alpha.sum()).backward()
    torch.cuda.synchronize()
    assert float(alpha[1:].abs().max()) == 0.0 and float(alpha[0].max()) > 0.0
    assert torch.isfinite(pr.grad).all()
    rgb0, alpha0 = batched.render_views("3d", p, vf[:1], W, H, bg, d["viewmats"][:1].to(DEV), d["Ks"][:1].to(DEV))
    assert torch.equal(rgb0[0], rgb[0]) and torch.equal(alpha0[0], alpha[0])


def test_3d_equal_depths_crowded_buckets_and_ties():
    """Depth ties: hundreds of Gaussians at exactly the same camera depth (identical means) and others spread out.  The
    one-pass bucket ranking gives such a view to the radix kernel (crowded bucket); either way equal depth words must keep
    Gaussian order -- the sorted keys / values are compared bit for bit with the oracle's stable sort."""
    _, _, _, synth = _mods()
    W, H = 144, 128
    vm, Ks = synth.ring_cameras(6, ds=8.0)
    p = synth.gaussians_3d(1200, 31)
    p[100:500, 0:3] = p[100, 0:3]          # 400 identical means: identical depth words in every view
    p[600:640, 0:3] = p[600, 0:3]          # a second, smaller tie group (stays below the bucket limit in most views)
    p[:, 3:6] += 0.5
    d = dict(params=p[None], view_frame=torch.zeros(3, dtype=torch.int32))
    _compare("3d", d["params"], d["view_frame"], W, H, (0.1, 0.2, 0.3), vm[:3], Ks[:3])


@pytest.mark.parametrize("wl", ["c2", "c3"])
def test_contributor_list_counters_match_the_images(wl):
    """The forward hands the backward a compact contributor list per pixel block.  Its counters must agree with what the
    forward itself composited: the backward replays exactly the forward's contributing (pixel, entry) pairs (= the sum of
    the per-pixel contributor counts), over no more entries than the forward staged, and every walked entry contributes."""
    _, _capi, batched, synth = _mods()
    d = synth.make_views(wl, 2, 6, seed=11)
    W, H, mode = d["width"], d["height"], d["mode"]
    p, vf = d["params"].to(DEV), d["view_frame"].to(DEV)
    vm = d["viewmats"].to(DEV) if d["viewmats"] is not None else None
    Ks = d["Ks"].to(DEV) if d["Ks"] is not None else None
    bg = torch.ones(3, device=DEV)
    w_rgb, w_a = synth.cotangents(len(vf), H, W, seed=7)
    _capi.raster_stats(torch.device(DEV), reset=True)
    _, _, counts, sv = batched.forward_raw(mode, p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD | _capi.FLAG_RASTER_STATS,
                                           want_counts=True)
    g = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb.to(DEV), w_a.to(DEV))
    sv.release()
    st = _capi.raster_stats(torch.device(DEV), reset=True)
    total = int(counts.sum().item())
    assert st["fwd"]["pairs_contributing"] == total > 0
    assert st["bwd"]["pairs_contributing"] == total == st["bwd"]["pairs_evaluated"]
    assert 0 < st["bwd"]["entries_walked"] <= st["bwd"]["entries_staged"] <= st["fwd"]["entries_staged"]
    assert bool(torch.isfinite(g).all())
