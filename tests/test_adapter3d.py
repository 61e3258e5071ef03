"""The 3D adapter stage (src/gaussian_renderer.py:175-211 of the reference) pinned to the reference's OWN code.

gsplat is absent from the reference tree, but the lines around its call are plain torch.
tests/golden/make_golden_adapter3d.py imports the unmodified reference class with gsplat.rendering.rasterization
replaced by a recorder and stores the tensors the reference hands to gsplat plus the autograd Jacobian of those
activated values w.r.t. the raw rows.  Checked here:
  * the oracle's adapter stage (CPU) and the product's device functions (GPU probe) reproduce the activated tensors,
  * both vector-Jacobian products agree with the reference's autograd Jacobian,
  * on the GPU, rendering the raw rows equals rendering the reference's activated tensors with
    PS_FLAG_ACTIVATED_INPUTS, and raw-mode gradient == (reference Jacobian)^T x activated-mode gradient:
    everything between gaussian_params and the gsplat boundary is the reference's, only the rasterization core
    behind that boundary remains pinned to the restated gsplat rules.
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, column_rel_err
from oracle import oracle as ora


def _fixture():
    return np.load(GOLDEN / "adapter3d_reference.npz")


def test_reference_call_shape_at_the_gsplat_boundary():
    z = _fixture()
    # keyword set of src/gaussian_renderer.py:196-208: nothing else is passed, so gsplat's defaults apply
    assert list(z["keywords"]) == ["Ks", "backgrounds", "colors", "height", "means", "opacities", "packed", "quats",
                                   "scales", "viewmats", "width"]
    assert not bool(z["packed"]) and int(z["width"]) == int(z["W"]) and int(z["height"]) == int(z["H"])
    assert z["viewmats"].shape == (1, 4, 4) and z["Ks"].shape == (1, 3, 3) and z["backgrounds"].shape == (1, 3)
    assert np.array_equal(z["viewmats"][0], z["viewmat"]) and np.array_equal(z["Ks"][0], z["K"])
    assert np.array_equal(z["backgrounds"][0], z["bg"])
    assert z["opacities"].shape == (len(z["params"]),)            # sigmoid(...).squeeze(-1)
    assert tuple(z["out_rgb_shape"]) == (int(z["H"]), int(z["W"]), 3)  # rgb[0]
    assert tuple(z["out_alpha_shape"]) == (int(z["H"]), int(z["W"]))   # alpha[0, ..., 0]


def _check_act(act, z):
    want = z["act"]
    assert np.array_equal(act[:, 0:3], z["means"])                       # means pass through untouched
    assert np.array_equal(act[:, 10:13], z["colors"])                    # clamp is exact
    assert np.abs(act[:, 3:6] - z["scales"]).max() <= 2e-6 * np.abs(z["scales"]).max()
    rel = np.abs(act - want) / np.maximum(np.abs(want), 1e-6)
    assert rel.max() <= 4e-6, rel.max()                                  # exp / sigmoid / division: a few ulp


def test_oracle_adapter_matches_reference_tensors_and_jacobian():
    z = _fixture()
    rng = np.random.default_rng(3)
    v = rng.normal(size=z["act"].shape)
    act, d = ora.adapter3d(z["params"], v)
    _check_act(act, z)
    want = np.einsum("iaj,ia->ij", z["jac"].astype(np.float64), v)
    assert column_rel_err(d, want).max() <= 1e-5


@pytest.mark.gpu
def test_device_adapter_matches_reference_tensors_and_jacobian():
    from pose_splatter_b200 import _capi
    z = _fixture()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = torch.from_numpy(z["params"]).to(dev)
    n = rows.shape[0]
    g = torch.Generator().manual_seed(3)
    v = torch.randn(n, 14, generator=g)
    act = torch.empty(n, 14, device=dev)
    d = torch.empty(n, 14, device=dev)
    _capi.check(_capi.load().ps_adapter3d_probe(_capi.context(dev), _capi.ptr(rows), n, _capi.ptr(v.to(dev)), _capi.ptr(act),
                                                _capi.ptr(d), _capi.stream_ptr(dev)), "ps_adapter3d_probe")
    torch.cuda.synchronize()
    _check_act(act.cpu().numpy(), z)
    want = np.einsum("iaj,ia->ij", z["jac"].astype(np.float64), v.double().numpy())
    assert column_rel_err(d.cpu().numpy(), want).max() <= 1e-4
    # bit-exact against the oracle's restatement of the same lines
    assert np.array_equal(act.cpu().numpy().view(np.uint32), ora.adapter3d(z["params"]).view(np.uint32))


@pytest.mark.gpu
def test_raw_render_equals_reference_adapter_composed_with_activated_render():
    """raw rows -> [reference adapter, fixture] -> activated tensors -> kernels(PS_FLAG_ACTIVATED_INPUTS)  must equal
    raw rows -> kernels, forward and backward (chain rule with the reference's autograd Jacobian)."""
    from pose_splatter_b200 import _capi, batched, synth
    z = _fixture()
    dev = torch.device("cuda", torch.cuda.current_device())
    W, H = 96, 80
    n = z["params"].shape[0]
    # place the fixture's rows in front of a camera: keep its adversarial quats / colours / opacities / scales,
    # means from the synthetic model-like cloud, log-scales brought to a visible range for most rows
    rows = torch.from_numpy(z["params"]).clone()
    cloud = synth.gaussians_3d(n, 9)
    vm, Ks = synth.ring_cameras(6, ds=12.0)
    assert np.array_equal(rows.numpy(), z["params"])
    rows_np = rows.numpy()
    # activated tensors exactly as the reference produced them, but means replaced consistently on both sides
    act = torch.from_numpy(z["act"]).clone()
    rows[:, 0:3] = cloud[:, 0:3]
    act[:, 0:3] = cloud[:, 0:3]
    keep = (act[:, 3:6].max(1).values < 0.05) & (act[:, 6:10].norm(dim=1) > 0)   # drop absurd scales / the zero quaternion
    rows, act, jac = rows[keep], act[keep], torch.from_numpy(z["jac"])[keep]
    m = rows.shape[0]
    assert m >= 40
    bg = torch.tensor(z["bg"], device=dev)
    vf = torch.zeros(2, dtype=torch.int32, device=dev)
    vmd, Kd = vm[:2].to(dev).contiguous(), Ks[:2].to(dev).contiguous()
    w_rgb, w_a = synth.cotangents(2, H, W, seed=12)
    w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)
    out = {}
    for name, p, opts in (("raw", rows, None), ("act", act, dict(activated=True))):
        pd = p[None].to(dev).contiguous()
        rgb, alpha, _, saved = batched.forward_raw("3d", pd, vf, vmd, Kd, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD, False, opts)
        g = batched.backward_raw(saved, pd, vf, vmd, Kd, bg, w_rgb, w_a)
        saved.release()
        out[name] = (rgb.cpu().numpy(), alpha.cpu().numpy(), g[0].cpu().double().numpy())
    assert np.abs(out["raw"][0] - out["act"][0]).max() <= 2e-5
    assert np.abs(out["raw"][1] - out["act"][1]).max() <= 2e-5
    assert (out["raw"][1] > 0.05).mean() > 0.005, "the case must actually render something"
    chained = np.einsum("iaj,ia->ij", jac.double().numpy(), out["act"][2])
    assert column_rel_err(out["raw"][2], chained).max() <= 1e-3
    del rows_np
