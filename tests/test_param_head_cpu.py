"""the oracle restatement of the parameter-head tail (oracle/param_head_ref.py) against fixtures generated
by running the reference's own src/model.py functions (tests/golden/make_golden_param_head.py): rows before and after
apply_pose_transform_3d, and the reference's autograd gradients mapped back through its linear head."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import param_head_ref as ref


def _load(mode):
    z = np.load(GOLDEN / f"param_head_{mode}.npz")
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def _oracle_rows(mode, z, net_out, probs, scale0, pose=True):
    if mode == "3d":
        return ref.rows_3d(net_out, probs, z["grid_sel"], scale0, z["voxel_size"], z["pt"],
                           angle=z["angle"] if pose else None, p_3d=z["p_3d"] if pose else None)
    return ref.rows_2d(net_out, probs, scale0, z["pt"])


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_rows_match_the_reference(mode):
    z = _load(mode)
    scale0 = torch.tensor([z["scale0"]])
    pre = _oracle_rows(mode, z, z["net_out"], z["probs_sel"], scale0, pose=False)
    assert (pre - z["rows_pre_pose"]).abs().max().item() <= 2e-6
    rows = _oracle_rows(mode, z, z["net_out"], z["probs_sel"], scale0)
    assert (rows - z["rows"]).abs().max().item() <= 2e-6
    if mode == "3d":  # the fixture exercises both clamps of the opacity and the colour clip
        assert (rows[:, 13] == rows[:, 13].max()).sum() >= 5 and (rows[:, 13] == rows[:, 13].min()).sum() >= 3
        assert (rows[:, 10:13] == 0.99).any()
        assert (rows[:, 6] >= 0).all() and ((rows[:, 6:10] ** 2).sum(1) - 1).abs().max() < 1e-5


@pytest.mark.parametrize("mode", ["3d", "2d"])
def test_gradients_match_the_reference_autograd(mode):
    z = _load(mode)
    net_out = z["net_out"].clone().requires_grad_(True)
    probs = z["probs_sel"].clone().requires_grad_(True)
    scale0 = torch.tensor([z["scale0"]], requires_grad=True)
    rows = _oracle_rows(mode, z, net_out, probs, scale0)
    (rows * z["cot"]).sum().backward()
    d_W1 = z["x"].T @ net_out.grad
    err = (d_W1 - z["d_W1"]).abs().max(0).values / z["d_W1"].abs().max(0).values.clamp_min(1e-20)
    assert err.max().item() <= 1e-3, err
    sel = z["sel"]
    d_vol = (z["W1"] @ net_out.grad.T)
    d_vol[0] += probs.grad * z["probs_sel"] * (1 - z["probs_sel"])
    want = z["d_volume"][:, sel]
    assert ((d_vol - want).abs().max() / want.abs().max()).item() <= 1e-3
    assert abs(scale0.grad.item() - float(z["d_scale"][0])) <= 1e-3 * abs(float(z["d_scale"][0]))


def test_degenerate_quaternion_is_the_identity_block():
    r = ref.quat_block(torch.tensor([[0.0, 0.0, 0.0, 0.0], [1e-9, 0.0, 0.0, 0.0]]))
    assert torch.equal(r, torch.eye(3).expand(2, 3, 3))


def _reference_selection(vol0, **kw):
    """the loop of src/model.py:185-204, restated for the test (thresholds, counts, random subsample)"""
    mt, pt = kw["mask_threshold"], kw["prob_threshold"]
    probs = torch.sigmoid(vol0 - mt)
    mask = probs > pt
    while mask.sum() > kw["max_n"]:
        mt += kw["mask_threshold_delta"]
        probs = torch.sigmoid(vol0 - mt)
        mask = probs > pt
    while mask.sum() < kw["min_n"]:
        mt -= kw["mask_threshold_delta"]
        probs = torch.sigmoid(vol0 - mt)
        mask = probs > pt
    if mask.sum() > kw["max_n"]:
        indices = torch.nonzero(mask, as_tuple=True)[0]
        rand_idx = torch.randperm(len(indices))[:kw["max_n"]]
        keep = indices[rand_idx]
        mask[:] = False
        mask[keep] = True
    return mask, probs, mt


@pytest.mark.parametrize("case", ["in_range", "too_many", "too_few", "overshoot_then_subsample"])
def test_select_voxels_replays_the_reference_loop(case):
    from pose_splatter_b200 import param_head
    g = torch.Generator().manual_seed(11)
    vol0 = torch.randn(20000, generator=g) * 2.0
    kw = dict(mask_threshold=0.25, prob_threshold=0.25, mask_threshold_delta=0.05, min_n=1024, max_n=16000)
    if case == "too_many":
        kw.update(max_n=3000, min_n=100)
    elif case == "too_few":
        vol0 = vol0 - 6.0
    elif case == "overshoot_then_subsample":  # a plateau: one step down jumps from < min_n to > max_n
        vol0 = torch.cat([torch.full((5000,), -1.0), torch.full((300,), 3.0)])
        kw.update(min_n=1024, max_n=2000)
    torch.manual_seed(5)
    want_mask, want_probs, want_mt = _reference_selection(vol0.clone(), **kw)
    torch.manual_seed(5)
    mask, probs, mt = param_head.select_voxels(vol0, kw["mask_threshold"], kw["prob_threshold"], kw["mask_threshold_delta"],
                                               kw["min_n"], kw["max_n"])
    assert mt == want_mt and torch.equal(mask, want_mask) and torch.equal(probs, want_probs)
    n = int(mask.sum())
    assert n <= kw["max_n"] and (n >= kw["min_n"] or case == "overshoot_then_subsample")


def test_select_voxels_matches_the_reference_fixture():
    """the fixture's selection (made by the reference's own method) is reproduced from its volume"""
    z = _load("3d")
    from pose_splatter_b200 import param_head
    vol0_grad = z["d_volume"][0]  # only used for its length
    g = torch.Generator().manual_seed(1)
    vol0 = torch.randn(len(vol0_grad), generator=g) * 2.0
    vol0[:7] = 40.0
    vol0[7:12] = 0.25 + float(np.log(0.25 / 0.75)) + 1e-7
    mask, probs, mt = param_head.select_voxels(vol0, 0.25, 0.25, 0.05, 16, 16000)
    assert torch.equal(mask, z["sel"]) and torch.allclose(probs[mask], z["probs_sel"], atol=0, rtol=0)
