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
