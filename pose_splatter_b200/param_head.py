"""The parameter-head tail between the model's MLP and render(), on the GPU in one launch each way.

Reference: src/model.py:207-257 (colour sigmoid + clip, log-scale offset `self.scale[0]`, opacity logit from the
occupancy probability, `grid[mask] + 2 * voxel_size * tanh(delta_means)`) and apply_pose_transform_3d :261-298 (yaw +
translation of the means; quaternion composition through a float64 `torch.linalg.eigh` per Gaussian, :368-421).
`gaussian_rows` returns the `[N,14]` / `[N,9]` rows `render()` / `render_views` take and is differentiable w.r.t. the
MLP output, the selected probabilities and the trainable scale offset.  Which voxels are selected (:185-204: threshold
search and random subsampling) stays with the caller.  CUDA tensors only: there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _capi


def _pose_args(angle, p_3d, row_frame, dev):
    """scalar pose -> (1, angle, host p, None, None); per-frame poses -> (1, 0, zeros, poses [F,5] on dev, row_frame)"""
    zero = (ctypes.c_float * 3)(0.0, 0.0, 0.0)
    if angle is None:
        return 0, 0.0, zero, None, None
    if row_frame is None:
        return 1, float(angle), (ctypes.c_float * 3)(*[float(x) for x in p_3d]), None, None
    ang = torch.as_tensor(angle, dtype=torch.float64).reshape(-1)
    p = torch.as_tensor(p_3d, dtype=torch.float32).reshape(-1, 3)
    # cos / sin in float64, then rounded to fp32: what torch.tensor([[c, -s, 0], ...]).to(float32) does (:276-277)
    poses = torch.cat([torch.cos(ang).float()[:, None], torch.sin(ang).float()[:, None], p], 1).contiguous().to(dev)
    return 1, 0.0, zero, poses, row_frame.to(dev, torch.int32).contiguous()


def _call_forward(mode, net_out, probs_sel, grid_sel, scale, voxel_size, pt, clip, pose_args):
    dev = net_out.device
    n, P = net_out.shape
    rows = torch.empty((n, P), dtype=torch.float32, device=dev)
    pose, angle, p_host, poses, row_frame = pose_args
    _capi.check(_capi.load().ps_param_head_forward(
        _capi.context(dev), _capi.MODE_3D if mode == "3d" else _capi.MODE_2D, n, _capi.ptr(net_out), _capi.ptr(probs_sel),
        _capi.ptr(grid_sel), _capi.ptr(scale), float(voxel_size), float(pt), float(clip[0]), float(clip[1]), pose,
        angle, p_host, _capi.ptr(poses), _capi.ptr(row_frame), 0 if poses is None else int(poses.shape[0]), _capi.ptr(rows),
        _capi.stream_ptr(dev)), "ps_param_head_forward")
    return rows


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net_out, probs_sel, scale, grid_sel, mode, voxel_size, pt, clip, angle, p_3d, row_frame):
        net_c, probs_c = net_out.detach().float().contiguous(), probs_sel.detach().float().contiguous()
        scale_c = scale.detach().float().reshape(1).contiguous()
        grid_c = None if grid_sel is None else grid_sel.detach().float().contiguous()
        pose_args = _pose_args(angle, p_3d, row_frame, net_c.device)
        rows = _call_forward(mode, net_c, probs_c, grid_c, scale_c, voxel_size, pt, clip, pose_args)
        ctx.save_for_backward(net_c, probs_c)
        ctx.meta = (mode, voxel_size, pt, clip, pose_args, scale.shape)
        return rows

    @staticmethod
    def backward(ctx, d_rows):
        net_c, probs_c = ctx.saved_tensors
        mode, voxel_size, pt, clip, (pose, angle, _, poses, row_frame), scale_shape = ctx.meta
        dev = net_c.device
        n = net_c.shape[0]
        d_rows = d_rows.float().contiguous()
        d_net, d_probs = torch.empty_like(net_c), torch.empty_like(probs_c)
        d_scale = torch.empty(1, dtype=torch.float32, device=dev)
        _capi.check(_capi.load().ps_param_head_backward(
            _capi.context(dev), _capi.MODE_3D if mode == "3d" else _capi.MODE_2D, n, _capi.ptr(net_c), _capi.ptr(probs_c),
            float(voxel_size), float(pt), float(clip[0]), float(clip[1]), pose, angle, _capi.ptr(poses), _capi.ptr(row_frame),
            0 if poses is None else int(poses.shape[0]), _capi.ptr(d_rows), _capi.ptr(d_net), _capi.ptr(d_probs), _capi.ptr(d_scale), _capi.stream_ptr(dev)),
            "ps_param_head_backward")
        return d_net, d_probs, d_scale.reshape(scale_shape), None, None, None, None, None, None, None, None


def gaussian_rows(mode: str, net_out: torch.Tensor, probs_sel: torch.Tensor, scale: torch.Tensor, voxel_size: float,
                  prob_threshold: float, grid_sel: torch.Tensor | None = None, color_clip=(0, 0.99), angle=None, p_3d=None,
                  row_frame: torch.Tensor | None = None):
    """MLP output -> gaussian_params rows.

    mode "3d": net_out [N,14] (quats 4 | scales 3 | opacity 1 | colours 3 | delta_means 3), grid_sel [N,3]; with `angle`
    (float) and `p_3d` (3 floats) the pose transform of apply_pose_transform_3d is applied as well.
    mode "2d": net_out [N,9] (means_2d 2 | scales_2d 2 | rotation 1 | colours 3 | opacity 1).
    probs_sel [N] = probs[mask]; scale = the trainable `self.scale` ([1]).
    Rows of several frames in one launch: row_frame [N] int + angle [F], p_3d [F,3] (one pose per frame).
    """
    mode = mode.lower()
    if mode not in ("2d", "3d"):
        raise ValueError(f"Unknown renderer mode: '{mode}'. Expected '2d' or '3d'.")
    P = 14 if mode == "3d" else 9
    if net_out.dim() != 2 or net_out.shape[1] != P:
        raise ValueError(f"Expected {P} parameters per Gaussian, got {tuple(net_out.shape)}")
    if net_out.device.type != "cuda":
        raise RuntimeError(f"pose_splatter_b200.param_head runs on CUDA tensors only (got {net_out.device}); there is no CPU path")
    if probs_sel.shape != (net_out.shape[0],):
        raise ValueError(f"Expected probs_sel [{net_out.shape[0]}], got {tuple(probs_sel.shape)}")
    if mode == "3d":
        if grid_sel is None or grid_sel.shape != (net_out.shape[0], 3):
            raise ValueError("3d mode needs grid_sel [N,3]")
        if (angle is None) != (p_3d is None):
            raise ValueError("angle and p_3d go together")
        if p_3d is not None and row_frame is None:
            p_3d = [float(x) for x in (p_3d.detach().cpu().reshape(-1).tolist() if isinstance(p_3d, torch.Tensor) else p_3d)]
            angle = float(angle)
        if row_frame is not None and (angle is None or row_frame.shape != (net_out.shape[0],)):
            raise ValueError("row_frame [N] needs angle [F] and p_3d [F,3]")
    else:
        grid_sel, angle, p_3d, row_frame = None, None, None, None
    return _Head.apply(net_out, probs_sel, scale, grid_sel, mode, float(voxel_size), float(prob_threshold),
                       (float(color_clip[0]), float(color_clip[1])), angle, p_3d, row_frame)


def select_voxels(volume0: torch.Tensor, mask_threshold: float, prob_threshold: float, mask_threshold_delta: float,
                  min_n: int, max_n: int, max_steps: int = 64):
    """Which voxels become Gaussians: the selection of src/model.py:185-204 with ONE device->host read.

    The reference raises `mt` by `delta` while more than `max_n` voxels pass `sigmoid(volume[0] - mt) > pt`, then lowers
    it while fewer than `min_n` pass -- a `mask.sum()` host sync per trial -- and finally subsamples at random if the
    count is still above `max_n`.  Here the counts of all trial thresholds (the same sequence of repeated float additions)
    are produced in one batched pass, read back once, and the loops are replayed on the host.  Host-side helper built
    from torch ops (no kernel of this library); returns (mask [n] bool, probs [n], mt) exactly as the reference leaves them.
    """
    pt = prob_threshold
    ups, downs = [float(mask_threshold)], []
    for _ in range(max_steps):
        ups.append(ups[-1] + mask_threshold_delta)
    # the downward walk starts wherever the upward walk stopped: tabulate it from every possible start lazily below
    trial = torch.tensor(ups, dtype=torch.float64, device=volume0.device).to(volume0.dtype)
    counts_up = (torch.sigmoid(volume0[None, :] - trial[:, None]) > pt).sum(1).tolist()  # the one host read (upward)
    k = 0
    while counts_up[k] > max_n:
        k += 1
        if k >= len(ups):
            raise RuntimeError("select_voxels: threshold search did not converge")
    mt = ups[k]
    count = counts_up[k]
    if count < min_n:
        downs = [mt]
        for _ in range(max_steps):
            downs.append(downs[-1] - mask_threshold_delta)
        trial = torch.tensor(downs, dtype=torch.float64, device=volume0.device).to(volume0.dtype)
        counts_dn = (torch.sigmoid(volume0[None, :] - trial[:, None]) > pt).sum(1).tolist()
        j = 0
        while counts_dn[j] < min_n:
            j += 1
            if j >= len(downs):
                raise RuntimeError("select_voxels: threshold search did not converge")
        mt, count = downs[j], counts_dn[j]
    probs = torch.sigmoid(volume0 - mt)
    mask = probs > pt
    if count > max_n:  # :198-203, same RNG calls as the reference
        indices = torch.nonzero(mask, as_tuple=True)[0]
        rand_idx = torch.randperm(len(indices))[:max_n].to(mask.device)
        keep = indices[rand_idx]
        mask[:] = False
        mask[keep] = True
    return mask, probs, mt
