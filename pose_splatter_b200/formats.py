"""On-disk formats either side of the renderer.

* Gaussian NPZ -- the only persisted form of real `gaussian_params` in the reference
  (scripts/visualization/export_gaussian_full.py:113-137 collects, :163-178 `save_npz` writes): activated values
  `means` (centred), `quaternions` (wxyz, as the network emits them: not normalised), `scales` (= exp(log_scales)),
  `opacities` (= sigmoid(logit), [N]), `colors`, `center` [1,3], `metadata`.  `save_gaussian_npz` writes the same
  archive; `load_gaussian_npz` reads it back into the [N,14] ACTIVATED row layout that
  `render_views(..., activated=True)` / PS_FLAG_ACTIVATED_INPUTS takes, so an exported frame replays through the
  renderer (and the benchmark) without the model.
* uint8 RGBA evaluation renders: written on the GPU by `render_views_rgba8` (scripts/utils/evaluate_model.py:101-113).

Host-side only: no kernel work here.
"""
from __future__ import annotations

import numpy as np
import torch

NPZ_KEYS = ("means", "quaternions", "scales", "opacities", "colors", "center")


def activate_rows(gaussian_params: torch.Tensor):
    """Raw [N,14] rows -> (means, quats, scales, opacities, colors) exactly as the reference's legacy accessor does
    (src/model.py:300-317): scales = exp(log_scales), opacities = sigmoid(logit)[N]; quats and colours untouched."""
    if gaussian_params.dim() != 2 or gaussian_params.shape[1] != 14:
        raise ValueError(f"Expected 14 parameters per Gaussian, got {tuple(gaussian_params.shape)}")
    p = gaussian_params.detach().float().cpu()
    return p[:, 0:3], p[:, 6:10], torch.exp(p[:, 3:6]), torch.sigmoid(p[:, 13]), p[:, 10:13]


def save_gaussian_npz(filename, means, quats, scales, opacities, colors, center=None):
    """Same archive as the reference's `save_npz` (export_gaussian_full.py:163-178).  With center=None the point
    cloud is centred here like :126-128 (center = mean of the means, subtracted before saving)."""
    means = np.asarray(means, np.float32)
    if center is None:
        center = means.mean(axis=0, keepdims=True)
        means = means - center
    np.savez_compressed(filename, means=means, quaternions=np.asarray(quats, np.float32),
                        scales=np.asarray(scales, np.float32), opacities=np.asarray(opacities, np.float32),
                        colors=np.asarray(colors, np.float32), center=np.asarray(center, np.float32),
                        metadata={"format": "gaussian_splatting_full", "num_gaussians": len(means), "version": "1.0"})


def load_gaussian_npz(filename) -> torch.Tensor:
    """NPZ -> [N,14] activated rows: means + center | scales | quats | colours | opacity."""
    # allow_pickle stays off: every array this loader needs is plain float data.  The reference's `metadata` entry is a
    # pickled dict (export_gaussian_full.py:176); unpickling an archive someone else wrote can run arbitrary code, so it
    # is not read -- the required keys identify the format.
    with np.load(filename, allow_pickle=False) as z:
        missing = [k for k in NPZ_KEYS if k not in z.files]
        if missing:
            raise ValueError(f"{filename}: not a gaussian_splatting_full archive (missing {missing})")
        means = z["means"].astype(np.float32) + z["center"].astype(np.float32).reshape(1, 3)
        n = len(means)
        cols = [means, z["scales"].astype(np.float32).reshape(n, 3), z["quaternions"].astype(np.float32).reshape(n, 4),
                z["colors"].astype(np.float32).reshape(n, 3), z["opacities"].astype(np.float32).reshape(n, 1)]
    return torch.from_numpy(np.concatenate(cols, axis=1))
