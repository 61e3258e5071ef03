"""Host-side multi-GPU logic on CPU: view sharding and the per-Gaussian gradient reduce, world_size 2 over gloo.

The renderer itself needs a GPU; here each rank's per-view gradients come from the CPU oracle (the checker),
so what is tested is the partitioning (pose_splatter_b200/dist.py) and the collective: when one frame's
cameras land on different ranks, the all-reduced d_params must equal the single-process gradient.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pose_splatter_b200 import dist as psd
from pose_splatter_b200 import synth


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("policy", ["frame", "view"])
def test_shard_views_is_a_partition(world, policy):
    F, C = 5, 6
    owned = [psd.shard_views(F, C, r, world, policy) for r in range(world)]
    flat = sorted(v for o in owned for v in o)
    assert flat == list(range(F * C)), "every view rendered exactly once"
    if policy == "frame":
        for o in owned:
            frames = {v // C for v in o}
            assert all(sum(1 for v in o if v // C == f) == C for f in frames), "whole frames per rank"
        assert not psd.frames_shared_across_ranks(F, C, world, policy)
    else:
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
        assert psd.frames_shared_across_ranks(F, C, world, policy) == (world > 1)


def test_shard_views_rejects_unknown_policy():
    with pytest.raises(ValueError):
        psd.shard_views(1, 6, 0, 1, "tile")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import oracle as ora
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.make_views("c2", n_frames=2, n_cams=6, seed=11, n=300)
        W, H = 72, 64
        vm, Ks = synth.ring_cameras(6, ds=16.0)
        vm, Ks = vm.repeat(2, 1, 1), Ks.repeat(2, 1, 1)
        V = len(d["view_frame"])
        w_rgb, w_a = synth.cotangents(V, H, W, seed=2)
        mine = psd.shard_views(2, 6, rank, world, "view")
        sub = ora.render_views("3d", d["params"].numpy(), d["view_frame"].numpy()[mine], W, H, np.ones(3, np.float32),
                               vm.numpy()[mine], Ks.numpy()[mine], w_rgb.numpy()[mine], w_a.numpy()[mine])
        g = torch.from_numpy(sub["d_params"]).float()
        psd.reduce_frame_grads(g)
        slowest = psd.max_over_ranks(float(rank + 1), "cpu")
        if rank == 0:
            np.savez(os.path.join(out_dir, "reduced.npz"), g=g.numpy(), slowest=slowest)
    finally:
        dist.destroy_process_group()


def test_split_frame_gradients_allreduce_gloo_world2(tmp_path):
    from oracle import oracle as ora
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "reduced.npz")
    d = synth.make_views("c2", n_frames=2, n_cams=6, seed=11, n=300)
    W, H = 72, 64
    vm, Ks = synth.ring_cameras(6, ds=16.0)
    vm, Ks = vm.repeat(2, 1, 1), Ks.repeat(2, 1, 1)
    w_rgb, w_a = synth.cotangents(12, H, W, seed=2)
    full = ora.render_views("3d", d["params"].numpy(), d["view_frame"].numpy(), W, H, np.ones(3, np.float32),
                            vm.numpy(), Ks.numpy(), w_rgb.numpy(), w_a.numpy())
    want = full["d_params"]
    scale = np.abs(want).reshape(-1, 14).max(0)
    err = (np.abs(z["g"] - want).reshape(-1, 14).max(0) / np.maximum(scale, 1e-20)).max()
    assert err < 1e-5, err
    assert float(z["slowest"]) == 2.0  # timing rule: max over ranks
