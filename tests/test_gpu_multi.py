"""Two-GPU test of the fused gradient reduce (ps_backward_peer): run with >= 2 visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


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
    import torch.distributed as dist
    from pose_splatter_b200 import _capi, batched, synth
    from pose_splatter_b200 import dist as psd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        F, C = 3, 6
        d = synth.make_views("c2", n_frames=F, n_cams=C, seed=17, n=2000)
        W, H = d["width"], d["height"]
        mine = torch.tensor(psd.shard_views(F, C, rank, world, "view"), dtype=torch.long)
        p = d["params"].to(dev)
        vf, vm, Ks = d["view_frame"][mine].to(dev), d["viewmats"][mine].to(dev), d["Ks"][mine].to(dev)
        w_rgb, w_a = synth.cotangents(F * C, H, W, seed=4)
        w_rgb, w_a = w_rgb[mine].to(dev).contiguous(), w_a[mine].to(dev).contiguous()
        bg = torch.ones(3, device=dev)
        # baseline: local backward + NCCL all-reduce
        _, _, _, sv = batched.forward_raw("3d", p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
        want = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
        psd.reduce_frame_grads(want)
        # fused: rows pushed straight into the owner's staging buffer over peer memory
        peer = psd.PeerGradBuffers(tuple(p.shape), dev)
        for _ in range(2):  # twice: the buffers are reused step after step
            peer.begin()
            batched.backward_peer_raw(sv, p, vm, Ks, bg, w_rgb, w_a, peer.rank_ptrs, peer.rank, peer.world)
            got = peer.end()
        torch.cuda.synchronize()
        errs = []
        for k, f in enumerate(peer.owned):
            scale = want[f].abs().amax(0).clamp_min(1e-20)
            errs.append(float(((got[k] - want[f]).abs().amax(0) / scale).max()))
        np.save(os.path.join(out_dir, f"err{rank}.npy"), np.array(errs + [len(peer.owned)]))
        sv.release()
    finally:
        dist.destroy_process_group()


def test_fused_peer_gradient_reduce_matches_allreduce(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    owned = 0
    for r in range(world):
        e = np.load(tmp_path / f"err{r}.npy")
        owned += int(e[-1])
        assert e[:-1].max() < 1e-4, e  # same rows, different summation order
    assert owned == 3
