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


@pytest.mark.parametrize("cull_rank", [None, 1])
def test_peer_push_protocol_two_ranks_emulated_on_one_gpu(cull_rank):
    """ps_backward_peer / ps_peer_sum only see pointers: two 'ranks' sharing ONE GPU exercise the whole protocol (slot
    addressing, owner = frame % world, every slot rewritten every step) on a single-GPU box.  cull_rank = 1: that rank's
    views all look away from the scene (M == 0) -- it must still push zeros, not fail and not leave a stale slot."""
    from pose_splatter_b200 import _capi, batched, synth
    from pose_splatter_b200 import dist as psd
    dev = torch.device("cuda", torch.cuda.current_device())
    world, F, C = 2, 3, 6
    d = synth.make_views("c2", n_frames=F, n_cams=C, seed=23, n=1500)
    W, H = d["width"], d["height"]
    p = d["params"].to(dev)
    N, P = p.shape[1], p.shape[2]
    bg = torch.ones(3, device=dev)
    w_rgb_all, w_a_all = synth.cotangents(F * C, H, W, seed=6)
    fpr = (F + world - 1) // world
    stages = [torch.full((world, fpr, N, P), float("nan"), device=dev) for _ in range(world)]  # stale garbage must vanish
    ptrs = torch.tensor([s.data_ptr() for s in stages], dtype=torch.int64, device=dev)
    want = torch.zeros_like(p)
    for step in range(2):
        for rank in range(world):
            mine = torch.tensor(psd.shard_views(F, C, rank, world, "view"), dtype=torch.long)
            vf, vm, Ks = d["view_frame"][mine].to(dev), d["viewmats"][mine].clone(), d["Ks"][mine].to(dev)
            if cull_rank == rank:
                vm[:, 2, :] *= -1.0  # camera looks the other way: every Gaussian is behind the near plane
            vm = vm.to(dev)
            w_rgb, w_a = w_rgb_all[mine].to(dev).contiguous(), w_a_all[mine].to(dev).contiguous()
            _, _, _, sv = batched.forward_raw("3d", p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
            if cull_rank == rank:
                assert int(sv.info().n_isect) == 0
            if step == 0:
                want += batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
            batched.backward_peer_raw(sv, p, vm, Ks, bg, w_rgb, w_a, ptrs, rank, world)
            sv.release()
        for rank in range(world):
            owned = [f for f in range(F) if f % world == rank]
            out = torch.empty(fpr, N, P, device=dev)
            batched.peer_sum_raw(stages[rank], out)
            for k, f in enumerate(owned):
                scale = want[f].abs().amax(0).clamp_min(1e-20)
                assert torch.isfinite(out[k]).all()
                assert float(((out[k] - want[f]).abs().amax(0) / scale).max()) < 1e-4
