"""Multi-GPU partitioning of (frame, camera) views: one process per GPU over torch.distributed.

Views are independent units: forward needs no exchange.  The only collective is
the sum of per-Gaussian gradients d_params[F,N,P] when the cameras of one frame land on
different ranks (3D; <= 896 KB per frame) -- one all-reduce per step over NCCL / NVLink.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_views(n_frames: int, n_cams: int, rank: int, world: int, policy: str = "frame") -> List[int]:
    """Global view ids v = frame * n_cams + cam owned by `rank`.

    policy "frame": whole frames per rank (frame % world == rank) -> no gradient exchange at all.
    policy "view" : views round-robin (v % world == rank) -> one frame's cameras are split across
                    ranks and their d_params must be summed (BASELINE.json configs[4]).
    """
    if policy == "frame":
        return [f * n_cams + c for f in range(n_frames) if f % world == rank for c in range(n_cams)]
    if policy == "view":
        return [v for v in range(n_frames * n_cams) if v % world == rank]
    raise ValueError(f"unknown shard policy {policy!r}")


def frames_shared_across_ranks(n_frames: int, n_cams: int, world: int, policy: str) -> bool:
    if world == 1 or policy == "frame":
        return False
    owners = [{(f * n_cams + c) % world for c in range(n_cams)} for f in range(n_frames)]
    return any(len(o) > 1 for o in owners)


def reduce_frame_grads(d_params: torch.Tensor, group=None) -> torch.Tensor:
    """Sum d_params [F,N,P] over the ranks that rendered views of the same frames (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d_params, op=dist.ReduceOp.SUM, group=group)
    return d_params


class PeerGradBuffers:
    """Fused cross-GPU gradient exchange (K8').  Frame f is owned by rank f % world.  Every rank holds a staging
    buffer [world, ceil(F/world), N, P] in symmetric memory, mapped into every process of the node (NVLink peer
    access).  The projection-backward kernel pushes each finished block of rows straight into slot
    [my rank][f // world] of the owner's buffer (ps_backward_peer) -- no all-reduce, no atomics; after a barrier the
    owner adds its `world` slots.

        bufs = PeerGradBuffers((F, N, P), device)
        bufs.begin()                                   # barrier: last step's staging has been consumed
        batched.backward_peer_raw(saved, params, viewmats, Ks, bg, d_rgb, d_alpha, bufs.rank_ptrs, bufs.rank, bufs.world)
        grads = bufs.end()                             # barrier + local sum: [len(bufs.owned), N, P] for frames bufs.owned
    """

    def __init__(self, shape, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import batched
        self._batched = batched
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        F, N, P = shape
        self.frames_per_rank = (F + self.world - 1) // self.world
        self.stage = symm_mem.empty(self.world, self.frames_per_rank, N, P, dtype=torch.float32, device=device)
        self.stage.zero_()
        self.handle = symm_mem.rendezvous(self.stage, group)
        self.rank_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.owned = [f for f in range(F) if f % self.world == self.rank]
        self.out = torch.empty(self.frames_per_rank, N, P, dtype=torch.float32, device=device)
        self.handle.barrier()

    def begin(self):
        self.handle.barrier()

    def end(self) -> torch.Tensor:
        self.handle.barrier()
        self._batched.peer_sum_raw(self.stage, self.out)
        return self.out[:len(self.owned)]


def max_over_ranks(value: float, device) -> float:
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
