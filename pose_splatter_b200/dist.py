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
    """Fused cross-GPU gradient sum (K8'): one d_params buffer [F,N,P] per rank, allocated as symmetric memory and
    mapped into every process of the node (NVLink peer access).  With it the projection-backward kernel adds each
    finished row straight into the buffer of the rank that owns the frame (ps_backward_peer) -- no all-reduce, no
    staging copy.  frame f is owned by rank `f % world`.

        bufs = PeerGradBuffers((F, N, P), device)
        bufs.begin()                      # zero own buffer, barrier
        batched.backward_peer_raw(saved, params, viewmats, Ks, bg, d_rgb, d_alpha, bufs.rank_ptrs, bufs.frame_owner)
        bufs.end()                        # barrier: bufs.buf[f] is complete for every owned frame f
    """

    def __init__(self, shape, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buf = symm_mem.empty(*shape, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.rank_ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.frame_owner = (torch.arange(shape[0], device=device) % self.world).to(torch.int32)
        self.owned = [f for f in range(shape[0]) if f % self.world == self.rank]

    def begin(self):
        self.buf.zero_()
        self.handle.barrier()

    def end(self):
        self.handle.barrier()


def max_over_ranks(value: float, device) -> float:
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
