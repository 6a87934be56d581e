"""Multi-GPU plumbing for voxel-based extraction: 1-D z-slab decomposition with halo planes
exchanged between slab neighbours.  One process per GPU; torch.distributed
(NCCL on GPUs, gloo in the CPU tests) carries the r halo planes per face -- the only data-path
communication -- plus one tiny all-reduce for the set of GLCM angles that are non-empty anywhere
in the ROI.  Feature maps stay sharded by slab.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_range(Z: int, rank: int, world: int):
    """contiguous plane range [z0, z1) owned by `rank` (first Z % world ranks get one more)."""
    base, extra = divmod(Z, world)
    z0 = rank * base + min(rank, extra)
    return z0, z0 + base + (1 if rank < extra else 0)


class SlabHalo:
    """A rank's slab of the level volume with `r` halo planes on both sides:
    buf[0:r] = planes owned by rank-1, buf[r:r+nz] = own planes, buf[r+nz:] = planes of rank+1
    (zeros at the volume boundary = 'unmasked', exactly how the kernels treat out-of-volume)."""

    def __init__(self, own: torch.Tensor, r: int, rank: int, world: int):
        self.r, self.rank, self.world = r, rank, world
        nz = own.shape[0]
        self.nz = nz
        self.buf = torch.zeros((nz + 2 * r,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
        self.buf[r:r + nz] = own

    def exchange(self):
        """send own boundary planes to both neighbours, receive theirs into the halo."""
        r, nz, buf = self.r, self.nz, self.buf
        if self.world == 1:
            return
        ops = []
        lo_send = buf[r:2 * r].contiguous()
        hi_send = buf[nz:nz + r].contiguous()
        lo_recv = torch.empty_like(lo_send)
        hi_recv = torch.empty_like(hi_send)
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, lo_send, self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, lo_recv, self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, hi_send, self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, hi_recv, self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.rank > 0:
            buf[0:r] = lo_recv
        if self.rank < self.world - 1:
            buf[r + nz:] = hi_recv


def allreduce_alive(alive_words: np.ndarray, device) -> np.ndarray:
    """bitwise OR of the per-rank 'angle is non-empty somewhere' masks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return alive_words
    bits = np.unpackbits(alive_words.view(np.uint8), bitorder="little").astype(np.int32)
    t = torch.from_numpy(bits).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = np.packbits(t.cpu().numpy().astype(np.uint8), bitorder="little").view(np.uint32)
    return out.copy()
