"""Multi-GPU plumbing for voxel-based extraction: 1-D z-slab decomposition with halo planes
exchanged between slab neighbours; ring-closed periodic halos for the wavelet pre-filter and the
z-slab <-> y-slab transposition the recursive-Gaussian z pass of the LoG pre-filter needs.  One process per GPU; torch.distributed
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
    """A rank's slab of a volume with halo planes on both sides:
    buf[0:lo] = planes owned by rank-1, buf[lo:lo+nz] = own planes, buf[lo+nz:] = `hi` planes of rank+1.
    Non-periodic (texture kernels, radius r on both sides): the halo at the volume boundary stays zero = 'unmasked',
    exactly how the kernels treat out-of-volume.  periodic=True closes the ring (rank 0 <-> rank world-1): the
    stationary wavelet transform is periodic along every axis (np.pad(..., "wrap") + periodized swtn, reference
    radiomics/imageoperations.py:914-919,935), and its 6-tap coif1 filter needs 2 planes below and 3 above."""

    def __init__(self, own: torch.Tensor, r: int, rank: int, world: int, hi: int | None = None, periodic: bool = False):
        self.lo, self.hi = int(r), int(r if hi is None else hi)
        self.r, self.rank, self.world, self.periodic = self.lo, rank, world, periodic
        nz = own.shape[0]
        self.nz = nz
        self.buf = torch.zeros((nz + self.lo + self.hi,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
        self.buf[self.lo:self.lo + nz] = own

    def exchange(self):
        """send own boundary planes to both neighbours, receive theirs into the halo."""
        lo, hi, nz, buf, rank, world = self.lo, self.hi, self.nz, self.buf, self.rank, self.world
        below = rank - 1 if rank > 0 else (world - 1 if self.periodic else None)
        above = rank + 1 if rank < world - 1 else (0 if self.periodic else None)
        if world == 1:
            if self.periodic:                      # one rank: the ring closes on itself
                if lo:
                    buf[0:lo] = buf[nz:nz + lo]
                if hi:
                    buf[lo + nz:] = buf[lo:lo + hi]
            return
        assert nz >= max(lo, hi), "slab thinner than the halo"
        ops = []
        up_send = buf[lo + nz - lo:lo + nz].contiguous() if lo else None      # my top `lo` planes are the lower halo of rank+1
        dn_send = buf[lo:lo + hi].contiguous() if hi else None                # my bottom `hi` planes are the upper halo of rank-1
        lo_recv = torch.empty_like(buf[0:lo]) if lo else None
        hi_recv = torch.empty_like(buf[lo + nz:]) if hi else None
        # order matters when both neighbours are the SAME rank (two ranks, periodic): messages between one pair match in
        # posting order -- "top planes up" pairs with the peer's "lower halo", "bottom planes down" with its "upper halo"
        if above is not None and lo:
            ops.append(dist.P2POp(dist.isend, up_send, above))
        if below is not None and lo:
            ops.append(dist.P2POp(dist.irecv, lo_recv, below))
        if below is not None and hi:
            ops.append(dist.P2POp(dist.isend, dn_send, below))
        if above is not None and hi:
            ops.append(dist.P2POp(dist.irecv, hi_recv, above))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if below is not None and lo:
            buf[0:lo] = lo_recv
        if above is not None and hi:
            buf[lo + nz:] = hi_recv


def _split(n: int, world: int):
    return [slab_range(n, k, world) for k in range(world)]


def zslab_to_yslab(own: torch.Tensor, Z: int, rank: int, world: int) -> torch.Tensor:
    """[nz_own, Y, X] z-slab of a (Z,Y,X) volume -> [Z, ny_own, X] y-slab of the same volume (grouped send/recv: every
    rank sends rank k the y-range of k out of its planes).  The recursive Gaussian along z is a sequential scan over the
    whole line (ITK's RecursiveGaussianImageFilter behind sitk.LaplacianRecursiveGaussian, reference
    radiomics/imageoperations.py:824-830), so that pass runs on y-slabs and the volume is transposed back afterwards."""
    nz, Y, X = own.shape
    if world == 1:
        return own
    zr, yr = _split(Z, world), _split(Y, world)
    y0, y1 = yr[rank]
    out = torch.empty((Z, y1 - y0, X), dtype=own.dtype, device=own.device)
    ops, recvs = [], []
    for k in range(world):
        blk = own[:, yr[k][0]:yr[k][1], :].contiguous()
        if k == rank:
            out[zr[rank][0]:zr[rank][1]] = blk
            continue
        rb = torch.empty((zr[k][1] - zr[k][0], y1 - y0, X), dtype=own.dtype, device=own.device)
        recvs.append((k, rb))
        ops.append(dist.P2POp(dist.isend, blk, k))
        ops.append(dist.P2POp(dist.irecv, rb, k))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for k, rb in recvs:
        out[zr[k][0]:zr[k][1]] = rb
    return out


def yslab_to_zslab(ys: torch.Tensor, Y: int, rank: int, world: int) -> torch.Tensor:
    """inverse of zslab_to_yslab: [Z, ny_own, X] -> [nz_own, Y, X]"""
    Z, ny, X = ys.shape
    if world == 1:
        return ys
    zr, yr = _split(Z, world), _split(Y, world)
    z0, z1 = zr[rank]
    out = torch.empty((z1 - z0, Y, X), dtype=ys.dtype, device=ys.device)
    ops, recvs = [], []
    for k in range(world):
        blk = ys[zr[k][0]:zr[k][1]].contiguous()
        if k == rank:
            out[:, yr[rank][0]:yr[rank][1], :] = blk
            continue
        rb = torch.empty((z1 - z0, yr[k][1] - yr[k][0], X), dtype=ys.dtype, device=ys.device)
        recvs.append((k, rb))
        ops.append(dist.P2POp(dist.isend, blk, k))
        ops.append(dist.P2POp(dist.irecv, rb, k))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for k, rb in recvs:
        out[:, yr[k][0]:yr[k][1], :] = rb
    return out


def allreduce_alive(alive_words: np.ndarray, device) -> np.ndarray:
    """bitwise OR of the per-rank 'angle is non-empty somewhere' masks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return alive_words
    bits = np.unpackbits(alive_words.view(np.uint8), bitorder="little").astype(np.int32)
    t = torch.from_numpy(bits).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = np.packbits(t.cpu().numpy().astype(np.uint8), bitorder="little").view(np.uint32)
    return out.copy()
