"""world_size-2 gloo tests (CPU) of the slab decomposition + halo exchange host logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyradiomics_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Z, r, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    vol = rng.integers(1, 33, (Z, 6, 5)).astype(np.uint8)
    z0, z1 = D.slab_range(Z, rank, world)
    slab = D.SlabHalo(torch.from_numpy(vol[z0:z1].copy()), r, rank, world)
    slab.exchange()
    # expected: global planes z0-r .. z1+r-1, zeros outside the volume
    exp = np.zeros((z1 - z0 + 2 * r, 6, 5), np.uint8)
    for k in range(z0 - r, z1 + r):
        if 0 <= k < Z:
            exp[k - (z0 - r)] = vol[k]
    ok = np.array_equal(slab.buf.numpy(), exp)
    alive = np.zeros(6, np.uint32)
    alive[0] = 1 << rank
    alive[1] = 0x10 if rank == 1 else 0
    red = D.allreduce_alive(alive, torch.device("cpu"))
    ok = ok and int(red[0]) == (1 << world) - 1 and int(red[1]) == 0x10
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,Z,r", [(2, 9, 1), (2, 8, 2), (3, 10, 1)])
def test_halo_exchange_and_alive_allreduce(world, Z, r):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(k, world, port, Z, r, q)) for k in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_slab_ranges_cover_volume():
    for Z in (1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            rs = [D.slab_range(Z, k, world) for k in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == Z
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))


def _worker_filters(rank, world, port, Z, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    Y, X = 7, 5
    vol = rng.normal(size=(Z, Y, X))
    z0, z1 = D.slab_range(Z, rank, world)
    # periodic asymmetric halo (coif1: 2 planes below, 3 above), ring closed between rank 0 and rank world-1
    slab = D.SlabHalo(torch.from_numpy(vol[z0:z1].copy()), 2, rank, world, hi=3, periodic=True)
    slab.exchange()
    exp = np.stack([vol[k % Z] for k in range(z0 - 2, z1 + 3)])
    ok = np.array_equal(slab.buf.numpy(), exp)
    # z-slab -> y-slab -> z-slab
    ys = D.zslab_to_yslab(torch.from_numpy(vol[z0:z1].copy()), Z, rank, world)
    y0, y1 = D.slab_range(Y, rank, world)
    ok = ok and np.array_equal(ys.numpy(), vol[:, y0:y1, :])
    back = D.yslab_to_zslab(ys, Y, rank, world)
    ok = ok and np.array_equal(back.numpy(), vol[z0:z1])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,Z", [(2, 8), (2, 11), (3, 13)])
def test_periodic_halo_ring_and_slab_transposition(world, Z):
    """multi-GPU pre-filters: the wavelet's ring-closed 2+3-plane halo and the LoG z pass's transposition"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_filters, args=(k, world, port, Z, q)) for k in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_periodic_halo_single_rank():
    vol = torch.arange(6 * 2 * 2, dtype=torch.float64).reshape(6, 2, 2)
    s = D.SlabHalo(vol, 2, 0, 1, hi=3, periodic=True)
    s.exchange()
    assert torch.equal(s.buf, torch.stack([vol[k % 6] for k in range(-2, 9)]))
