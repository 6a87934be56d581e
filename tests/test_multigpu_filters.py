"""2-GPU test (NCCL) of the multi-GPU pre-filters: wavelet with the ring-closed periodic halo and LoG with
the z-slab <-> y-slab transposition must reproduce the single-GPU derived images BIT FOR BIT, and the whole config-4
chain on slabs (global bin edges, halo of the packed levels) the single-GPU feature maps.  Skipped with fewer than two
GPUs (the driver's 1-GPU run); run it with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_filters.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, q):
    import torch.distributed as dist
    from pyradiomics_b200 import distributed as D, pipeline as PL
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import scipy.ndimage as ndi
        rng = np.random.default_rng(3)
        vol = (ndi.gaussian_filter(rng.normal(size=shape), 1.5) * 400 + 300).astype(np.float32)
        Z = shape[0]
        z0, z1 = D.slab_range(Z, rank, world)
        full = torch.from_numpy(vol).to(dev)
        own = full[z0:z1].contiguous()
        ref = dict(PL.derived_images(full, sigmas=(1.0, 2.5)))
        got = dict(PL.derived_images_slab(own, Z, rank, world, sigmas=(1.0, 2.5)))
        bad = [n for n in ref if not torch.equal(got[n].contiguous(), ref[n][z0:z1].contiguous())]
        ok = set(got) == set(ref) and not bad
        # the whole chain: filters -> global binning -> fused kernels on the slab
        mask = torch.ones(shape, dtype=torch.uint8, device=dev)
        maps_ref, maps_got = {}, {}
        PL.voxel_suite_with_filters(full, mask, classes=("glcm", "gldm"), sigmas=(1.0,), binWidth=25,
                                    consume=lambda n, c, t: maps_ref.__setitem__((n, c), t[:, z0:z1].clone()))
        PL.voxel_suite_with_filters_slab(own, mask[z0:z1].contiguous(), Z, rank, world, classes=("glcm", "gldm"), sigmas=(1.0,),
                                         binWidth=25, consume=lambda n, c, t: maps_got.__setitem__((n, c), t.clone()))
        bad2 = [k for k in maps_ref if not torch.equal(maps_got[k].view(torch.int64), maps_ref[k].view(torch.int64))]
        q.put((rank, bool(ok and not bad2 and len(maps_got) == len(maps_ref)), bad[:3], [str(b) for b in bad2[:3]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("shape", [(24, 20, 28), (21, 19, 23)])
def test_two_gpu_prefilters_equal_single_gpu(shape):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(k, 2, port, shape, q)) for k in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
