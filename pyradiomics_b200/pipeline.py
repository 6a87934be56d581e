"""Device-resident end-to-end pipelines over the hot path (BASELINE.json configs 4 and 5):

* ``voxel_suite_with_filters`` -- Original + wavelet (8 sub-bands) + LoG (one per sigma) derived
  images -> per-image gray-level discretisation -> the five fused voxel-based texture kernels,
  everything on one GPU without host round trips (what ``RadiomicsFeatureExtractor.execute(...,
  voxelBased=True)`` does image type by image type, reference radiomics/featureextractor.py:371-392).
* ``segment_batch`` -- segment-based matrices + features for a list of independent cases, sharded
  round-robin over the ranks of the process group with no collective (the reference's own
  parallel model: one case per worker, radiomics/scripts/__init__.py:393-404).
"""
from __future__ import annotations

import torch

from . import _lib, featureclasses as FC, imageoperations as IO, voxel
from ._lib import CLASSES


def derived_images(x: torch.Tensor, spacing_zyx=(1.0, 1.0, 1.0), wavelet="coif1", sigmas=(1.0, 2.0, 3.0),
                   original=True):
    """yields (name, CUDA tensor) like the reference's imageType generators"""
    if original:
        yield "original", x
    if wavelet:
        lo, hi = IO.wavelet_filters(wavelet)
        xp, crop = IO._wrap_pad_even(x.to(torch.float64), (2, 1, 0))
        dec = IO.swt_level1_device(xp, (2, 1, 0), lo, hi)
        for key, t in dec.items():
            if key != "aaa":
                yield "wavelet-" + key.replace("a", "L").replace("d", "H"), t[crop]
        yield "wavelet-LLL", dec["aaa"][crop]
    for s in sigmas or ():
        yield f"log-sigma-{str(float(s)).replace('.', '-')}-mm-3D", IO.log_filter_device(x, float(s), spacing_zyx)


def voxel_suite_with_filters(image: torch.Tensor, mask: torch.Tensor, classes=CLASSES, spacing_zyx=(1.0, 1.0, 1.0),
                             wavelet="coif1", sigmas=(1.0, 2.0, 3.0), consume=None, **kw):
    """image: CUDA tensor (Z,Y,X) of raw intensities, mask: CUDA uint8/bool.  For every derived
    image: bin (binWidth/binCount in kw) -> pack -> fused kernels.  `consume(name, cls, maps)` is
    called with each float64 [F,Z,Y,X] result (maps are reused buffers unless consume keeps them);
    returns the list of (image name, Ng, number of levels)."""
    msk = (mask != 0).to(torch.uint8).contiguous()
    outs = {}
    info = []
    for name, img in derived_images(image, spacing_zyx, wavelet, sigmas):
        lev32, _ = IO.bin_image_device(img.contiguous(), msk, **kw)
        Ng = int(lev32.max().item())
        lev, presence = voxel.pack_levels(lev32, msk, Ng)
        nlev = int((presence > 0).sum().item())
        s = _lib.make_settings(Ng, nlev, spacing_zyx=spacing_zyx, **kw)
        for c in classes:
            nf = _lib.lib().rb_num_features(_lib.CLASS_ID[c])
            if c not in outs:
                outs[c] = torch.empty((nf,) + tuple(lev.shape), dtype=torch.float64, device=lev.device)
            maps = voxel.voxel_features(c, lev, s, out=outs[c], out_z0=0)
            if consume is not None:
                consume(name, c, maps)
        info.append((name, Ng, nlev))
    return info


def segment_batch(cases, classes=tuple(FC.FEATURE_CLASSES), rank=0, world=1, **kw):
    """cases: sequence of (image ndarray, mask ndarray).  Rank `rank` of `world` processes the cases
    k with k % world == rank on its current CUDA device; returns {case index: {class: {feature: value}}}."""
    res = {}
    for k, (img, msk) in enumerate(cases):
        if k % world != rank:
            continue
        res[k] = {c: {f: float(v) for f, v in FC.FEATURE_CLASSES[c](img, msk, **kw).execute().items()} for c in classes}
    return res


# ---------------------------------------------------------------------------------------------- multi-GPU pre-filters
def derived_images_slab(own: torch.Tensor, Z: int, rank: int, world: int, spacing_zyx=(1.0, 1.0, 1.0), wavelet="coif1",
                        sigmas=(1.0, 2.0, 3.0), original=True):
    """`derived_images` for a volume that is sharded into z-slabs over the ranks of the default process group
    (SURVEY.md section 8e): yields (name, this rank's slab of the derived image).
      * wavelet: the transform is periodic, so the slab gets (F-1-F/2) planes from the rank below and F/2 from the rank
        above, ring-closed between rank 0 and the last rank (distributed.SlabHalo(periodic=True)); an odd global Z is
        wrap-padded by handing the last rank a copy of rank 0's first plane, like the reference pads before transforming;
      * LoG: the recursive Gaussian along z is a sequential scan over whole lines -- z-slabs are transposed to y-slabs
        for that pass and back (distributed.zslab_to_yslab); the x and y passes are local.  Bit-identical to one GPU."""
    import torch.distributed as dist
    from . import distributed as D
    if original:
        yield "original", own
    nz, Y, X = own.shape
    if wavelet:
        lo, hi = IO.wavelet_filters(wavelet)
        F = int(lo.size)
        low, up = F - 1 - F // 2, F // 2
        x64 = own.to(torch.float64)
        last = world - 1
        if Z % 2 and world > 1:                     # global wrap-pad along z: the last rank appends global plane 0
            if rank == 0:
                dist.send(x64[0:1].contiguous(), last)
            if rank == last:
                first = torch.empty_like(x64[0:1])
                dist.recv(first, 0)
                x64 = torch.cat([x64, first], 0)
        elif Z % 2:
            x64 = torch.cat([x64, x64[0:1]], 0)
        xp, crop = IO._wrap_pad_even(x64, (2, 1))           # y, x: local
        slab = D.SlabHalo(xp, low, rank, world, hi=up, periodic=True)
        slab.exchange()
        dec = IO.swt_level1_device(slab.buf, (2, 1, 0), lo, hi, z_range=(low, low + xp.shape[0]))
        crop = (slice(0, nz),) + tuple(crop[1:])
        for key, t in dec.items():
            if key != "aaa":
                yield "wavelet-" + key.replace("a", "L").replace("d", "H"), t[crop]
        yield "wavelet-LLL", dec["aaa"][crop]

    def z_pass(t, sigma_vox, order, scale):
        ys = D.zslab_to_yslab(t, Z, rank, world)
        return D.yslab_to_zslab(IO._rg_pass(ys.contiguous(), 0, sigma_vox, order, scale=scale), Y, rank, world)

    for s in sigmas or ():
        yield (f"log-sigma-{str(float(s)).replace('.', '-')}-mm-3D",
               IO.log_filter_device(own, float(s), spacing_zyx, z_pass=z_pass if world > 1 else None))


def voxel_suite_with_filters_slab(own: torch.Tensor, own_mask: torch.Tensor, Z: int, rank: int, world: int, classes=CLASSES,
                                  spacing_zyx=(1.0, 1.0, 1.0), wavelet="coif1", sigmas=(1.0, 2.0, 3.0), consume=None, **kw):
    """BASELINE.json config 4 on z-slabs: every derived image is binned with the WHOLE ROI's edges (all-reduced min / max
    and gray-level presence), the packed levels exchange one halo plane per face, and the fused texture kernels run on
    the slab.  `consume(name, cls, maps[F, nz, Y, X])` sees each result; returns [(image name, Ng, number of levels)]."""
    import torch.distributed as dist
    from . import distributed as D
    msk = (own_mask != 0).to(torch.uint8).contiguous()
    dev = own.device
    r = int(kw.get("kernelRadius", 1))

    def reduce_minmax(mn, mx):
        if world == 1:
            return mn, mx
        t = torch.tensor([-mn, mx], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return -float(t[0].item()), float(t[1].item())

    info, outs = [], {}
    nz = own.shape[0]
    for name, img in derived_images_slab(own, Z, rank, world, spacing_zyx, wavelet, sigmas):
        lev32, _ = IO.bin_image_device(img.contiguous(), msk, minmax_reduce=reduce_minmax, **kw)
        ngt = lev32.max().to(torch.int64).reshape(1)
        if world > 1:
            dist.all_reduce(ngt, op=dist.ReduceOp.MAX)
        Ng = int(ngt.item())
        lev, presence = voxel.pack_levels(lev32, msk, Ng)
        pres = presence.to(torch.int64)
        if world > 1:
            dist.all_reduce(pres, op=dist.ReduceOp.SUM)
        nlev = int((pres > 0).sum().item())
        s = _lib.make_settings(Ng, nlev, spacing_zyx=spacing_zyx, **kw)
        slab = D.SlabHalo(lev, r, rank, world)
        slab.exchange()
        alive = D.allreduce_alive(voxel.glcm_alive_angles(slab.buf, s), dev) if "glcm" in classes else None
        for c in classes:
            nf = _lib.lib().rb_num_features(_lib.CLASS_ID[c])
            if c not in outs:
                outs[c] = torch.empty((nf, nz) + tuple(lev.shape[1:]), dtype=torch.float64, device=dev)
            maps = voxel.voxel_features(c, slab.buf, s, z0=r, z1=r + nz, out=outs[c], out_z0=r, alive=alive if c == "glcm" else None)
            if consume is not None:
                consume(name, c, maps)
        info.append((name, Ng, nlev))
    return info
