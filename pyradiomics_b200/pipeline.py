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

import numpy as np
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
        dec = IO.swt_level1_device(x.to(torch.float64), (2, 1, 0), lo, hi)
        for key, t in dec.items():
            if key != "aaa":
                yield "wavelet-" + key.replace("a", "L").replace("d", "H"), t
        yield "wavelet-LLL", dec["aaa"]
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
