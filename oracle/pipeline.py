"""TEST INFRASTRUCTURE ONLY -- end-to-end CPU oracle of the hot path: gray-level
discretisation -> texture matrix -> features, segment-based or voxel-based, mirroring the
reference call stack.  `cm` selects the matrix backend: the C port
(oracle/cmatrices_oracle.py, default) or the compiled reference (oracle/build_ref.load()).
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cmatrices_oracle as _port  # noqa: E402
import features_np as F  # noqa: E402

CLASS_NAMES = ("glcm", "glrlm", "glszm", "gldm", "ngtdm")


def bin_edges(values, binWidth=25, binCount=None):
    """reference radiomics/imageoperations.py:67-153 (getBinEdges)."""
    values = np.asarray(values)
    if binCount is not None:
        e = np.histogram(values, binCount)[1]
        e[-1] += 1
        return e
    lo, hi = values.min(), values.max()
    low = lo - (lo % binWidth)
    e = np.arange(low, hi + 2 * binWidth, binWidth)
    if len(e) == 1:
        e = np.array([e[0] - 0.5, e[0] + 0.5])
    return e


def bin_image(image, mask, binWidth=25, binCount=None, **_):
    """reference radiomics/imageoperations.py:156-174 (binImage) + base.py:119-125."""
    out = np.zeros(image.shape, dtype=np.int64)
    e = bin_edges(image[mask], binWidth, binCount)
    out[mask] = np.digitize(image[mask], e)
    levels = np.unique(out[mask])
    return out, e, levels, int(levels.max())


def extract(cname, image, mask, voxelBased=False, spacing_zyx=(1.0, 1.0, 1.0), cm=None, voxels=None, **kw):
    """Returns {feature: float} (segment) or {feature: array[Nvox]} for the voxel list
    `voxels` (default all masked voxels, C order)."""
    cm = cm or _port
    maskArr = np.asarray(mask, bool)
    binmask = maskArr if kw.get("maskedKernel", True) or not voxelBased else np.ones_like(maskArr)
    img, _, levels, Ng = bin_image(np.asarray(image), binmask, kw.get("binWidth", 25), kw.get("binCount"))
    f2, f2d = kw.get("force2D", False), kw.get("force2Ddimension", 0)
    dist = np.array(kw.get("distances", [1]))
    extra = []
    if voxelBased:
        if voxels is None:
            voxels = np.array(np.where(maskArr)).astype(np.int32)
        extra = [kw.get("kernelRadius", 1), voxels]
    if cname == "glcm":
        P, ang = cm.calculate_glcm(img, binmask, dist, Ng, f2, f2d, *extra)
        w = F.angle_weights(ang, spacing_zyx, kw.get("weightingNorm"), "glcm")
        p = F.glcm_matrix(P, levels, kw.get("symmetricalGLCM", True), w)
        res = F.glcm_features(p, levels, Ng)
    elif cname == "glrlm":
        P, ang = cm.calculate_glrlm(img, binmask, Ng, int(max(img.shape)), f2, f2d, *extra)
        w = F.angle_weights(ang, spacing_zyx, kw.get("weightingNorm"), "glrlm")
        res = F.glrlm_features(P, levels, w)
    elif cname == "glszm":
        P = cm.calculate_glszm(img, binmask, Ng, int(binmask.sum()), f2, f2d, *extra)
        res = F.glszm_features(P, levels)
    elif cname == "gldm":
        P = cm.calculate_gldm(img, binmask, dist, Ng, int(kw.get("gldm_a", 0)), f2, f2d, *extra)
        res = F.gldm_features(P, levels)
    elif cname == "ngtdm":
        P = cm.calculate_ngtdm(img, binmask, dist, Ng, f2, f2d, *extra)
        res = F.ngtdm_features(P)
    else:
        raise KeyError(cname)
    if not voxelBased:
        res = {k: float(np.squeeze(v)) for k, v in res.items()}
    return res
