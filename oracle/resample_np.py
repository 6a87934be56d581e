"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's resampling step (radiomics/imageoperations.py:448-612):
the grid arithmetic of resampleImage (:493-566) in numpy, and SimpleITK's ResampleImageFilter(sitkBSpline) /
(sitkNearestNeighbor) restated with SciPy's cubic B-spline machinery (spline_filter + map_coordinates, mirror boundaries --
the same Unser recursive filter ITK's BSplineDecompositionImageFilter implements; SimpleITK is not installed here).

PINNED on the reference's own baseline: with the value cast done by clamping + TRUNCATION (ITK's ResampleImageFilter), the
first-order features of the `breast1_resampling` column of data/baseline/baseline_firstorder.csv (23 ROI voxels after
resampling to 2 mm) are reproduced exactly -- Mean 126.17391304347827, Energy 375154 (tests/test_resample_cpu.py); with
rounding instead they are not.  Only tests/ may import this.
"""
from __future__ import annotations

import numpy as np


def grid(mask, spacing_xyz, new_spacing_xyz, padDistance=5, label=1, offset_xyz=(0, 0, 0), full_size_xyz=None):
    """(newSize xyz, start xyz, step xyz, newSpacing): output voxel k samples the input at continuous index start + k*step.
    `offset_xyz` / `full_size_xyz`: the arrays are a crop of a larger image starting at that index -- the output grid is
    anchored at the FULL image's index 0 (imageoperations.py:520-548), so a committed crop must say where it sat."""
    sp = np.array(spacing_xyz, float)
    new = np.array(new_spacing_xyz, float)
    new = np.where(new == 0, sp, new)
    idx = np.array(np.where(np.asarray(mask) == label))
    lo, hi = idx.min(1)[::-1], idx.max(1)[::-1]
    off = np.array(offset_xyz, float)
    bb = np.concatenate([lo + off, hi - lo + 1]).astype(float)
    nd = len(sp)
    size = np.array(np.asarray(mask).shape[::-1] if full_size_xyz is None else full_size_xyz, float)
    new = np.where(bb[nd:] != 1, new, sp)
    ratio = sp / new
    L = np.floor((bb[:nd] - 0.5) * ratio - padDistance)
    U = np.ceil((bb[:nd] + bb[nd:] - 0.5) * ratio + padDistance)
    maxU = np.ceil(size * ratio) - 1
    L = np.where(L < 0, 0, L)
    U = np.where(U > maxU, maxU, U)
    return (U - L + 1).astype(int), 0.5 * (new - sp) / sp + L / ratio - off, new / sp, new


def resample(image, mask, spacing_xyz, new_spacing_xyz, padDistance=5, label=1, order=3, offset_xyz=(0, 0, 0), full_size_xyz=None):
    import scipy.ndimage as ndi
    image, mask = np.asarray(image), np.asarray(mask)
    newSize, start, step, new = grid(mask, spacing_xyz, new_spacing_xyz, padDistance, label, offset_xyz, full_size_xyz)
    g = [start[d] + step[d] * np.arange(newSize[d]) for d in range(3)]          # x, y, z
    zz, yy, xx = np.meshgrid(g[2], g[1], g[0], indexing="ij")
    if order == 3:
        coef = ndi.spline_filter(image.astype(np.float64), order=3, mode="mirror")
        val = ndi.map_coordinates(coef, [zz, yy, xx], order=3, mode="mirror", prefilter=False)
    else:
        val = ndi.map_coordinates(image.astype(np.float64), [zz, yy, xx], order=order, mode="nearest")
    inside = np.ones(val.shape, bool)
    for c, n in ((zz, image.shape[0]), (yy, image.shape[1]), (xx, image.shape[2])):
        inside &= (c >= -0.5) & (c < n - 0.5)
    val = np.where(inside, val, 0.0)
    if np.issubdtype(image.dtype, np.integer):
        info = np.iinfo(image.dtype)
        val = np.trunc(np.clip(val, info.min, info.max))
    out = val.astype(image.dtype)
    nn = [np.floor(c + 0.5).astype(int) for c in g]
    ins = [(c >= -0.5) & (c < n - 0.5) for c, n in zip(g, mask.shape[::-1])]
    mz, my, mx = np.meshgrid(nn[2], nn[1], nn[0], indexing="ij")
    iz, iy, ix = np.meshgrid(ins[2], ins[1], ins[0], indexing="ij")
    m = np.where(iz & iy & ix, mask[np.clip(mz, 0, mask.shape[0] - 1), np.clip(my, 0, mask.shape[1] - 1), np.clip(mx, 0, mask.shape[2] - 1)], 0)
    return out, m.astype(mask.dtype), new
