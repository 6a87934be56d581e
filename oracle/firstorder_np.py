"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's first-order class
(reference radiomics/firstorder.py:40-474), segment-based and voxel-based.

Voxel-based semantics follow firstorder.py:40-122: the kernel of a centre voxel is its
(2r+1)^3 window (offsets limited to |d| < ROI-bbox size per dimension, force2D dimension
collapsed), values outside the mask / volume are NaN and ignored by the nan-aware statistics.

ONE deliberate deviation: the reference indexes its *unpadded* discretised array with the
*padded* kernel coordinates (firstorder.py:109, `self.discretizedImageArray[kernelCoords]`
after `voxelCoordinates + kernelRadius`), so its voxel-mode Entropy / Uniformity histogram a
window shifted by +kernelRadius (or raise IndexError near the far border).  Here p_i is taken
from the same window as every other feature -- the segment-mode definition applied per voxel.
Pinned by: data/baseline/baseline_firstorder.csv (segment mode, all 18 features) and reference
voxel-mode runs for the 16 features that do not use p_i (tests/golden/voxel_firstorder.npz).
"""
from __future__ import annotations

import numpy as np

import pipeline as PL

EPS = np.spacing(1)
NAMES = ["10Percentile", "90Percentile", "Energy", "Entropy", "InterquartileRange", "Kurtosis", "Maximum",
         "MeanAbsoluteDeviation", "Mean", "Median", "Minimum", "Range", "RobustMeanAbsoluteDeviation",
         "RootMeanSquared", "Skewness", "TotalEnergy", "Uniformity", "Variance"]


def _features(T, p_i, shift, voxel_volume):
    """T: [V, K] target values (NaN = not in kernel); p_i: [V, L] normalised level histogram."""
    with np.errstate(invalid="ignore", divide="ignore"):
        f = {}
        sh = T + shift
        f["Energy"] = np.nansum(sh ** 2, 1)
        f["TotalEnergy"] = f["Energy"] * voxel_volume
        f["Entropy"] = -np.sum(p_i * np.log2(p_i + EPS), 1)
        f["Minimum"] = np.nanmin(T, 1)
        f["Maximum"] = np.nanmax(T, 1)
        p10, p25, p75, p90 = (np.nanpercentile(T, q, axis=1) for q in (10, 25, 75, 90))
        f["10Percentile"], f["90Percentile"] = p10, p90
        f["Mean"] = np.nanmean(T, 1)
        f["Median"] = np.nanmedian(T, 1)
        f["InterquartileRange"] = p75 - p25
        f["Range"] = f["Maximum"] - f["Minimum"]
        mu = np.nanmean(T, 1, keepdims=True)
        f["MeanAbsoluteDeviation"] = np.nanmean(np.abs(T - mu), 1)
        R = T.copy()
        R[(T < p10[:, None]) | (T > p90[:, None])] = np.nan
        f["RobustMeanAbsoluteDeviation"] = np.nanmean(np.abs(R - np.nanmean(R, 1, keepdims=True)), 1)
        n = np.sum(~np.isnan(T), 1).astype(float)
        f["RootMeanSquared"] = np.sqrt(np.nansum(sh ** 2, 1) / n)
        m2 = np.nanmean((T - mu) ** 2, 1)
        m3 = np.nanmean((T - mu) ** 3, 1)
        m4 = np.nanmean((T - mu) ** 4, 1)
        m2s = np.where(m2 == 0, 1.0, m2)
        f["Skewness"] = m3 / m2s ** 1.5
        f["Kurtosis"] = m4 / m2s ** 2.0
        f["Variance"] = np.nanstd(T, 1) ** 2
        f["Uniformity"] = np.nansum(p_i ** 2, 1)
    return f


def kernel_offsets(shape, mask, r, masked=True, force2D=False, force2Ddimension=0):
    import cmatrices_oracle as O
    if masked:
        idx = np.array(np.where(mask))
        size = idx.max(1) - idx.min(1) + 1
    else:
        size = np.array(shape)
    bb = np.minimum(size, 2 * r + 1)
    off = O.generate_angles(bb, np.arange(1, r + 1), True, force2D, force2Ddimension)
    return np.concatenate([off, np.zeros((1, off.shape[1]), off.dtype)], 0)


def extract(image, mask, voxelBased=False, spacing_xyz=(1.0, 1.0, 1.0), **kw):
    image = np.asarray(image)
    label = np.asarray(mask, bool)
    shift = kw.get("voxelArrayShift", 0)
    vv = float(np.multiply.reduce(spacing_xyz))
    binmask = label if (kw.get("maskedKernel", True) or not voxelBased) else np.ones_like(label)
    lev, _, levels, _ = PL.bin_image(image, binmask, kw.get("binWidth", 25), kw.get("binCount"))
    if not voxelBased:
        T = image[label].astype(float)[None, :]
        _, cnt = np.unique(lev[label], return_counts=True)
        p = (cnt / cnt.sum())[None, :]
        return {k: float(np.squeeze(v)) for k, v in _features(T, p, shift, vv).items()}
    r = kw.get("kernelRadius", 1)
    off = kernel_offsets(image.shape, binmask, r, kw.get("maskedKernel", True), kw.get("force2D", False), kw.get("force2Ddimension", 0))
    img = image.astype(float)
    img[~binmask] = np.nan
    img = np.pad(img, r, constant_values=np.nan)
    levp = np.pad(np.where(binmask, lev, 0), r, constant_values=0)
    vox = np.array(np.where(label)) + r
    coords = tuple(off.T[:, None, :] + vox[:, :, None])
    T = img[coords]
    L = levp[coords]
    p = np.stack([(L == g).sum(1) for g in levels], 1).astype(float)
    s = p.sum(1, keepdims=True)
    p = p / np.where(s == 0, 1, s)
    return _features(T, p, shift, vv)
