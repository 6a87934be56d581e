"""Shared helpers of the parity tests."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

import pipeline as PL  # oracle/pipeline.py (tests may use the oracle; the product may not)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# relative tolerance north_star states for derived float features
RTOL = 1e-5


def voxel_goldens():
    out = []
    for fn in sorted(glob.glob(os.path.join(GOLDEN, "voxel_*.npz"))):
        z = np.load(fn)
        if "settings" not in z.files:        # e.g. voxel_firstorder.npz has its own tests
            continue
        out.append((os.path.basename(fn)[6:-4], z, json.loads(str(z["settings"]))))
    return out


def binned(z, kw):
    lev, _, levels, Ng = PL.bin_image(z["image"], z["mask"], kw.get("binWidth", 25), kw.get("binCount"))
    return lev, levels, Ng


def ref_map(z, cname, fname):
    """Golden map of the reference; MCC uses the voxelBatch=1 run (see make_golden.py)."""
    if cname == "glcm" and fname == "MCC":
        return z["glcm_MCC_voxelBatch1"]
    return z[f"{cname}_{fname}"]


def assert_maps_close(got, ref, what, rtol=RTOL, atol=1e-9):
    """Feature-map comparison with the tolerance of BASELINE.json (1e-5 relative); NaNs must
    coincide.  `atol` absorbs values that are pure rounding noise in the reference itself
    (e.g. entropies of a one-entry matrix: -log2(1+eps) ~ 3e-16, Imc2 of independent margins)."""
    got = np.asarray(got, float)
    ref = np.asarray(ref, float)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    ok = np.isclose(got, ref, rtol=rtol, atol=atol, equal_nan=True)
    if not ok.all():
        bad = np.argwhere(~ok)
        b = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {ok.size} voxels differ; first {b}: got {got[b]!r} ref {ref[b]!r}")
