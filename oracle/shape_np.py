"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's shape path.

* ``coefficients(mask, spacing)`` follows radiomics/src/cshape.c:22-242 (calculate_coefficients +
  calculate_meshDiameter): marching cube over every 2x2x2 neighbourhood, surface area and signed
  origin volume per triangle, mesh vertices on the three cube edges that meet at corner (z+1, y+1, x),
  O(V^2) diameters.  The triangle table is the product's generated one (csrc/mc_table.inc, derived from
  geometry and pinned on the reference's single-cube outputs, see csrc/gen_mc_table.py); parity of this
  restatement is pinned by tests/golden/shape_cube_probes.npz, shape_random.npz and shape_expect.json
  (all produced by the compiled reference, tests/golden/make_golden.py --shape-only).
* ``features(mask, spacing)`` follows radiomics/shape.py:54-424.

Pure Python loops: only for small masks.  Only tests/ may import this.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "pyradiomics_b200", "csrc"))
import gen_mc_table  # noqa: E402

_MID2, _TRI = gen_mc_table.load_table()


def coefficients(mask, spacing):
    m = np.asarray(mask) != 0
    sp = np.asarray(spacing, dtype=np.float64)
    Z, Y, X = m.shape
    area = vol6 = 0.0
    verts = []
    for iz in range(Z - 1):
        for iy in range(Y - 1):
            for ix in range(X - 1):
                cfg = 0
                for c in range(8):
                    if m[iz + (c >> 2 & 1), iy + (c >> 1 & 1), ix + (c & 1)]:
                        cfg |= 1 << c
                own = cfg >> 6 & 1                                   # corner (1,1,0), cshape.c:94-112
                if (cfg >> 7 & 1) != own:
                    verts.append(((iz + 1.0) * sp[0], (iy + 1.0) * sp[1], (ix + 0.5) * sp[2]))
                if (cfg >> 4 & 1) != own:
                    verts.append(((iz + 1.0) * sp[0], (iy + 0.5) * sp[1], (ix + 0.0) * sp[2]))
                if (cfg >> 2 & 1) != own:
                    verts.append(((iz + 0.5) * sp[0], (iy + 1.0) * sp[1], (ix + 0.0) * sp[2]))
                row = _TRI[cfg]
                for k in range(0, 15, 3):
                    if row[k] < 0:
                        break
                    a, b, c3 = ((np.array([iz, iy, ix], dtype=np.float64) + 0.5 * _MID2[row[k + v]]) * sp for v in range(3))
                    vol6 += float(np.dot(np.cross(a, b), c3))                    # cshape.c:143-149
                    area += 0.5 * float(np.linalg.norm(np.cross(a - c3, b - c3)))    # cshape.c:158-178
    dia = [0.0, 0.0, 0.0, 0.0]
    if verts:
        v = np.array(verts)
        for i in range(len(v)):                                       # cshape.c:192-242
            d = v[i] - v[: i + 1]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
            for q in range(3):
                sel = d2[v[: i + 1, q] == v[i, q]]
                if sel.size:
                    dia[q] = max(dia[q], float(sel.max()))
            dia[3] = max(dia[3], float(d2.max()))
    return area, vol6 / 6.0, tuple(float(np.sqrt(x)) for x in dia)


def features(mask, spacing_zyx):
    """the 14 active + 3 deprecated shape features of a (not yet padded) ROI mask, shape.py:54-424"""
    sp = np.asarray(spacing_zyx, dtype=np.float64)
    m = np.pad(np.asarray(mask) != 0, 1)
    sa, vol, dia = coefficients(m, sp)
    idx = np.array(np.where(m), dtype=np.float64).T
    n = len(idx)
    phys = idx * sp[None, :]
    phys -= phys.mean(0)
    phys /= np.sqrt(n)
    ev = np.linalg.eigvals(phys.T.copy() @ phys)
    ev[(ev < 0) & (ev > -1e-10)] = 0
    ev = np.sort(ev)
    f = {"MeshVolume": vol, "VoxelVolume": n * float(np.prod(sp)), "SurfaceArea": sa, "SurfaceVolumeRatio": sa / vol,
         "Sphericity": (36 * np.pi * vol ** 2) ** (1.0 / 3.0) / sa,
         "Compactness1": vol / (sa ** 1.5 * np.sqrt(np.pi)), "Compactness2": 36.0 * np.pi * vol ** 2 / sa ** 3,
         "SphericalDisproportion": sa / (36 * np.pi * vol ** 2) ** (1.0 / 3.0),
         "Maximum3DDiameter": dia[3], "Maximum2DDiameterSlice": dia[0], "Maximum2DDiameterColumn": dia[1],
         "Maximum2DDiameterRow": dia[2]}
    neg = lambda k: ev[k] < 0
    f["MajorAxisLength"] = np.nan if neg(2) else float(np.sqrt(ev[2]) * 4)
    f["MinorAxisLength"] = np.nan if neg(1) else float(np.sqrt(ev[1]) * 4)
    f["LeastAxisLength"] = np.nan if neg(0) else float(np.sqrt(ev[0]) * 4)
    f["Elongation"] = np.nan if (neg(1) or neg(2)) else float(np.sqrt(ev[1] / ev[2]))
    f["Flatness"] = np.nan if (neg(0) or neg(2)) else float(np.sqrt(ev[0] / ev[2]))
    return f
