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


def coefficients2d(mask, spacing):
    """CPU restatement of calculate_coefficients2D + calculate_meshDiameter2D (radiomics/src/cshape.c:420-595) for an
    already zero-padded 2-D mask: marching squares with edge-midpoint vertices.  Per 2x2 neighbourhood the reference's
    16-entry line table amounts to: 1 or 3 inside corners -> one corner cut (:460-500); two adjacent corners -> one
    straight cut; the two diagonal configurations -> two corner cuts that keep the inside corners apart (probed on the
    compiled reference: [[1,0],[0,1]] has surface 1.0).  The surface is accumulated like the reference does, as half the
    sum of the cross products of the ORIENTED segment end points (:483), and the vertices kept for the diameter are the
    midpoints of the crossed left / bottom square edges (:514-535).  Pinned on the compiled reference (_cshape) by
    tests/golden/shape2d_golden.npz."""
    m = np.asarray(mask) != 0
    sy, sx = float(spacing[0]), float(spacing[1])
    Y, X = m.shape
    per = 0.0
    cross = 0.0
    verts = []
    mid = {0: (0.0, 0.5), 1: (0.5, 1.0), 2: (1.0, 0.5), 3: (0.5, 0.0)}        # edge midpoints: top, right, bottom, left
    corner = [(0, 0), (0, 1), (1, 1), (1, 0)]                                  # p0..p3, clockwise from the origin
    for iy in range(Y - 1):
        for ix in range(X - 1):
            ins = [bool(m[iy + dy, ix + dx]) for dy, dx in corner]
            k = sum(ins)
            if k in (0, 4):
                continue
            # segments: walk the square's edges clockwise (edge e joins corner e and e+1); a segment starts where the
            # boundary goes inside -> outside and ends at the next outside -> inside crossing, so the inside stays on
            # one side of every oriented segment (the diagonal cases pair each inside corner with its own two edges)
            outs = [e for e in range(4) if ins[e] and not ins[(e + 1) % 4]]
            for e in outs:
                f = e
                while True:
                    f = (f + 1) % 4
                    if not ins[f] and ins[(f + 1) % 4]:
                        break
                if k == 2 and ins[0] == ins[2]:          # diagonal: keep the corners apart -> the closing edge is the one BEFORE e
                    f = (e - 1) % 4
                a = ((iy + mid[e][0]) * sy, (ix + mid[e][1]) * sx)
                b = ((iy + mid[f][0]) * sy, (ix + mid[f][1]) * sx)
                cross += a[0] * b[1] - b[0] * a[1]
                per += float(np.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2))
            if ins[0] != ins[3]:
                verts.append(((iy + 0.5) * sy, ix * sx))
            if ins[3] != ins[2]:
                verts.append(((iy + 1.0) * sy, (ix + 0.5) * sx))
    v = np.array(verts, dtype=np.float64).reshape(-1, 2)
    d2 = 0.0
    for i in range(len(v)):
        if i:
            dy = v[i, 0] - v[:i, 0]
            dx = v[i, 1] - v[:i, 1]
            d2 = max(d2, float((dy * dy + dx * dx).max()))
    return per, abs(cross) / 2.0, float(np.sqrt(d2))


def features2d(mask, spacing_yx):
    """the 9 active + 1 deprecated 2-D shape features of a (not yet padded) 2-D ROI mask, shape2D.py:40-300"""
    sp = np.asarray(spacing_yx, dtype=np.float64)
    m = np.pad(np.asarray(mask) != 0, 1)
    per, sur, dia = coefficients2d(m, sp)
    idx = np.array(np.where(m), dtype=np.float64).T
    n = len(idx)
    phys = idx * sp[None, :]
    phys -= phys.mean(0)
    phys /= np.sqrt(n)
    ev = np.linalg.eigvals(phys.T.copy() @ phys).real
    ev[(ev < 0) & (ev > -1e-10)] = 0
    ev = np.sort(ev)
    sph = (2 * np.sqrt(np.pi * sur)) / per
    return {"MeshSurface": sur, "PixelSurface": n * float(np.prod(sp)), "Perimeter": per, "PerimeterSurfaceRatio": per / sur,
            "Sphericity": sph, "SphericalDisproportion": 1.0 / sph, "MaximumDiameter": dia,
            "MajorAxisLength": np.nan if ev[1] < 0 else float(np.sqrt(ev[1]) * 4),
            "MinorAxisLength": np.nan if ev[0] < 0 else float(np.sqrt(ev[0]) * 4),
            "Elongation": np.nan if (ev[0] < 0 or ev[1] < 0) else float(np.sqrt(ev[0] / ev[1]))}
