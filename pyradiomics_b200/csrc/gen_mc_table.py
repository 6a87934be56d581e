"""Generates csrc/mc_table.inc: the marching-cubes triangle table of the shape path.

The reference meshes the ROI with a 2x2x2 marching cube whose iso-surface vertices sit on the
midpoints of cube edges (radiomics/src/cshape.c:22-190, table at cshape.c:250-619).  This file does
NOT transcribe that table.  It rebuilds the triangulation from geometry:

  1. corners i <-> (z, y, x) = (i>>2 & 1, i>>1 & 1, i & 1); 12 edges between corners differing in one bit;
  2. on every cube face the active edge midpoints are joined by segments (1 or 3 inside corners: one
     segment; 2 adjacent: one; 2 diagonal: two segments, either pairing -- the ambiguous face);
  3. the segments close into polygons; every polygon is triangulated (all triangulations are candidates)
     and oriented with its normals pointing from the inside corners to the outside ones (shape.py:24-27);
  4. the candidate whose (surface area, signed origin volume) equals the reference's output for that
     single-cube mask under three spacings (tests/golden/shape_cube_probes.npz, produced by
     tests/golden/make_golden.py from the compiled reference) is kept -- this pins the pairing on ambiguous
     faces and the split of non-planar polygons, the only freedom a midpoint marching cube has.

Area is additive per cube; the volume of a closed mesh depends on the per-cube triangles only through
their origin volume at offset 0 and their vector area (fixed by the polygon boundaries), so matching the
probes makes whole-mesh parity follow (tests/test_shape_cpu.py re-checks all 256 x 3 probes and random
masks against the reference).

usage: python gen_mc_table.py            (needs tests/golden/shape_cube_probes.npz)
"""
from __future__ import annotations

import itertools
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PROBES = os.path.join(HERE, "..", "..", "tests", "golden", "shape_cube_probes.npz")
OUT = os.path.join(HERE, "mc_table.inc")


def corner(i):
    return (i >> 2 & 1, i >> 1 & 1, i & 1)


# edges: (corner a, corner b) with a < b differing in exactly one bit; id = position in this list
EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]
EDGE_ID = {e: k for k, e in enumerate(EDGES)}
EDGE_MID2 = [tuple(ca + cb for ca, cb in zip(corner(a), corner(b))) for a, b in EDGES]   # midpoint * 2

# faces: axis d fixed at value v; corners in cyclic order
FACES = []
for d in range(3):
    o = [k for k in range(3) if k != d]
    for v in (0, 1):
        cyc = []
        for (p, q) in ((0, 0), (0, 1), (1, 1), (1, 0)):
            c = [0, 0, 0]
            c[d] = v; c[o[0]] = p; c[o[1]] = q
            cyc.append(c[0] << 2 | c[1] << 1 | c[2])
        FACES.append(cyc)


def eid(a, b):
    return EDGE_ID[(min(a, b), max(a, b))]


def face_segments(cfg, cyc):
    """list of alternatives; each alternative is a list of (edge, edge) segments on this face"""
    ins = [cfg >> c & 1 for c in cyc]
    n = sum(ins)
    if n in (0, 4):
        return [[]]
    if n in (1, 3):
        k = ins.index(1) if n == 1 else ins.index(0)
        return [[(eid(cyc[k], cyc[(k + 1) % 4]), eid(cyc[k], cyc[(k - 1) % 4]))]]
    # two inside
    k = ins.index(1)
    if ins[(k + 1) % 4] or ins[(k - 1) % 4]:              # adjacent pair
        if not ins[(k + 1) % 4]:
            k = (k - 1) % 4                               # make k, k+1 the inside pair
        return [[(eid(cyc[k], cyc[(k - 1) % 4]), eid(cyc[(k + 1) % 4], cyc[(k + 2) % 4]))]]
    # diagonal: k and k+2 inside
    a = [(eid(cyc[k], cyc[(k + 1) % 4]), eid(cyc[k], cyc[(k - 1) % 4])),
         (eid(cyc[(k + 2) % 4], cyc[(k + 3) % 4]), eid(cyc[(k + 2) % 4], cyc[(k + 1) % 4]))]
    b = [(eid(cyc[k], cyc[(k + 1) % 4]), eid(cyc[(k + 2) % 4], cyc[(k + 1) % 4])),
         (eid(cyc[k], cyc[(k - 1) % 4]), eid(cyc[(k + 2) % 4], cyc[(k + 3) % 4]))]
    return [a, b]


def polygons(segs):
    adj = {}
    for u, v in segs:
        adj.setdefault(u, []).append(v)
        adj.setdefault(v, []).append(u)
    assert all(len(v) == 2 for v in adj.values()), adj
    seen, polys = set(), []
    for s in sorted(adj):
        if s in seen:
            continue
        poly, prev, cur = [s], None, s
        seen.add(s)
        while True:
            nxt = [w for w in adj[cur] if w != prev] or adj[cur]
            nx = nxt[0]
            if nx == s:
                break
            poly.append(nx); seen.add(nx)
            prev, cur = cur, nx
        polys.append(poly)
    return polys


def triangulations(poly):
    """all triangulations of a polygon given as a vertex list (recursive on the edge poly[0]-poly[-1])"""
    n = len(poly)
    if n < 3:
        return [[]]
    if n == 3:
        return [[tuple(poly)]]
    out = []
    for k in range(1, n - 1):
        for left in triangulations(poly[:k + 1]):
            for right in triangulations(poly[k:]):
                out.append(left + [(poly[0], poly[k], poly[-1])] + right)
    return out


def measure(tris, spacing):
    """(area, origin volume * 6) of a triangle list under a spacing, like cshape.c:121-178"""
    area = vol6 = 0.0
    for t in tris:
        a, b, c = (np.array(EDGE_MID2[e], dtype=float) * 0.5 * spacing for e in t)
        vol6 += float(np.dot(np.cross(a, b), c))
        area += 0.5 * float(np.linalg.norm(np.cross(a - c, b - c)))
    return area, vol6


_TRI_CACHE = {}
_ORIENT_CACHE = {}


def tri_measures(t, spacings):
    """[(area, 6 * origin volume) per spacing] of ONE oriented triangle, cached"""
    key = (t, spacings.tobytes())
    if key not in _TRI_CACHE:
        _TRI_CACHE[key] = np.array([measure([t], sp) for sp in spacings])
    return _TRI_CACHE[key]


def orient_outward(tris, cfg):
    """order every triangle so that its normal (a - c) x (b - c) points from the inside corners to the outside
    ones (the reference's convention, shape.py:24-27): every mesh vertex sits on an edge with one inside and one
    outside corner, and the normal must have a positive component along (outside - inside) summed over the
    triangle's three vertices.  The probes only see the origin volume of a cube at offset 0, which cannot tell a
    polygon from its mirror-oriented twin when two polygons have equal and opposite origin volumes; the whole-mesh
    volume (origin volume + offset . vector area) can."""
    out = []
    for t in tris:
        if (t, cfg) in _ORIENT_CACHE:
            out.append(_ORIENT_CACHE[(t, cfg)])
            continue
        pts = [np.array(EDGE_MID2[e], dtype=float) * 0.5 for e in t]
        nrm = np.cross(pts[0] - pts[2], pts[1] - pts[2])
        d = np.zeros(3)
        for e in t:
            a, b = EDGES[e]
            ins, outs = (a, b) if cfg >> a & 1 else (b, a)
            d += np.array(corner(outs), dtype=float) - np.array(corner(ins), dtype=float)
        _ORIENT_CACHE[(t, cfg)] = t if float(np.dot(nrm, d)) > 0 else (t[0], t[2], t[1])
        out.append(_ORIENT_CACHE[(t, cfg)])
    return out


def build():
    pr = np.load(PROBES)
    spacings, probes = pr["spacings"], pr["probes"]
    table = []
    for cfg in range(256):
        target = np.array([[probes[cfg, k, 0], probes[cfg, k, 1] * 6.0] for k in range(3)])
        found = None
        for choice in itertools.product(*[face_segments(cfg, cyc) for cyc in FACES]):
            segs = [s for alt in choice for s in alt]
            if not segs:
                cand_sets = []
            else:
                cand_sets = []
                for poly in polygons(segs):
                    cands = []
                    for tri in triangulations(poly):
                        tt = orient_outward(tri, cfg)
                        cands.append((tt, sum((tri_measures(t, spacings) for t in tt), np.zeros((len(spacings), 2)))))
                    cand_sets.append(cands)
            for combo in itertools.product(*cand_sets):
                tot = sum((c[1] for c in combo), np.zeros((3, 2)))
                if np.allclose(tot, target, rtol=1e-10, atol=1e-10):
                    found = [t for c in combo for t in c[0]]
                    break
            if found is not None:
                break
        assert found is not None, f"no candidate triangulation reproduces the reference for configuration {cfg}"
        assert len(found) <= 5
        table.append(found)
    return table


def write(table):
    with open(OUT, "w") as f:
        f.write("// GENERATED by gen_mc_table.py -- do not edit.  Corner i = (z,y,x) bits (i>>2, i>>1, i); edge ids below.\n")
        f.write("// MC_EDGE_MID2[e] = 2 * (midpoint of edge e) in (z, y, x); MC_TRI[cfg] = up to 5 triangles of edge ids, -1 ends.\n")
        f.write("static const signed char MC_EDGE_MID2[12][3] = {" + ", ".join("{%d, %d, %d}" % m for m in EDGE_MID2) + "};\n")
        f.write("static const signed char MC_TRI[256][16] = {\n")
        for cfg, tris in enumerate(table):
            flat = [e for t in tris for e in t]
            flat += [-1] * (16 - len(flat))
            f.write("  {" + ", ".join("%d" % v for v in flat) + "},\n")
        f.write("};\n")


def load_table():
    """parse mc_table.inc back (used by the CPU tests)"""
    import re
    txt = open(OUT).read()
    mid = re.search(r"MC_EDGE_MID2\[12\]\[3\] = \{(.*?)\};", txt, re.S).group(1)
    mids = np.array([int(v) for v in re.findall(r"-?\d+", mid)]).reshape(12, 3)
    body = txt[txt.index("MC_TRI[256][16]"):]
    rows = re.findall(r"\{([-\d, ]+)\},", body)
    tri = np.array([[int(v) for v in r.split(",")] for r in rows])
    assert tri.shape == (256, 16)
    return mids, tri


if __name__ == "__main__":
    t = build()
    write(t)
    print("wrote", OUT, "max triangles", max(len(x) for x in t))
