"""Shape path on the CPU box: the generated marching-cubes table and the oracle restatement against the
reference's outputs (tests/golden/shape_*.{npz,json}, produced by the compiled reference `_cshape` and
the reference RadiomicsShape class, tests/golden/make_golden.py --shape-only)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import shape_np  # noqa: E402

G = os.path.join(HERE, "golden")


def test_table_reproduces_every_single_cube_probe():
    pr = np.load(os.path.join(G, "shape_cube_probes.npz"))
    for cfg in range(256):
        m = np.zeros((2, 2, 2), dtype=bool)
        for i in range(8):
            if cfg >> i & 1:
                m[i >> 2 & 1, i >> 1 & 1, i & 1] = True
        for k, sp in enumerate(pr["spacings"]):
            sa, vol, _ = shape_np.coefficients(m, sp)
            assert sa == pytest.approx(pr["probes"][cfg, k, 0], rel=1e-12, abs=1e-13), (cfg, k)
            assert vol == pytest.approx(pr["probes"][cfg, k, 1], rel=1e-12, abs=1e-13), (cfg, k)


def test_table_is_well_formed():
    mids, tri = shape_np._MID2, shape_np._TRI
    assert mids.shape == (12, 3) and tri.shape == (256, 16)
    assert (tri[0] == -1).all() and (tri[255] == -1).all()
    for cfg in range(256):
        row = tri[cfg]
        n = int((row >= 0).sum())
        assert n % 3 == 0 and (row[n:] == -1).all()
        for e in row[:n]:                       # a triangle corner sits on an edge whose ends differ in the mask
            a = [int(v) for v in (mids[e] // 2)]
            b = [int(v) for v in ((mids[e] + 1) // 2)]
            ia, ib = a[0] << 2 | a[1] << 1 | a[2], b[0] << 2 | b[1] << 1 | b[2]
            assert (cfg >> ia & 1) != (cfg >> ib & 1), (cfg, e)


@pytest.mark.parametrize("name", ["blob", "noise", "sparse", "touching_border", "single", "plane"])
def test_oracle_coefficients_match_reference(name):
    d = np.load(os.path.join(G, "shape_random.npz"))
    sa, vol, dia = shape_np.coefficients(d[f"{name}_mask"], d[f"{name}_spacing"])
    ref = d[f"{name}_coeff"]
    assert sa == pytest.approx(ref[0], rel=1e-11, abs=1e-12)
    assert vol == pytest.approx(ref[1], rel=1e-11, abs=1e-9)
    assert list(dia) == [pytest.approx(v, rel=1e-15, abs=0) for v in ref[2:6]]


@pytest.mark.parametrize("case", ["brain2", "breast1", "lung1"])
def test_oracle_features_match_reference_class(case):
    exp = json.load(open(os.path.join(G, "shape_expect.json")))[case]
    seg = np.load(os.path.join(G, "segment_cases.npz"))
    got = shape_np.features(seg[f"{case}_mask"], seg[f"{case}_spacing"][::-1])
    for f, v in exp["features"].items():
        assert got[f] == pytest.approx(v, rel=1e-9), f
    for f, v in exp["baseline"].items():          # the stored CSV, at the reference's own 3 % (tests/testUtils.py:266-275)
        assert got[f] == pytest.approx(v, rel=0.03), f


def test_committed_table_is_what_the_generator_derives():
    """csrc/mc_table.inc is the output of gen_mc_table.build() on the committed probes (geometry + black-box selection),
    not a hand-edited or transcribed table."""
    import gen_mc_table as g
    table = g.build()
    mids, tri = g.load_table()
    assert [tuple(m) for m in mids] == list(g.EDGE_MID2)
    for cfg, tris in enumerate(table):
        flat = [e for t in tris for e in t]
        assert list(tri[cfg][:len(flat)]) == flat and (tri[cfg][len(flat):] == -1).all(), cfg


SHAPE2D = ["disc", "noise", "sparse", "touching_border", "single", "checker", "ring"]


@pytest.mark.parametrize("name", SHAPE2D)
def test_shape2d_restatement_matches_compiled_reference_goldens(name):
    """oracle/shape_np.coefficients2d against the reference's calculate_coefficients2D (tests/golden/shape2d_golden.npz,
    written by make_golden.py --shape2d-only from the compiled _cshape)"""
    import shape_np as S
    d = np.load(os.path.join(G, "shape2d_golden.npz"))
    per, sur, dia = S.coefficients2d(np.pad(d[name + "_mask"], 1), d[name + "_spacing"])
    ref = d[name + "_coeff"]
    assert per == pytest.approx(ref[0], rel=1e-13) and sur == pytest.approx(ref[1], rel=1e-12, abs=1e-14)
    assert dia == ref[2]
