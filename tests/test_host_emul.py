"""CPU check of the DEVICE arithmetic: pyradiomics_b200/csrc/vox_features.cuh is __host__
__device__, so tests/host_emul/emul.cpp compiles it with g++ and the per-voxel feature math the
CUDA kernels run is compared with the reference's voxel-mode golden maps without a GPU.
(Test-only build; the product never runs this code on the CPU.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cmatrices_oracle as O
from helpers import assert_maps_close, binned, ref_map, voxel_goldens
from pyradiomics_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {
    "glcm": ["Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast", "Correlation",
             "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id", "Idm", "Idmn", "Idn", "Imc1", "Imc2",
             "InverseVariance", "JointAverage", "JointEnergy", "JointEntropy", "MCC", "MaximumProbability", "SumAverage",
             "SumEntropy", "SumSquares"],
    "glrlm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance", "HighGrayLevelRunEmphasis",
              "LongRunEmphasis", "LongRunHighGrayLevelEmphasis", "LongRunLowGrayLevelEmphasis", "LowGrayLevelRunEmphasis",
              "RunEntropy", "RunLengthNonUniformity", "RunLengthNonUniformityNormalized", "RunPercentage", "RunVariance",
              "ShortRunEmphasis", "ShortRunHighGrayLevelEmphasis", "ShortRunLowGrayLevelEmphasis"],
    "glszm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance", "HighGrayLevelZoneEmphasis",
              "LargeAreaEmphasis", "LargeAreaHighGrayLevelEmphasis", "LargeAreaLowGrayLevelEmphasis", "LowGrayLevelZoneEmphasis",
              "SizeZoneNonUniformity", "SizeZoneNonUniformityNormalized", "SmallAreaEmphasis", "SmallAreaHighGrayLevelEmphasis",
              "SmallAreaLowGrayLevelEmphasis", "ZoneEntropy", "ZonePercentage", "ZoneVariance"],
    "gldm": ["DependenceEntropy", "DependenceNonUniformity", "DependenceNonUniformityNormalized", "DependenceVariance",
             "GrayLevelNonUniformity", "GrayLevelVariance", "HighGrayLevelEmphasis", "LargeDependenceEmphasis",
             "LargeDependenceHighGrayLevelEmphasis", "LargeDependenceLowGrayLevelEmphasis", "LowGrayLevelEmphasis",
             "SmallDependenceEmphasis", "SmallDependenceHighGrayLevelEmphasis", "SmallDependenceLowGrayLevelEmphasis"],
    "ngtdm": ["Busyness", "Coarseness", "Complexity", "Contrast", "Strength"],
}


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "host_emul", "libemul.so")
    src = os.path.join(HERE, "host_emul", "emul.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, src])
    return C.CDLL(so)


def alive_mask_bruteforce(lev, centers, ang, r3):
    """which GLCM angles have a co-occurrence inside at least one kernel window (numpy restatement
    of the reference's 'delete empty angles', radiomics/glcm.py:187-196)."""
    m = lev != 0
    out = np.zeros(_lib.ALIVE_WORDS, np.uint32)
    for ai, a in enumerate(ang):
        pm = np.zeros_like(m)
        src = tuple(slice(max(0, -a[d]), lev.shape[d] - max(0, a[d])) for d in range(3))
        dst = tuple(slice(max(0, a[d]), lev.shape[d] + min(0, a[d])) for d in range(3))
        pm[src] = m[src] & m[dst]
        for c in zip(*np.where(centers)):
            lo = [max(c[d] - r3[d], c[d] - r3[d] - a[d], 0) for d in range(3)]
            hi = [min(c[d] + r3[d], c[d] + r3[d] - a[d], lev.shape[d] - 1) for d in range(3)]
            if all(lo[d] <= hi[d] for d in range(3)) and pm[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1].any():
                out[ai >> 5] |= np.uint32(1 << (ai & 31))
                break
    return out


def test_feature_name_tables_match_library_order():
    for cls, names in NAMES.items():
        assert names == sorted(names, key=lambda s: s) or cls == "glcm"  # 'MCC' < 'Ma...' in ASCII
        assert len(names) == {"glcm": 24, "glrlm": 16, "glszm": 16, "gldm": 14, "ngtdm": 5}[cls]


@pytest.mark.parametrize("name,z,kw", voxel_goldens(), ids=[g[0] for g in voxel_goldens()])
def test_device_math_on_host_matches_reference_maps(emul, name, z, kw):
    lev, levels, Ng = binned(z, kw)
    lev16 = np.ascontiguousarray(lev, dtype=np.uint16)
    s = _lib.make_settings(Ng, len(levels), spacing_zyx=z["spacing"][::-1], **kw)
    Zs, Ys, Xs = lev.shape
    ang = O.generate_angles(lev.shape, kw.get("distances", [1]), 0, s.force2D, s.force2Ddimension)
    r3 = [0 if (s.force2D and s.force2Ddimension == k) else s.kernelRadius for k in range(3)]
    alive = alive_mask_bruteforce(lev, z["mask"], ang, r3)
    for cid, cname in enumerate(_lib.CLASSES):
        out = np.zeros((len(NAMES[cname]), Zs, Ys, Xs))
        rc = emul.emul_voxel_features(cid, lev16.ctypes.data_as(C.c_void_p), None, Zs, Ys, Xs, C.byref(s),
                                      alive.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        for k, f in enumerate(NAMES[cname]):
            assert_maps_close(out[k], ref_map(z, cname, f), f"{name}/{cname}/{f}", rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("name,r", [("r1", 1), ("r2", 2)])
def test_firstorder_device_math_on_host(emul, name, r):
    """first-order window statistics (csrc/firstorder.cuh) against the reference's voxel-mode run"""
    import firstorder_np as FO
    import pipeline as PL
    z = np.load(os.path.join(HERE, "golden", "voxel_firstorder.npz"))
    img, m = z["image"], z[name + "_mask"]
    lev, _, _, _ = PL.bin_image(img, m, 25)
    lev16 = np.ascontiguousarray(np.where(m, lev, 0), dtype=np.uint16)
    imgd = np.ascontiguousarray(img, dtype=np.float64)
    mk = np.ascontiguousarray(m, dtype=np.uint8)
    Zs, Ys, Xs = img.shape
    out = np.zeros((18, Zs, Ys, Xs))
    emul.emul_firstorder(imgd.ctypes.data_as(C.c_void_p), mk.ctypes.data_as(C.c_void_p), lev16.ctypes.data_as(C.c_void_p),
                         Zs, Ys, Xs, r, r, r, C.c_double(100.0), C.c_double(float(np.prod(z["spacing"]))),
                         out.ctypes.data_as(C.c_void_p))
    ref = FO.extract(img, m, voxelBased=True, spacing_xyz=z["spacing"], kernelRadius=r, binWidth=25, voxelArrayShift=100)
    for k, f in enumerate(FO.NAMES):
        assert np.allclose(out[k][m], ref[f], rtol=1e-10, atol=1e-9), f
        if f not in ("Entropy", "Uniformity"):
            assert np.allclose(out[k][m], z[f"{name}_{f}"][m], rtol=1e-9, atol=1e-8), f


@pytest.mark.parametrize("kind", ["uniform", "smooth"])
def test_glcm_fast_math_equals_generic_math_on_host(emul, kind):
    """the r=1 GLCM fast path (sorting networks, Lanczos eigen-tasks with float-stored vectors) against the
    generic entry-list / Householder path on a 24^3 volume with holes -- catches solver regressions
    (e.g. dropping the local re-orthogonalisation produced a 6e-5 Ritz error) without a GPU."""
    rng = np.random.default_rng(2)
    shape = (24, 24, 24)
    if kind == "uniform":
        lev = rng.integers(1, 33, shape)
    else:
        zz, yy, xx = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        f = np.sin(zz / 2.7) + np.cos(yy / 3.1) + np.sin(xx / 2.3 + 1) + 0.25 * rng.normal(size=shape)
        lev = np.digitize(f, np.quantile(f, np.linspace(0, 1, 33)[1:-1])) + 1
        lev[5:9, 3:20, 7] = 0
    lev = np.ascontiguousarray(lev, dtype=np.uint16)
    s = _lib.make_settings(32, 32)
    Zs, Ys, Xs = shape
    fast = np.zeros((24, Zs, Ys, Xs))
    gen = np.zeros((24, Zs, Ys, Xs))
    assert emul.emul_glcm_fast(lev.ctypes.data_as(C.c_void_p), Zs, Ys, Xs, C.byref(s), None, fast.ctypes.data_as(C.c_void_p)) == 0
    assert emul.emul_voxel_features(0, lev.ctypes.data_as(C.c_void_p), None, Zs, Ys, Xs, C.byref(s), None, gen.ctypes.data_as(C.c_void_p)) == 0
    for k, f in enumerate(NAMES["glcm"]):
        atol = 1e-6 if f in ("MCC", "Imc2", "Imc1") else 1e-9
        assert np.allclose(fast[k], gen[k], rtol=1e-7, atol=atol, equal_nan=True), f
