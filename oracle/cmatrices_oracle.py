"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/cmatrices_port.c with the call
signatures of the reference's ``radiomics.cMatrices`` module (reference
radiomics/src/_cmatrices.c:41-50,104,255,450,601,749,892), so parity tests read like the
reference's.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liboracle.so")
_SRC = os.path.join(HERE, "cmatrices_port.c")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, _SRC])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_glszm_zones.restype = C.c_int
        _lib.oracle_free.argtypes = [C.c_void_p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _prep(image, mask):
    img = np.ascontiguousarray(image, dtype=np.int32)
    msk = np.ascontiguousarray(mask, dtype=np.bool_).view(np.int8)
    if img.ndim != msk.ndim:
        raise ValueError("Expected image and mask to have equal number of dimensions.")
    if img.shape != msk.shape:
        raise ValueError("Dimensions of image and mask do not match.")
    if img.ndim not in (2, 3):
        raise ValueError("oracle port handles 2-D and 3-D arrays")
    nd = img.ndim
    size3 = np.array((1,) * (3 - nd) + img.shape, dtype=np.int32)
    return img, msk, nd, size3


def generate_angles(size, distances, bidirectional, force2D, force2Ddimension):
    size = np.asarray(size, dtype=np.int32)
    nd = size.shape[0]
    size3 = np.concatenate([np.ones(3 - nd, np.int32), size]).astype(np.int32)
    dist = np.ascontiguousarray(distances, dtype=np.int32)
    f2 = (force2Ddimension + 3 - nd) if force2D else -1
    buf = np.zeros((max(1, (2 * int(dist.max()) + 1) ** 3), 3), np.int32)
    na = lib().oracle_build_angles(_p(size3, C.c_int), _p(dist, C.c_int), C.c_int(dist.size),
                                   C.c_int(1 if bidirectional else 0), C.c_int(f2),
                                   _p(buf, C.c_int), C.c_int(buf.shape[0]))
    if na <= 0:
        raise RuntimeError("Error getting angle count.")
    return np.ascontiguousarray(buf[:na, 3 - nd:])


def _voxels(voxels, nd, kernelRadius):
    if voxels is None:
        return None, 1
    if kernelRadius <= 0:
        raise RuntimeError("Expecting kernelRadius > 0")
    v = np.ascontiguousarray(voxels, dtype=np.int32)
    if v.ndim != 2 or v.shape[0] != nd:
        raise RuntimeError("Expecting voxel indices array to be 2-dimensional")
    if nd == 2:
        v = np.concatenate([np.zeros((1, v.shape[1]), np.int32), v]).copy()
    return v, v.shape[1]


def _angles3(ang, nd):
    if nd == 3:
        return np.ascontiguousarray(ang, dtype=np.int32)
    return np.ascontiguousarray(np.concatenate([np.zeros((ang.shape[0], 1), np.int32), ang], 1), dtype=np.int32)


def _f2(force2D, dim, nd):
    return (dim + 3 - nd) if force2D else -1


def calculate_glcm(image, mask, distances, Ng, force2D, force2Ddimension, kernelRadius=0, voxels=None):
    img, msk, nd, size3 = _prep(image, mask)
    ang = generate_angles(img.shape, distances, 0, force2D, force2Ddimension)
    v, nvox = _voxels(voxels, nd, kernelRadius)
    out = np.zeros((nvox, Ng, Ng, ang.shape[0]))
    a3 = _angles3(ang, nd)
    ok = lib().oracle_glcm(_p(img, C.c_int), _p(msk, C.c_char), _p(size3, C.c_int), _p(a3, C.c_int),
                           C.c_int(a3.shape[0]), C.c_int(Ng), None if v is None else _p(v, C.c_int),
                           C.c_int(nvox), C.c_int(kernelRadius), C.c_int(_f2(force2D, force2Ddimension, nd)),
                           _p(out, C.c_double))
    if not ok:
        raise IndexError("Calculation of GLCM Failed.")
    return out, ang


def calculate_glrlm(image, mask, Ng, Nr, force2D, force2Ddimension, kernelRadius=0, voxels=None):
    img, msk, nd, size3 = _prep(image, mask)
    ang = generate_angles(img.shape, [1], 0, force2D, force2Ddimension)
    v, nvox = _voxels(voxels, nd, kernelRadius)
    out = np.zeros((nvox, Ng, Nr, ang.shape[0]))
    a3 = _angles3(ang, nd)
    ok = lib().oracle_glrlm(_p(img, C.c_int), _p(msk, C.c_char), _p(size3, C.c_int), _p(a3, C.c_int),
                            C.c_int(a3.shape[0]), C.c_int(Ng), C.c_int(Nr),
                            None if v is None else _p(v, C.c_int), C.c_int(nvox), C.c_int(kernelRadius),
                            C.c_int(_f2(force2D, force2Ddimension, nd)), _p(out, C.c_double))
    if not ok:
        raise IndexError("Calculation of GLRLM Failed.")
    return out, ang


def calculate_glszm(image, mask, Ng, Ns, force2D, force2Ddimension, kernelRadius=0, voxels=None):
    img, msk, nd, size3 = _prep(image, mask)
    ang = generate_angles(img.shape, [1], 1, force2D, force2Ddimension)
    v, nvox = _voxels(voxels, nd, kernelRadius)
    a3 = _angles3(ang, nd)
    offs = np.zeros(nvox + 1, dtype=np.int64)
    zptr = C.POINTER(C.c_int)()
    mx = lib().oracle_glszm_zones(_p(img, C.c_int), _p(msk, C.c_char), _p(size3, C.c_int), _p(a3, C.c_int),
                                  C.c_int(a3.shape[0]), None if v is None else _p(v, C.c_int),
                                  C.c_int(nvox), C.c_int(kernelRadius),
                                  C.c_int(_f2(force2D, force2Ddimension, nd)), _p(offs, C.c_long), C.byref(zptr))
    if mx < 0:
        raise IndexError("Calculation of GLSZM Failed.")
    mx = max(mx, 1)
    out = np.zeros((nvox, Ng, mx))
    ok = lib().oracle_glszm_fill(zptr, _p(offs, C.c_long), C.c_int(nvox), C.c_int(Ng), C.c_int(mx), _p(out, C.c_double))
    lib().oracle_free(zptr)
    if not ok:
        raise IndexError("Error filling GLSZM.")
    return out


def calculate_ngtdm(image, mask, distances, Ng, force2D, force2Ddimension, kernelRadius=0, voxels=None):
    img, msk, nd, size3 = _prep(image, mask)
    ang = generate_angles(img.shape, distances, 1, force2D, force2Ddimension)
    v, nvox = _voxels(voxels, nd, kernelRadius)
    out = np.zeros((nvox, Ng, 3))
    a3 = _angles3(ang, nd)
    ok = lib().oracle_ngtdm(_p(img, C.c_int), _p(msk, C.c_char), _p(size3, C.c_int), _p(a3, C.c_int),
                            C.c_int(a3.shape[0]), C.c_int(Ng), None if v is None else _p(v, C.c_int),
                            C.c_int(nvox), C.c_int(kernelRadius), C.c_int(_f2(force2D, force2Ddimension, nd)),
                            _p(out, C.c_double))
    if not ok:
        raise IndexError("Calculation of NGTDM Failed.")
    return out


def calculate_gldm(image, mask, distances, Ng, alpha, force2D, force2Ddimension, kernelRadius=0, voxels=None):
    img, msk, nd, size3 = _prep(image, mask)
    ang = generate_angles(img.shape, distances, 1, force2D, force2Ddimension)
    v, nvox = _voxels(voxels, nd, kernelRadius)
    out = np.zeros((nvox, Ng, 2 * ang.shape[0] + 1))
    a3 = _angles3(ang, nd)
    ok = lib().oracle_gldm(_p(img, C.c_int), _p(msk, C.c_char), _p(size3, C.c_int), _p(a3, C.c_int),
                           C.c_int(a3.shape[0]), C.c_int(Ng), C.c_int(int(alpha)),
                           None if v is None else _p(v, C.c_int), C.c_int(nvox), C.c_int(kernelRadius),
                           C.c_int(_f2(force2D, force2Ddimension, nd)), _p(out, C.c_double))
    if not ok:
        raise IndexError("Calculation of GLDM Failed.")
    return out
