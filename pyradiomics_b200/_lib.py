"""ctypes binding of libb200radiomics.so (include/b200radiomics.h).  There is no fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200_RADIOMICS_LIB: developer override to A/B a differently-compiled build of the same library
LIB_PATH = os.environ.get("B200_RADIOMICS_LIB") or os.path.join(HERE, "libb200radiomics.so")

RB_OK, RB_ERR_CUDA, RB_ERR_LEVEL_RANGE, RB_ERR_ARG, RB_ERR_NOMEM, RB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
CLASSES = ("glcm", "glrlm", "glszm", "gldm", "ngtdm")
CLASS_ID = {n: i for i, n in enumerate(CLASSES)}
WEIGHTING = {None: 0, "infinity": 1, "euclidean": 2, "manhattan": 3, "no_weighting": 4}
ALIVE_WORDS = 6


class VoxelSettings(C.Structure):
    _fields_ = [
        ("kernelRadius", C.c_int), ("force2D", C.c_int), ("force2Ddimension", C.c_int),
        ("ndist", C.c_int), ("distances", C.c_int * 8), ("symmetricalGLCM", C.c_int),
        ("weighting", C.c_int), ("spacing_zyx", C.c_double * 3), ("gldm_a", C.c_int),
        ("initValue", C.c_double), ("Ng", C.c_int), ("n_roi_levels", C.c_int),
    ]


class B200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is not built -- run `python -m pyradiomics_b200.build` (needs nvcc); "
                "pyradiomics_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rb_last_error.restype = C.c_char_p
        L.rb_version.restype = C.c_char_p
        L.rb_feature_name.restype = C.c_char_p
        L.rb_feature_name.argtypes = [C.c_int, C.c_int]
        _lib = L
    return _lib


def check(rc, what=""):
    """Map rb_status to the exception types the reference binding raises
    (reference radiomics/src/_cmatrices.c:159,219,1045,1079,1093)."""
    if rc >= 0:
        return rc
    msg = (lib().rb_last_error() or b"").decode()
    if rc == RB_ERR_LEVEL_RANGE:
        raise IndexError(f"Calculation of {what or 'matrix'} Failed. ({msg})")
    if rc == RB_ERR_ARG:
        raise ValueError(msg)
    if rc == RB_ERR_NOMEM:
        raise MemoryError(msg)
    raise B200Error(f"{what}: rb_status {rc}: {msg}")


def feature_names(cls):
    cid = CLASS_ID[cls] if isinstance(cls, str) else cls
    n = lib().rb_num_features(cid)
    return [lib().rb_feature_name(cid, i).decode() for i in range(n)]


def make_settings(Ng, n_roi_levels, **kw):
    s = VoxelSettings()
    s.kernelRadius = int(kw.get("kernelRadius", 1))
    s.force2D = int(bool(kw.get("force2D", False)))
    s.force2Ddimension = int(kw.get("force2Ddimension", 0))
    d = [int(x) for x in kw.get("distances", [1])]
    if not 1 <= len(d) <= 8:
        raise ValueError("1..8 distances supported")
    s.ndist = len(d)
    for i, v in enumerate(d):
        s.distances[i] = v
    s.symmetricalGLCM = int(bool(kw.get("symmetricalGLCM", True)))
    wn = kw.get("weightingNorm")
    s.weighting = WEIGHTING.get(wn, 4)  # unknown names weigh 1 like the reference (glcm.py:176-181)
    sp = kw.get("spacing_zyx", (1.0, 1.0, 1.0))
    for i in range(3):
        s.spacing_zyx[i] = float(sp[i])
    s.gldm_a = int(kw.get("gldm_a", 0))
    s.initValue = float(kw.get("initValue", 0))
    s.Ng = int(Ng)
    s.n_roi_levels = int(n_roi_levels)
    return s
