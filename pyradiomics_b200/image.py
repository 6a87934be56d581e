"""Tiny image abstraction so the plugin classes accept what pyradiomics passes (SimpleITK images)
as well as plain NumPy arrays: SimpleITK is used when it is importable, otherwise ``ArrayImage``
carries the array + spacing with the handful of methods the feature classes touch
(reference radiomics/base.py:86,207-208,242-245)."""
from __future__ import annotations

import numpy as np

try:  # optional
    import SimpleITK as _sitk
    if not hasattr(_sitk, "ReadImage"):      # a stub module (tests) is not the real thing
        _sitk = None
except ImportError:  # pragma: no cover
    _sitk = None


class ArrayImage:
    """(z,y,x) NumPy array + spacing in SimpleITK (x,y,z) order."""

    def __init__(self, array, spacing=None, origin=None):
        self.array = np.asarray(array)
        nd = self.array.ndim
        self.spacing = tuple(float(s) for s in (spacing if spacing is not None else (1.0,) * nd))
        self.origin = tuple(float(o) for o in (origin if origin is not None else (0.0,) * nd))

    def GetOrigin(self):
        return self.origin

    def GetSize(self):
        return tuple(int(s) for s in self.array.shape[::-1])

    def GetSpacing(self):
        return self.spacing

    def GetDimension(self):
        return self.array.ndim

    def CopyInformation(self, other):
        self.spacing = tuple(other.GetSpacing())
        if hasattr(other, "GetOrigin"):
            self.origin = tuple(other.GetOrigin())


def as_array(img):
    if isinstance(img, ArrayImage):
        return img.array
    if isinstance(img, np.ndarray):
        return img
    if _sitk is not None and isinstance(img, _sitk.Image):
        return _sitk.GetArrayFromImage(img)
    if hasattr(img, "_arr"):                 # array-backed stand-ins
        return np.asarray(img._arr)
    raise TypeError(f"cannot interpret {type(img)} as an image")


def spacing_xyz(img):
    if isinstance(img, np.ndarray):
        return (1.0,) * img.ndim
    return tuple(img.GetSpacing())


def origin_xyz(img):
    if isinstance(img, np.ndarray) or not hasattr(img, "GetOrigin"):
        return (0.0,) * np.ndim(as_array(img))
    return tuple(img.GetOrigin())


def size_xyz(img):
    return tuple(int(s) for s in as_array(img).shape[::-1])


def like(ref, array):
    """wrap `array` as an image of the same kind / geometry as `ref`."""
    if _sitk is not None and isinstance(ref, _sitk.Image):
        out = _sitk.GetImageFromArray(array)
        out.CopyInformation(ref)
        return out
    return ArrayImage(array, spacing_xyz(ref), origin_xyz(ref))
