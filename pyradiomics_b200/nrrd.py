"""Feature-map output assembly: NRRD files like the ones the reference's
voxel script writes with ``sitk.WriteImage(map, target, True)`` (radiomics/scripts/voxel.py:66-72), without
building a SimpleITK image per map.  75 float64 maps of a 512^3 case are 80 GB: ``write_maps`` streams
them map by map from the (pinned) host buffers of ``voxel.HostExtractor`` and compresses several maps
concurrently (zlib releases the GIL), so compression overlaps the next case's GPU work.

NRRD0004, little-endian, ``encoding: gzip`` (or ``raw``), ``sizes`` in x y z order with the array stored
z-major exactly as NumPy holds it -- what ITK's NrrdImageIO reads back as the same (z, y, x) array.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import zlib

import numpy as np

_TYPES = {"float64": "double", "float32": "float", "int16": "short", "int32": "int", "uint8": "unsigned char",
          "uint16": "unsigned short", "int8": "signed char", "uint32": "unsigned int", "int64": "int64", "uint64": "uint64"}


def header(shape_zyx, dtype, spacing_xyz=(1.0, 1.0, 1.0), origin_xyz=(0.0, 0.0, 0.0), encoding="gzip") -> bytes:
    nd = len(shape_zyx)
    dirs = []
    for d in range(nd):
        v = ["0"] * nd
        v[d] = repr(float(spacing_xyz[d]))
        dirs.append("(" + ",".join(v) + ")")
    lines = ["NRRD0004", "# written by pyradiomics_b200.nrrd", f"type: {_TYPES[np.dtype(dtype).name]}", f"dimension: {nd}",
             "space: left-posterior-superior" if nd == 3 else f"space dimension: {nd}",
             "sizes: " + " ".join(str(int(s)) for s in shape_zyx[::-1]),
             "space directions: " + " ".join(dirs), "kinds: " + " ".join(["domain"] * nd), "endian: little",
             f"encoding: {encoding}", "space origin: (" + ",".join(repr(float(o)) for o in origin_xyz[:nd]) + ")"]
    return ("\n".join(lines) + "\n\n").encode("ascii")


def write_nrrd(path, array, spacing_xyz=(1.0, 1.0, 1.0), origin_xyz=(0.0, 0.0, 0.0), compress=True, level=1,
               chunk_bytes=64 << 20):
    """one array (z, y, x) -> one .nrrd file.  gzip member written in chunks so that no second full copy of a
    multi-GB map is ever held."""
    a = np.asarray(array)
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    a = np.ascontiguousarray(a)
    with open(path, "wb") as f:
        f.write(header(a.shape, a.dtype, spacing_xyz, origin_xyz, "gzip" if compress else "raw"))
        flat = a.reshape(-1).view(np.uint8)
        if not compress:
            f.write(memoryview(flat))
            return path
        co = zlib.compressobj(level, zlib.DEFLATED, 31)            # wbits 31: gzip container
        for i in range(0, flat.size, chunk_bytes):
            f.write(co.compress(memoryview(flat[i:i + chunk_bytes])))
        f.write(co.flush())
    return path


def write_maps(out_dir, maps, feature_names, spacing_xyz=(1.0, 1.0, 1.0), origin_xyz=(0.0, 0.0, 0.0), prefix="original",
               compress=True, level=1, workers=8):
    """maps: {class: array-like [F, Z, Y, X]} (e.g. the return value of HostExtractor.run, torch or NumPy);
    feature_names: {class: [F names]}.  Writes <prefix>_<class>_<Feature>.nrrd (the reference's feature keys,
    featureextractor.py:626) and returns {feature key: path}."""
    os.makedirs(out_dir, exist_ok=True)
    jobs = {}
    with cf.ThreadPoolExecutor(max_workers=max(1, int(workers))) as ex:
        for cname, arr in maps.items():
            a = arr.numpy() if hasattr(arr, "numpy") else np.asarray(arr)
            for k, fname in enumerate(feature_names[cname]):
                key = f"{prefix}_{cname}_{fname}"
                path = os.path.join(out_dir, key + ".nrrd")
                jobs[key] = ex.submit(write_nrrd, path, a[k], spacing_xyz, origin_xyz, compress, level)
        return {k: j.result() for k, j in jobs.items()}
