"""Voxel-based (feature-map) extraction on device-resident volumes: the headline path.

Mirrors what the reference does per feature class in voxel-based mode (reference
radiomics/base.py:98-111,200-245 + the per-class _calculateMatrix/_calculateCoefficients/get*
methods) but as one fused CUDA kernel per class that never materialises per-voxel matrices.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CLASS_ID, CLASSES, check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_levels(image: torch.Tensor, mask: torch.Tensor, Ng: int):
    """int32 gray-level volume + mask (both CUDA) -> compact level volume (uint8 when Ng <= 255,
    else int16 storage of uint16), 0 outside the mask; returns (levels, presence[Ng] int32 counts).
    Raises IndexError when a masked voxel is outside 1..Ng (reference _cmatrices.c:219)."""
    assert image.is_cuda and mask.is_cuda and image.shape == mask.shape
    image = image.to(torch.int32).contiguous()
    mask = (mask != 0).to(torch.uint8).contiguous()
    lb = lib().rb_level_bytes(int(Ng))
    lev = torch.empty(image.shape, dtype=torch.uint8 if lb == 1 else torch.int16, device=image.device)
    presence = torch.zeros(int(Ng), dtype=torch.int32, device=image.device)
    status = torch.zeros(1, dtype=torch.int32, device=image.device)
    check(lib().rb_pack_levels_dev(_ptr(image), _ptr(mask), C.c_longlong(image.numel()), int(Ng), _ptr(lev),
                                   _ptr(presence), _ptr(status), _stream()), "pack_levels")
    if int(status.item()) & 1:
        raise IndexError("gray level outside 1..Ng inside the mask")
    return lev, presence


def level_bytes(lev: torch.Tensor) -> int:
    return 1 if lev.dtype == torch.uint8 else 2


def glcm_alive_angles(lev, settings, centers=None):
    Z, Y, X = lev.shape
    alive = torch.zeros(_lib.ALIVE_WORDS, dtype=torch.int32, device=lev.device)
    check(lib().rb_glcm_alive_angles_dev(_ptr(lev), level_bytes(lev), _ptr(centers), Z, Y, X, C.byref(settings),
                                         _ptr(alive), _stream()), "glcm_alive_angles")
    return alive.cpu().numpy().view(np.uint32).copy()


def voxel_features(cls: str, lev: torch.Tensor, settings, *, centers=None, z0=0, z1=None, out=None, out_z0=None,
                   alive=None, status=None):
    """Launch the fused kernel of one class on planes [z0,z1) of `lev` (Z,Y,X).  Returns `out`:
    float64 tensor [F, z1-z0, Y, X] (allocated when None).  Asynchronous on the current stream."""
    cid = CLASS_ID[cls]
    Z, Y, X = lev.shape
    z1 = Z if z1 is None else z1
    nf = lib().rb_num_features(cid)
    if out is None:
        out = torch.empty((nf, z1 - z0, Y, X), dtype=torch.float64, device=lev.device)
        out_z0 = z0
    assert out.dtype == torch.float64 and out.is_contiguous() and out.shape[0] == nf
    if out_z0 is None:
        out_z0 = z0
    if cls == "glcm" and alive is None:
        alive = glcm_alive_angles(lev, settings, centers)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=lev.device)
    alive_p = alive.ctypes.data_as(C.c_void_p) if alive is not None else None
    check(lib().rb_voxel_features_dev(cid, _ptr(lev), level_bytes(lev), _ptr(centers), Z, Y, X, int(z0), int(z1),
                                      C.byref(settings), alive_p, _ptr(out), 0, C.c_longlong(out.stride(0)),
                                      int(out_z0), _ptr(status), _stream()), cls)
    return out


def extract_maps(image, mask, classes=CLASSES, **kw):
    """Convenience: discretised int volume + mask (numpy or CUDA tensors) -> {class: {feature: map}}
    with maps as float64 CUDA tensors (Z,Y,X).  `image` must already hold gray levels 1..Ng inside
    the mask (see imageoperations.bin_image for the discretisation kernel)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    img = torch.as_tensor(np.ascontiguousarray(image) if isinstance(image, np.ndarray) else image).to(dev)
    msk = torch.as_tensor(np.ascontiguousarray(mask) if isinstance(mask, np.ndarray) else mask).to(dev)
    if img.ndim == 2:
        img, msk = img[None], msk[None]
        if kw.get("force2D"):
            kw = dict(kw, force2Ddimension=kw.get("force2Ddimension", 0) + 1)
    Ng = int(torch.where(msk != 0, img, torch.zeros_like(img)).max().item())
    lev, presence = pack_levels(img, msk, Ng)
    n_levels = int((presence > 0).sum().item())
    settings = _lib.make_settings(Ng, n_levels, **kw)
    res = {}
    for cls in classes:
        out = voxel_features(cls, lev, settings)
        res[cls] = {name: out[i] for i, name in enumerate(_lib.feature_names(cls))}
    return res


def _runs(idx):
    """[(first feature index, count, position in idx)] for every run of consecutive indices"""
    out, k = [], 0
    while k < len(idx):
        j = k
        while j + 1 < len(idx) and idx[j + 1] == idx[j] + 1:
            j += 1
        out.append((idx[k], j - k + 1, k))
        k = j + 1
    return out


def class_maps_to_host(cls: str, lev: torch.Tensor, settings, feature_idx=None, *, centers=None, alive=None, z0=0, z1=None,
                       zchunk=64, out_dtype=torch.float64, host=None, copy_stream=None, status=None, progress=None, sync=True):
    """Output assembly of one class (the reference's per-batch `featureMaps[tuple(voxelCoords)] = ...`,
    radiomics/base.py:205-209,232-234, without the batch loop): the fused kernel runs over planes [z0,z1) in z-chunks
    into a two-slot device ring; every finished chunk leaves for the host on `copy_stream` -- ONE strided DMA per run
    of consecutive selected features (rb_memcpy2d_async) -- while the next chunk computes.  Only the maps in
    `feature_idx` (default: all) are copied.  `out_dtype` float32 converts on the device first (half the PCIe bytes).
    Returns the page-locked host tensor [len(feature_idx), z1-z0, Y, X] (allocated from torch's caching pinned
    allocator when `host` is None: the caller owns it, dropping it recycles the block).  sync=False returns without
    waiting for the last copies (the caller synchronises `copy_stream` before touching `host`), so a following class
    starts computing while this one's tail is still on the wire."""
    cid = CLASS_ID[cls]
    Z, Y, X = lev.shape
    z1 = Z if z1 is None else int(z1)
    nz = z1 - int(z0)
    nf = lib().rb_num_features(cid)
    idx = list(range(nf)) if feature_idx is None else [int(k) for k in feature_idx]
    dev = lev.device
    if host is None:
        host = torch.empty((len(idx), nz, Y, X), dtype=out_dtype, pin_memory=True)
    assert host.shape == (len(idx), nz, Y, X) and host.dtype == out_dtype and host.is_contiguous()
    if not idx or nz <= 0:
        return host
    if cls == "glcm" and alive is None:
        alive = glcm_alive_angles(lev, settings, centers)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    alive_p = alive.ctypes.data_as(C.c_void_p) if alive is not None else None
    cur = torch.cuda.current_stream(dev)
    copy_stream = copy_stream or torch.cuda.Stream(device=dev)
    zc = max(1, min(int(zchunk), nz))
    plane = Y * X
    ring = [torch.empty((nf, zc, Y, X), dtype=torch.float64, device=dev) for _ in range(2 if nz > zc else 1)]
    f32 = out_dtype == torch.float32
    ring32 = [torch.empty((len(idx), zc, Y, X), dtype=torch.float32, device=dev) for _ in ring] if f32 else None
    for t in ring + (ring32 or []):
        t.record_stream(copy_stream)                     # the caching allocator must not recycle them under the DMA
    copied = [None, None]
    L = lib()
    for i, za in enumerate(range(int(z0), z1, zc)):
        zb = min(za + zc, z1)
        slot = i % len(ring)
        if copied[slot] is not None:
            cur.wait_event(copied[slot])                 # the DMA of the chunk that used this slot has finished
        buf = ring[slot]
        check(L.rb_voxel_features_dev(cid, _ptr(lev), level_bytes(lev), _ptr(centers), Z, Y, X, int(za), int(zb),
                                      C.byref(settings), alive_p, _ptr(buf), 0, C.c_longlong(buf.stride(0)), int(za),
                                      _ptr(status), C.c_void_p(cur.cuda_stream)), cls)
        width = (zb - za) * plane
        if f32:
            for first, count, pos in _runs(idx):
                check(L.rb_maps_to_f32_dev(C.c_void_p(buf.data_ptr() + first * buf.stride(0) * 8), C.c_longlong(buf.stride(0)),
                                           C.c_void_p(ring32[slot].data_ptr() + pos * ring32[slot].stride(0) * 4),
                                           C.c_longlong(ring32[slot].stride(0)), C.c_longlong(width), C.c_longlong(count),
                                           C.c_void_p(cur.cuda_stream)), "maps_to_f32")
        done = torch.cuda.Event()
        done.record(cur)
        copy_stream.wait_event(done)
        off = (za - int(z0)) * plane
        if f32:
            src = ring32[slot]
            check(L.rb_memcpy2d_async(C.c_void_p(host.data_ptr() + off * 4), C.c_ulonglong(host.stride(0) * 4), _ptr(src),
                                      C.c_ulonglong(src.stride(0) * 4), C.c_ulonglong(width * 4), C.c_ulonglong(len(idx)), 2,
                                      C.c_void_p(copy_stream.cuda_stream)), "memcpy2d")
        else:
            for first, count, pos in _runs(idx):
                check(L.rb_memcpy2d_async(C.c_void_p(host.data_ptr() + (pos * host.stride(0) + off) * 8),
                                          C.c_ulonglong(host.stride(0) * 8),
                                          C.c_void_p(buf.data_ptr() + first * buf.stride(0) * 8),
                                          C.c_ulonglong(buf.stride(0) * 8), C.c_ulonglong(width * 8), C.c_ulonglong(count), 2,
                                          C.c_void_p(copy_stream.cuda_stream)), "memcpy2d")
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        copied[slot] = ev
        if progress is not None:
            progress(zb - za)
    if sync:
        copy_stream.synchronize()
    return host


class HostExtractor:
    """End-to-end voxel-based extraction with HOST buffers (what a pyradiomics user holds):
    int32 gray levels + mask in, float64 feature maps out, all transfers inside.  Pinned staging is
    allocated once and reused.  The 80 GB of result maps is what bounds this path (PCIe), so the
    device->host stream is kept busy from the first milliseconds: classes run in order of
    (bytes out / compute time), every class is cut into z-chunks (class_maps_to_host), and a chunk's
    maps are copied on a second stream while the next chunk / class computes.

    `shape` is the (Z,Y,X) block handed to this GPU; `z0:z1` (default everything) selects the
    planes whose maps are computed and returned -- a multi-GPU caller passes its slab plus halo
    planes read from the host volume and keeps only the interior (no collective needed)."""

    ORDER = ("gldm", "glszm", "glrlm", "ngtdm", "glcm")      # cheap-and-wide first, GLCM last

    def __init__(self, shape, classes=CLASSES, device=None, z0=0, z1=None, zchunk=64, out_dtype=torch.float64):
        self.shape = tuple(int(s) for s in shape)
        self.classes = tuple(c for c in self.ORDER if c in classes)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.z0, self.z1 = int(z0), int(self.shape[0] if z1 is None else z1)
        self.zchunk = int(zchunk)
        self.out_dtype = out_dtype
        self.out_shape = (self.z1 - self.z0,) + self.shape[1:]
        self.nf = {c: lib().rb_num_features(CLASS_ID[c]) for c in self.classes}
        n = int(np.prod(self.shape))
        self.d_img = torch.empty(self.shape, dtype=torch.int32, device=self.dev)
        self.d_msk = torch.empty(self.shape, dtype=torch.uint8, device=self.dev)
        self.h_img = torch.empty(self.shape, dtype=torch.int32, pin_memory=True)
        self.h_msk = torch.empty(self.shape, dtype=torch.uint8, pin_memory=True)
        self.h_out = {c: torch.empty((self.nf[c],) + self.out_shape, dtype=out_dtype, pin_memory=True) for c in self.classes}
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.h2d_bytes = n * 5
        self.d2h_bytes = sum(self.nf.values()) * int(np.prod(self.out_shape)) * (4 if out_dtype == torch.float32 else 8)

    def run(self, image: np.ndarray, mask: np.ndarray, Ng: int, n_roi_levels: int, alive=None, **kw):
        """returns {class: pinned tensor [F, z1-z0, Y, X]} (valid until the next run())."""
        self.h_img.numpy()[...] = image
        self.h_msk.numpy()[...] = mask
        self.d_img.copy_(self.h_img, non_blocking=True)
        self.d_msk.copy_(self.h_msk, non_blocking=True)
        lev, _ = pack_levels(self.d_img, self.d_msk, Ng)
        settings = _lib.make_settings(Ng, n_roi_levels, **kw)
        if alive is None and "glcm" in self.classes:
            from . import distributed as D
            alive = D.allreduce_alive(glcm_alive_angles(lev, settings), self.dev)   # OR over the slabs' ranks
        for c in self.classes:
            class_maps_to_host(c, lev, settings, None, alive=alive if c == "glcm" else None, z0=self.z0, z1=self.z1,
                               zchunk=self.zchunk, out_dtype=self.out_dtype, host=self.h_out[c], copy_stream=self.copy_stream,
                               sync=False)
        self.copy_stream.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
        return self.h_out


def extract_to_nrrd(lev: torch.Tensor, settings, out_dir, classes=CLASSES, prefix="original", spacing_xyz=(1.0, 1.0, 1.0),
                    origin_xyz=(0.0, 0.0, 0.0), compress=True, level=1, workers=8, out_dtype=torch.float64, centers=None,
                    zchunk=64, features=None):
    """Voxel driver + output assembly in one pipeline (the reference: extractor.execute(..., voxelBased=True) then one
    sitk.WriteImage(map, target, True) per map, radiomics/scripts/voxel.py:62-72): the fused kernels of one class stream
    their maps chunk by chunk into page-locked host memory (class_maps_to_host) while a pool of writer threads gzips the
    PREVIOUS class's maps into <prefix>_<class>_<Feature>.nrrd files (zlib releases the GIL), so compression and disk
    overlap the GPU and the PCIe stream.  `features` = {class: [names]} restricts what is copied and written.
    Returns {feature key: path}."""
    import concurrent.futures as cf
    import os

    from . import nrrd
    os.makedirs(out_dir, exist_ok=True)
    jobs, keep = {}, []
    copy_stream = torch.cuda.Stream(device=lev.device)
    with cf.ThreadPoolExecutor(max_workers=max(1, int(workers))) as ex:
        for c in [c for c in HostExtractor.ORDER if c in classes]:
            names = _lib.feature_names(c)
            want = names if not features or c not in features else [n for n in names if n in set(features[c])]
            idx = [names.index(n) for n in want]
            if not idx:
                continue
            host = class_maps_to_host(c, lev, settings, idx, centers=centers, zchunk=zchunk, out_dtype=out_dtype,
                                      copy_stream=copy_stream, sync=True)
            keep.append(host)
            arr = host.numpy()
            for pos, n in enumerate(want):
                key = f"{prefix}_{c}_{n}"
                jobs[key] = ex.submit(nrrd.write_nrrd, os.path.join(out_dir, key + ".nrrd"), arr[pos], spacing_xyz, origin_xyz,
                                      compress, level)
        return {k: j.result() for k, j in jobs.items()}
