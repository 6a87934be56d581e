"""Secondary BASELINE.json configurations on one GPU (numbers for profiles/r01_configs.md):
config 2 (GLCM-only voxel maps, 256^3), config 4 (wavelet + LoG + suite, 512^3), config 5(ii)
(segment-based full suite of independent 256^3 cases, per-GPU share of the batch)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyradiomics_b200 import _lib, cmatrices as B, pipeline as PP, voxel

dev = torch.device("cuda")
rows = []


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# config 2
for kind in ("uniform", "smooth"):
    N = 256
    rng = np.random.default_rng(0)
    if kind == "uniform":
        lev = torch.from_numpy(rng.integers(1, 33, (N, N, N)).astype(np.uint8)).to(dev)
    else:
        f = torch.randn(1, 1, N, N, N, device=dev)
        f = torch.nn.functional.conv3d(f, torch.ones(1, 1, 7, 7, 7, device=dev) / 343, padding=3)[0, 0]
        q = torch.quantile(f.flatten()[::16], torch.linspace(0, 1, 33, device=dev)[1:-1])
        lev = (torch.bucketize(f, q) + 1).to(torch.uint8)
    s = _lib.make_settings(32, 32)
    out = torch.empty((24, N, N, N), dtype=torch.float64, device=dev)
    ms = timed(lambda: voxel.voxel_features("glcm", lev, s, out=out, out_z0=0))
    rows.append(f"| 2 | voxel-based GLCM maps, {N}^3 Ng=32 r=1 ({kind}) | {ms:.1f} ms | {N ** 3 / ms * 1e3:.3e} voxels/s |")
    del out

# config 4
N = 512
torch.manual_seed(0)
f = torch.randn(1, 1, N, N, N, device=dev)
x = torch.nn.functional.conv3d(f, torch.ones(1, 1, 5, 5, 5, device=dev) / 125, padding=2)[0, 0]
del f
x = ((x - x.min()) / (x.max() - x.min()) * 800).to(torch.float32).contiguous()
m = torch.ones((N, N, N), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
imgs = list(PP.derived_images(x))
torch.cuda.synchronize()
tf = time.perf_counter() - t0
nimg = len(imgs)
del imgs
torch.cuda.empty_cache()
t0 = time.perf_counter()
info = PP.voxel_suite_with_filters(x, m, binWidth=25)
torch.cuda.synchronize()
tt = time.perf_counter() - t0
rows.append(f"| 4 | coif1 wavelet (8 bands) + LoG sigma 1,2,3 + original -> binWidth 25 -> full suite, {N}^3 | filters {tf * 1e3:.0f} ms; total {tt:.2f} s for {nimg} images | {N ** 3 * nimg / tt:.3e} voxel-suites/s |")
del x, m
torch.cuda.empty_cache()

# config 5 (ii): segment-based full suite on 256^3 cases (host-buffer cMatrices API, as a CLI worker would call it)
N = 256
rng = np.random.default_rng(1)
img = rng.integers(1, 33, (N, N, N)).astype(np.int32)
msk = np.ones(img.shape, bool)
d1 = np.array([1])


def seg_case():
    B.calculate_glcm(img, msk, d1, 32, False, 0)
    B.calculate_glrlm(img, msk, 32, N, False, 0)
    B.calculate_glszm(img, msk, 32, N ** 3, False, 0)
    B.calculate_gldm(img, msk, d1, 32, 0, False, 0)
    B.calculate_ngtdm(img, msk, d1, 32, False, 0)


seg_case()
t0 = time.perf_counter()
for _ in range(3):
    seg_case()
tc = (time.perf_counter() - t0) / 3
rows.append(f"| 5(ii) | segment-based five matrices of one {N}^3 case (host buffers in, matrices out) | {tc * 1e3:.0f} ms / case | {N ** 3 / tc:.3e} voxels/s per GPU; a batch of 64 cases on 8 GPUs = 8 cases each, no collective |")
print("| config | workload | time | throughput |\n|---|---|---|---|")
print("\n".join(rows))
