"""BASELINE.json config 4 on one GPU: wavelet (8 sub-bands) + LoG (sigma 1,2,3) + original, each
discretised (binWidth 25) and run through the full voxel-based suite; prints timings per stage."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyradiomics_b200 import imageoperations as IO, pipeline as PP

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
torch.manual_seed(0)
f = torch.randn(1, 1, N, N, N, device=dev)
k = torch.ones(1, 1, 5, 5, 5, device=dev) / 125
x = torch.nn.functional.conv3d(f, k, padding=2)[0, 0]
x = ((x - x.min()) / (x.max() - x.min()) * 800).to(torch.float32).contiguous()
m = torch.ones((N, N, N), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
imgs = list(PP.derived_images(x))
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"filters: {len(imgs)} derived images of {N}^3 in {(t1 - t0) * 1e3:.1f} ms "
      f"({N ** 3 * (len(imgs) - 1) / (t1 - t0):.3e} output voxels/s)")
del imgs
t0 = time.perf_counter()
info = PP.voxel_suite_with_filters(x, m, binWidth=25)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"config4 end to end: {len(info)} images x 75 maps in {(t1 - t0):.2f} s -> {N ** 3 * len(info) / (t1 - t0):.3e} voxel-suites/s")
for n, Ng, nl in info:
    print(f"  {n:28s} Ng={Ng:4d} levels={nl}")
