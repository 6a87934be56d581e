"""Developer timing probe: per-class throughput of the fused voxel kernels on a synthetic volume."""
import sys
import time

import torch

sys.path.insert(0, ".")
from pyradiomics_b200 import _lib, voxel

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
dev = torch.device("cuda")
torch.manual_seed(0)
if kind == "uniform":
    lev = torch.randint(1, 33, (N, N, N), device=dev, dtype=torch.uint8)
else:
    f = torch.randn(1, 1, N, N, N, device=dev)
    k = torch.ones(1, 1, 7, 7, 7, device=dev) / 343
    f = torch.nn.functional.conv3d(f, k, padding=3)[0, 0]
    q = torch.quantile(f.flatten()[:: max(1, f.numel() // 1000000)], torch.linspace(0, 1, 33, device=dev)[1:-1])
    lev = (torch.bucketize(f, q) + 1).to(torch.uint8)
s = _lib.make_settings(32, 32)
tot = 0
for cname in _lib.CLASSES:
    out = voxel.voxel_features(cname, lev, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    voxel.voxel_features(cname, lev, s, out=out, out_z0=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot += ms
    print(f"{cname:6s} N={N} {kind}: {ms:9.2f} ms  {N**3/ms*1e3:.3e} vox/s", flush=True)
    del out
print(f"suite  N={N} {kind}: {tot:9.2f} ms  {N**3/tot*1e3:.3e} vox/s")
