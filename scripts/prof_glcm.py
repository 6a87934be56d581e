import sys, torch
sys.path.insert(0, ".")
from pyradiomics_b200 import _lib, voxel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
cls = sys.argv[3] if len(sys.argv) > 3 else "glcm"
torch.manual_seed(0)
dev = "cuda"
if kind == "uniform":
    lev = torch.randint(1, 33, (N, N, N), device=dev, dtype=torch.uint8)
else:
    f = torch.randn(1, 1, N, N, N, device=dev)
    k = torch.ones(1, 1, 7, 7, 7, device=dev) / 343
    f = torch.nn.functional.conv3d(f, k, padding=3)[0, 0]
    q = torch.quantile(f.flatten()[:: max(1, f.numel() // 1000000)], torch.linspace(0, 1, 33, device=dev)[1:-1])
    lev = (torch.bucketize(f, q) + 1).to(torch.uint8)
s = _lib.make_settings(32, 32)
out = voxel.voxel_features(cls, lev, s)
torch.cuda.synchronize()
out = voxel.voxel_features(cls, lev, s, out=out, out_z0=0)
torch.cuda.synchronize()
