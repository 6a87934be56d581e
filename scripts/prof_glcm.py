import sys, torch
sys.path.insert(0, ".")
from pyradiomics_b200 import _lib, voxel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
lev = torch.randint(1, 33, (N, N, N), device="cuda", dtype=torch.uint8)
s = _lib.make_settings(32, 32)
out = voxel.voxel_features("glcm", lev, s)
torch.cuda.synchronize()
out = voxel.voxel_features("glcm", lev, s, out=out, out_z0=0)
torch.cuda.synchronize()
