"""developer diagnostic: determinism and fast-vs-generic agreement of the GLCM MCC map"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from pyradiomics_b200 import _lib, voxel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
for kind in (sys.argv[2:] or ["uniform", "smooth"]):
    torch.manual_seed(1)
    if kind == "uniform":
        lev = torch.randint(1, 33, (N, N, N), device="cuda", dtype=torch.uint8)
    else:
        f = torch.randn(1, 1, N, N, N, device="cuda")
        f = torch.nn.functional.conv3d(f, torch.ones(1, 1, 5, 5, 5, device="cuda") / 125, padding=2)[0, 0]
        q = torch.quantile(f.flatten(), torch.linspace(0, 1, 33, device="cuda")[1:-1])
        lev = (torch.bucketize(f, q) + 1).to(torch.uint8)
    s = _lib.make_settings(32, 32)
    k = _lib.feature_names("glcm").index("MCC")
    runs = [voxel.voxel_features("glcm", lev, s)[k].cpu().numpy() for _ in range(3)]
    os.environ["B200_RADIOMICS_FORCE_GENERIC"] = "1"
    gen = voxel.voxel_features("glcm", lev, s)[k].cpu().numpy()
    del os.environ["B200_RADIOMICS_FORCE_GENERIC"]
    for i, r in enumerate(runs):
        d = np.abs(r - gen)
        bad = np.argwhere(~(d < 1e-6))
        print(kind, "run", i, "max|fast-gen|", np.nanmax(d), "n>1e-6:", len(bad), "nan:", int(np.isnan(r).sum()),
              "differs from run0:", int((r != runs[0]).sum()), [(tuple(b), float(r[tuple(b)]), float(gen[tuple(b)])) for b in bad[:3]])
