"""developer probe / ncu target: the five segment-mode matrix builders (cMatrices drop-in, host buffers) on one synthetic
N^3 case; prints host wall time per class.  python scripts/prof_segment.py N kind [repeat]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from pyradiomics_b200 import cmatrices
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lev = bench.synth_volume(N, kind) if kind == "uniform" else None
if lev is None:
    rng = np.random.default_rng(0)
    import scipy.ndimage as ndi
    f = ndi.gaussian_filter(rng.standard_normal((N, N, N), dtype=np.float32), 3.0)
    lev = (np.digitize(f, np.quantile(f.ravel(), np.linspace(0, 1, 33)[1:-1])) + 1).astype(np.int32)
m = np.ones(lev.shape, bool)
for r in range(rep):
    t = {}
    t0 = time.perf_counter(); cmatrices.calculate_glcm(lev, m, [1], 32, False, -1); t["glcm"] = time.perf_counter() - t0
    t0 = time.perf_counter(); cmatrices.calculate_glrlm(lev, m, 32, N, False, -1); t["glrlm"] = time.perf_counter() - t0
    t0 = time.perf_counter(); cmatrices.calculate_glszm(lev, m, 32, int(m.sum()), False, -1); t["glszm"] = time.perf_counter() - t0
    t0 = time.perf_counter(); cmatrices.calculate_gldm(lev, m, [1], 32, 0, False, -1); t["gldm"] = time.perf_counter() - t0
    t0 = time.perf_counter(); cmatrices.calculate_ngtdm(lev, m, [1], 32, False, -1); t["ngtdm"] = time.perf_counter() - t0
    print(kind, N, "host wall ms:", {k: round(v * 1e3, 1) for k, v in t.items()}, flush=True)
