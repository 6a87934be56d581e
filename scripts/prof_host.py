"""Developer probe: where does the HOST time of the plugin calls go?  cProfile of (a) the segment-based suite on a few
256^3 cases and (b) one voxel-based extraction, both through the feature classes with host buffers.
    python scripts/prof_host.py [n_segment=256] [n_voxel=256]"""
import cProfile
import io
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyradiomics_b200 import featureclasses as FC

CLASSES = ("glcm", "glrlm", "glszm", "gldm", "ngtdm")
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)


def top(pr, n=28):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(n)
    return "\n".join(l[:150] for l in s.getvalue().splitlines() if l.strip())


raws = [((rng.integers(1, 33, (ns, ns, ns)).astype(np.int16) - 1) * 25 + 3) for _ in range(4)]
mask = np.ones((ns, ns, ns), np.uint8)


def seg(raw, **kw):
    FC.clear_device_cache()
    return {c: FC.FEATURE_CLASSES[c](raw, mask, binWidth=25, **kw).execute() for c in CLASSES}


seg(raws[0])
torch.cuda.synchronize()
for kw in ({}, {"b200_image_key": "k"}):
    t0 = time.perf_counter()
    for i, r in enumerate(raws[1:]):
        if kw:
            kw["b200_image_key"] = i
        seg(r, **kw)
    torch.cuda.synchronize()
    print(f"segment suite {ns}^3 {kw and 'with image key' or 'content fingerprint'}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per case")
pr = cProfile.Profile()
pr.enable()
for r in raws[1:]:
    seg(r)
torch.cuda.synchronize()
pr.disable()
print(top(pr))

raw = (rng.integers(1, 33, (nv, nv, nv)).astype(np.int16) - 1) * 25 + 3
maskv = np.ones((nv, nv, nv), np.uint8)


def vox(**kw):
    FC.clear_device_cache()
    return {c: FC.FEATURE_CLASSES[c](raw, maskv, voxelBased=True, binWidth=25, **kw).execute() for c in CLASSES}


vox(); vox()
for kw in ({}, {"b200_image_key": "v"}):
    t0 = time.perf_counter()
    vox(**kw)
    print(f"voxel suite {nv}^3 {kw and 'with image key' or 'content fingerprint'}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
vox(b200_image_key="w")
pr.disable()
print(top(pr))
