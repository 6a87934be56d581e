"""Host-side placement for the output stream of voxel-based extraction: the 75 float64 maps of a 512^3 case are 80 GB of
device->host traffic, and on a two-socket host a rank whose page-locked buffers live on the far NUMA node pushes every
byte through the inter-socket link.  Binding a rank to the CPUs next to its GPU BEFORE it allocates makes the pinned
pages local (first touch).  (The reference's own parallel model -- one worker process per case,
radiomics/scripts/__init__.py:393-404 -- leaves placement to the OS.)"""
from __future__ import annotations

import os
import re
import subprocess


def gpu_cpu_affinity(index: int):
    """CPU ids close to GPU `index` (NVML's ideal affinity; falls back to the `nvidia-smi topo -m` table); [] if unknown"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if int(word) >> b & 1]
        if cpus:
            return cpus
    except Exception:
        pass
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        for line in txt.splitlines():
            if re.match(rf"\x1b?\[?[0-9;]*m?GPU{int(index)}\s", line) or line.startswith(f"GPU{int(index)}\t"):
                m = re.search(r"\t(\d+(?:-\d+)?(?:,\d+(?:-\d+)?)*)\t", line + "\t")
                if m:
                    cpus = []
                    for part in m.group(1).split(","):
                        a, _, b = part.partition("-")
                        cpus += list(range(int(a), int(b or a) + 1))
                    return cpus
    except Exception:
        pass
    return []


def bind_to_gpu(index: int):
    """restrict this process to the CPUs next to GPU `index`; returns the CPU list used ([] = left unbound)"""
    cpus = gpu_cpu_affinity(index)
    if not cpus:
        return []
    try:
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & set(allowed)) or sorted(allowed)
        os.sched_setaffinity(0, use)
        return use
    except (AttributeError, OSError):
        return []
