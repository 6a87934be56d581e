#!/usr/bin/env python
"""bench.py -- voxels/s of full-texture-suite voxel-based extraction (BASELINE.json metric).

A "step" is one pass of the hot path over the whole synthetic volume: GLCM + GLRLM + GLSZM + GLDM
+ NGTDM fused kernels -> 75 float64 feature maps (reference dtypes), kernelRadius 1, Ng 32.
  value  : device-resident (levels in HBM -> maps in HBM), CUDA events, max over ranks
  e2e    : the same through the host-buffer API (pinned int32 image + mask in, 75 float64 maps out,
           H2D/D2H inside the timed region)
  N > 1  : the volume is split into z-slabs, one per GPU ("strong" scaling); every step exchanges
           the halo planes with NCCL send/recv and all-reduces the GLCM alive-angle mask
  --impl reference : the reference's own CPU path (compiled `_cmatrices` from oracle/_ref for the
           matrices + the numpy feature port) on the host cores, bounded voxel sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

ALG_BYTES = {"glcm": 5 + 8 * 24, "glrlm": 5 + 8 * 16, "glszm": 5 + 8 * 16, "gldm": 5 + 8 * 14, "ngtdm": 5 + 8 * 5}
CLASSES = ("glcm", "glrlm", "glszm", "gldm", "ngtdm")
METRIC = "voxels/s full-texture voxel-based"


def synth_volume(n, kind):
    """iid uniform levels 1..32 (seed 0), or the smooth variant."""
    rng = np.random.default_rng(0)
    if kind == "uniform":
        return rng.integers(1, 33, (n, n, n), dtype=np.int32)
    import scipy.ndimage as ndi
    f = ndi.gaussian_filter(rng.standard_normal((n, n, n), dtype=np.float32), 3.0)
    q = np.quantile(f.ravel()[:: max(1, f.size // 2000000)], np.linspace(0, 1, 33)[1:-1])
    return (np.digitize(f, q) + 1).astype(np.int32)


# ----------------------------------------------------------------------------------- CPU arm
_CPU = {}


def _cpu_worker(args):
    """full suite (matrices + features) for a list of centre voxels, reference-style dense path"""
    import features_np as F
    vox, Ng, levels = args
    img, msk, cm = _CPU["img"], _CPU["msk"], _CPU["cm"]
    d1 = np.array([1])
    t0 = time.perf_counter()
    P, _ = cm.calculate_glcm(img, msk, d1, Ng, False, 0, 1, vox)
    F.glcm_features(F.glcm_matrix(P, levels), levels, Ng)
    del P
    P, _ = cm.calculate_glrlm(img, msk, Ng, int(max(img.shape)), False, 0, 1, vox)
    F.glrlm_features(P, levels)
    del P
    F.glszm_features(cm.calculate_glszm(img, msk, Ng, int(msk.size), False, 0, 1, vox), levels)
    F.gldm_features(cm.calculate_gldm(img, msk, d1, Ng, 0, False, 0, 1, vox), levels)
    F.ngtdm_features(cm.calculate_ngtdm(img, msk, d1, Ng, False, 0, 1, vox))
    return time.perf_counter() - t0


def cpu_arm_setup(vol):
    import build_ref
    try:
        cm, kind = build_ref.load(), "reference"
    except ImportError:
        import cmatrices_oracle as cm
        kind = "port"
    _CPU.update(img=vol, msk=np.ones(vol.shape, bool), cm=cm)
    return kind


def cpu_arm_step(vol, workers, per_worker, seed, pool):
    rng = np.random.default_rng(seed)
    n = vol.shape[0]
    vox = rng.integers(0, n, (3, workers * per_worker)).astype(np.int32)
    levels = np.arange(1, 33)
    jobs = [(np.ascontiguousarray(vox[:, i * per_worker:(i + 1) * per_worker]), 32, levels) for i in range(workers)]
    t0 = time.perf_counter()
    if pool is None:
        for j in jobs:
            _cpu_worker(j)
    else:
        pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return workers * per_worker / dt, dt


def run_reference(args):
    """--impl reference: rank 0 only; other ranks exit quietly."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    vol = synth_volume(args.size, args.kind)
    kind = cpu_arm_setup(vol)
    cores = min(os.cpu_count() or 1, args.cpu_workers)
    per_worker = args.cpu_voxels_per_worker
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    for w in range(args.warmup):
        cpu_arm_step(vol, cores, max(8, per_worker // 8), 100 + w, pool)
    t0 = time.perf_counter()
    nvox = 0
    for k in range(args.steps):
        v, dt = cpu_arm_step(vol, cores, per_worker, k, pool)
        nvox += cores * per_worker
    total = time.perf_counter() - t0
    if pool:
        pool.close()
    value = nvox / total
    sample = f"{cores * per_worker} random centre voxels of the {args.size}^3 volume per step, batches of {per_worker}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"full suite GLCM+GLRLM+GLSZM+GLDM+NGTDM voxel-based, synthetic {args.size}^3 Ng=32 "
                        f"kernelRadius=1 ({args.kind} levels), 75 float64 maps",
            "size": args.size, "Ng": 32, "kernelRadius": 1, "levels": args.kind,
            "l2": "inputs+outputs (>=25 GB per class) far exceed the 126 MB L2; no flush needed",
            "parallelism": f"z-slabs x{args.gpus}" if args.gpus > 1 else "single GPU"}


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        mx = max((float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    vol = synth_volume(args.size, args.kind)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (fork-safe)
        import multiprocessing as mp
        kind = cpu_arm_setup(vol)
        cores = min(os.cpu_count() or 1, args.cpu_workers)
        pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
        cpu_arm_step(vol, cores, 8, 99, pool)
        v, dt = cpu_arm_step(vol, cores, args.cpu_voxels_per_worker, 0, pool)
        if pool:
            pool.close()
        cpu_baseline = {"value": v, "unit": "voxels/s", "cores": cores, "kind": kind,
                        "sample": f"{cores * args.cpu_voxels_per_worker} random centre voxels of the same {args.size}^3 "
                                  f"volume, full suite matrices+features, {dt:.1f} s wall"}

    from pyradiomics_b200 import _lib, distributed as D, voxel

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Z = args.size
    z0, z1 = D.slab_range(Z, rank, world)
    r = 1
    settings = _lib.make_settings(32, 32)
    own = torch.from_numpy(vol[z0:z1].astype(np.uint8)).to(dev)
    slab = D.SlabHalo(own, r, rank, world)
    nz = z1 - z0
    outs = {c: torch.empty((_lib.lib().rb_num_features(_lib.CLASS_ID[c]), nz, Z, Z), dtype=torch.float64, device=dev)
            for c in CLASSES}
    ev = {c: [] for c in CLASSES}
    launches = 0
    zchunk = max(1, (48 << 20) // (Z * Z * 13))
    glcm_chunks = -(-nz // zchunk)

    def step(record):
        nonlocal launches
        slab.exchange()
        buf = slab.buf
        alive = voxel.glcm_alive_angles(buf, settings)
        alive = D.allreduce_alive(alive, dev)
        launches += 1                                  # glcm_alive_kernel
        for c in CLASSES:
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            voxel.voxel_features(c, buf, settings, z0=r, z1=r + nz, out=outs[c], out_z0=r, alive=alive)
            # GLCM = (features, eigen-solve, finish) kernels per plane chunk of the task queue
            launches += 5 * glcm_chunks if c == "glcm" else 1   # per plane chunk: phase A, 3 eigen-solve kernels, finish
            if record:
                e1.record()
                ev[c].append((e0, e1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = 0
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    for _ in range(args.steps):
        step(True)
    s1.record()
    barrier()
    ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    per_class_ms = {c: float(np.mean([a.elapsed_time(b) for a, b in ev[c]])) for c in CLASSES}
    nvox_total = Z ** 3
    value = nvox_total * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel (longest class kernel), algorithmic bytes / event time
    dom = max(per_class_ms, key=per_class_ms.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = ALG_BYTES[dom] * (nz * Z * Z) / (per_class_ms[dom] * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if prof.get("kernel_class") == dom and prof.get("voxels"):
            traffic = prof["dram_bytes_per_launch"] * (nz * Z * Z) / prof["voxels"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": f"{dom} fused voxel kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "ms_per_launch": per_class_ms[dom], "per_class_ms": per_class_ms,
                "suite_frac": 625.0 * value / world / 1e9 / peak,
                "note": "compute-bound fp64/integer kernel: see DESIGN.md section 'roofline'"}

    # ---- e2e through the host-buffer API: every rank takes its slab (+ halo planes) of the HOST
    # volume, H2D, discretised-level packing, the five fused kernels, D2H of its slab's 75 maps
    e2e = None
    if not args.no_e2e:
        del outs
        torch.cuda.empty_cache()
        h0, h1 = max(z0 - r, 0), min(z1 + r, Z)
        hx = voxel.HostExtractor((h1 - h0, Z, Z), CLASSES, dev, z0=z0 - h0, z1=z1 - h0)
        blk = np.ascontiguousarray(vol[h0:h1])
        msk = np.ones(blk.shape, np.uint8)
        hx.run(blk, msk, 32, 32)                             # warm-up (also faults the pinned pages in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hx.run(blk, msk, 32, 32)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        hb = torch.tensor([hx.h2d_bytes, hx.d2h_bytes], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
        dtv = float(dt.item())
        e2e = {"value": nvox_total / dtv, "unit": "voxels/s", "h2d_bytes_per_step": int(hb[0].item()),
               "d2h_bytes_per_step": int(hb[1].item()), "ms_per_step": dtv * 1e3, "steps": args.e2e_steps,
               "api": "pyradiomics_b200.voxel.HostExtractor.run (pinned host buffers per rank, D2H overlapped per class, max over ranks)"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--kind", default="uniform", choices=["uniform", "smooth"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-workers", type=int, default=64)
    ap.add_argument("--cpu-voxels-per-worker", type=int, default=96)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
