#!/usr/bin/env python
"""bench.py -- waterfall Gpixel/s of create_dataset + evaluate_segmentation (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--baselines B]

One "step" = one pass of the hot path over one synthetic cube:
    Preprocessor(cube_c64, magnitude=True).create_dataset(patch_size=128, stretch="SQRT",
        flag_sigma=5, use_custom_flags=False)  +  evaluate_segmentation(labels, truth)
on BASELINE.json configs[1] (45 baselines x 4 pols x 1024 ch x 1024 times, complex64,
SQRT stretch, 4-way augmentation).  N > 1 (torchrun, one rank per GPU): every rank owns its
own 45-baseline shard (baselines shard with no data-path exchange; weak scaling) and the
{TP, FP, FN} counts are all-reduced over NCCL.

JSON keys follow the driver contract; see DESIGN.md "Measurement" for the definitions of
`value` (inputs resident in HBM), `e2e` (host buffers, copies inside the timed region),
`roofline` (write_patches kernel, algorithmic bytes / CUDA-event time / measured HBM peak)
and `cpu_baseline` (the NumPy oracle port on the host cores, bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(n_bl=45, n_pol=4, channels=1024, times=1024, patch=128, stretch="SQRT", sigma=5, rot=4)
# other BASELINE.json configs, runnable for the record (`--workload c3`): per-GPU shard of the
# VLA-scale cube (351 baselines over 8 GPUs = 44 per GPU), LOG10 stretch
WORKLOADS = {
    "c2": dict(WORKLOAD, name="configs[1]"),
    "c3": dict(n_bl=44, n_pol=4, channels=4096, times=2048, patch=128, stretch="LOG10", sigma=5, rot=4, name="configs[2] (one of 8 baseline shards)"),
    # long-track cube, P = 256 (big-tile path): the per-GPU shard (44 baselines, 153 GB of output) is
    # streamed through create_dataset in baseline chunks; one step = one 8-baseline chunk
    "c5": dict(n_bl=8, n_pol=4, channels=1024, times=16384, patch=256, stretch=None, sigma=3, rot=4,
               name="configs[4] (8-baseline chunk of one of 8 baseline shards)"),
}
METRIC = "waterfall Gpixel/s (create_dataset+metrics)"
_JSON_FD = None  # the real stdout; fd 1 itself is pointed at stderr while the benchmark runs


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its
    version banner / NCCL_DEBUG output to stdout): keep a private handle on the real stdout and
    send everything else that is written to fd 1 to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())
ACTIVE = WORKLOAD  # set from --workload in main(); inherited by the forked CPU-baseline workers


# ------------------------------------------------------------------------------------- CPU side
def _cpu_cube(n_bl, seed):
    from tests.cubes import make_cube
    return make_cube(n_bl=n_bl, n_pol=ACTIVE["n_pol"], channels=ACTIVE["channels"],
                     times=ACTIVE["times"], seed=seed, dtype=np.complex64)


def _cpu_step(args):
    """Reference path on one baseline slice: np.abs -> create_dataset -> evaluate_segmentation."""
    import oracle
    cube, truth_seed = args
    np.random.seed(truth_seed)
    ds = oracle.create_dataset(np.abs(cube), None, patch_size=ACTIVE["patch"], stretch=ACTIVE["stretch"],
                               flag_sigma=ACTIVE["sigma"], use_custom_flags=False, num_workers=0)
    truth = ds.labels ^ (np.random.default_rng(truth_seed).random(ds.labels.shape) < 0.01)
    oracle.evaluate_segmentation(ds.labels, truth)
    return cube.size


def cpu_baseline(n_bl=1, procs=1):
    """Times the oracle port on `procs` processes, each over `n_bl` baselines. -> Gpixel/s."""
    cubes = [_cpu_cube(n_bl, 100 + i)[0] for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        npix = _cpu_step((cubes[0], 1))
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            npix = sum(pool.map(_cpu_step, [(c, i) for i, c in enumerate(cubes)]))
    dt = time.perf_counter() - t0
    return npix / dt / 1e9, npix, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = os.cpu_count() or 1
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(1, procs)
    times, npix = [], 0
    for _ in range(args.steps):
        v, npix, dt = cpu_baseline(1, procs)
        times.append(dt)
    total = sum(times)
    value = npix * len(times) / total / 1e9
    sample = (f"{procs} processes x 1 baseline x 4 pols x {ACTIVE['channels']}x{ACTIVE['times']} per step "
              "(one Preprocessor per process)")
    line = {
        "metric": METRIC, "value": value, "unit": "Gpixel/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ACTIVE.get('name', 'configs[1]')} {ACTIVE['n_bl']}bl x {ACTIVE['n_pol']}pol x "
                               f"{ACTIVE['channels']}ch x {ACTIVE['times']}t complex64, {ACTIVE['stretch']}, "
                               f"MAD sigma={ACTIVE['sigma']}, R={ACTIVE['rot']}, P={ACTIVE['patch']}",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML every 5 ms from a host thread
    (nvidia-smi's own loop cannot go below ~50 ms, longer than a short timed region);
    `window()` keeps the samples whose host timestamp falls inside the timed region."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._thr, self._h = index, [], False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            return
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _loop(self):
        nv = self._nv
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.rows.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop = True
        if self._thr:
            self._thr.join(timeout=1.0)

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        scope = "timed region"
        if not rows:
            rows = [r for r in self.rows if r[2] > 250.0]
            scope = "under load (warm-up + timed region; region shorter than the sampling period)"
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({n for r in rows for bit, n in names.items() if r[3] & bit})
        sm = [r[1] for r in rows]
        pw = [r[2] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------- GPU side
def run_ours(args):
    import torch
    import torch.distributed as dist

    from rfi_toolbox_b200 import Preprocessor, evaluate_segmentation
    from rfi_toolbox_b200.utils.synth import device_cube

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from rfi_toolbox_b200.utils.device import bind_host_to_device
    cpus = bind_host_to_device(dev) if world > 1 else None  # pinned host buffers on the GPU's own socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = True if world > 1 else None

    w = dict(WORKLOADS[args.workload])
    if args.baselines:
        w["n_bl"] = args.baselines
    # Philox seed 1234, one subsequence per baseline of the sharded cube
    cube, mask = device_cube(w["n_bl"], w["n_pol"], w["channels"], w["times"], seed=1234, device=dev,
                             first_baseline=rank * w["n_bl"])
    npix = cube.numel()
    kw = dict(patch_size=w["patch"], stretch=w["stretch"], flag_sigma=w["sigma"], use_custom_flags=False,
              augmentation_rotations=w["rot"])

    def step(data, seed, profile=False):
        np.random.seed(seed)
        pre = Preprocessor(data, None, magnitude=True, pin=True)
        pre.profile = profile
        ds = pre.create_dataset(**kw)
        return pre, ds

    # ground truth for the metric: the labels of a first pass with 1 % of pixels toggled
    pre, ds = step(cube, 0)
    truth = ds.labels ^ (torch.rand(ds.labels.shape, device=dev) < 0.01).to(torch.uint8)
    n_kept = len(ds)
    del ds, pre

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input arm
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        pre, ds = step(cube, 0)
        evaluate_segmentation(ds.labels, truth, group=group)
        del ds, pre
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = {"stats": [], "write": []}
    t0 = time.perf_counter()
    e0.record()
    evs = []
    for i in range(args.steps):
        pre, ds = step(cube, 0, profile=True)
        m = evaluate_segmentation(ds.labels, truth, group=group)
        evs.append(pre.events)
        del ds, pre  # one dataset resident at a time (configs[2]: 76 GB of patches per step)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    wall = t1 - t0
    sampler.stop()
    clocks = sampler.window(t0, t1)
    dev_ms = e0.elapsed_time(e1)
    for ev in evs:
        for k in kern_ms:
            kern_ms[k].append(ev[k][0].elapsed_time(ev[k][1]))
    elapsed = max(dev_ms / 1e3, 0.0)
    if world > 1:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = world * npix * args.steps / elapsed / 1e9

    # ---- end-to-end arm: host (pinned) cube -> H2D -> create_dataset -> metrics -> D2H of the metric
    host = torch.empty(cube.shape, dtype=cube.dtype, pin_memory=True)
    host.copy_(cube)
    for i in range(max(1, min(args.warmup, 2))):
        pre, ds = step(host, 0)
        evaluate_segmentation(ds.labels, truth, group=group)
        del ds, pre
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    e0.record()
    for i in range(e2e_steps):
        pre, ds = step(host, 0)
        m = evaluate_segmentation(ds.labels, truth, group=group)
        del ds, pre
    e1.record()
    barrier()
    e2e_elapsed = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([e2e_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = world * npix * e2e_steps / e2e_elapsed / 1e9
    n_tiles = npix // (w["patch"] ** 2)
    h2d = cube.numel() * cube.element_size() + n_tiles * w["rot"] * 8
    d2h = n_tiles * 4 + 24

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    k = n_kept / (n_tiles * w["rot"])
    # complex64 through the real branch: both writers read the exact float32 magnitudes phase 1 left in
    # its scratch (4 B / px), not the complex cube (8 B / px); the cube itself is read once, by phase 1
    alg_bytes = npix * (4 + w["rot"] * k * 13.0)
    write_ms = float(np.mean(kern_ms["write"]))
    stats_ms = float(np.mean(kern_ms["stats"]))
    achieved = alg_bytes / (write_ms * 1e-3) / 1e9
    # DRAM bytes of one launch from the committed `ncu --set full` capture of this very workload
    # (profiles/traffic.json, written by scripts/ncu_summary.py); null for any other workload
    traffic = None
    tj = ROOT / "profiles" / "traffic.json"
    if tj.exists() and not args.baselines and args.workload in ("c2", "c5"):
        kname = "write_patches_kernel" if args.workload == "c2" else "big_write_kernel"
        traffic = json.loads(tj.read_text()).get(kname, {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "write_patches_kernel" if w["patch"] == 128 else "big_write_kernel",
                "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": write_ms,
                "stats_kernel_ms": stats_ms, "stats_kernel_gbs": npix * (cube.element_size() + 4) / (stats_ms * 1e-3) / 1e9,
                "bytes_per_pixel": {"phase1": cube.element_size() + 4, "writer": 4 + w["rot"] * k * 13.0,
                                    "path_minimum": cube.element_size() + w["rot"] * k * 13.0},
                "step_ms_device": dev_ms / args.steps}

    cpu_bl = 2 if w["channels"] * w["times"] <= 1 << 20 else 1
    cpu_v, cpu_npix, cpu_dt = cpu_baseline(cpu_bl, 1)
    line = {
        "metric": METRIC, "value": value, "unit": "Gpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w['name']} {w['n_bl']}bl x {w['n_pol']}pol x {w['channels']}ch x {w['times']}t "
                               f"complex64 per GPU, magnitude fused, {w['stretch']}, MAD sigma={w['sigma']}, "
                               f"R={w['rot']}, P={w['patch']}",
                   "pixels_per_step_per_gpu": npix, "patches_kept": n_kept,
                   "l2": f"input cube {cube.numel() * 8 / 1e9:.1f} GB and {n_kept * w['patch'] ** 2 * 13 / 1e9:.1f} GB of output "
                         "per step exceed the 126 MB L2; no flush needed",
                   "parallelism": (f"baseline-sharded x{world}; TP/FP/FN summed inside the counting kernel over NVLink peer "
                                   "memory (NCCL only for rendezvous / barriers)") if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "Gpixel/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "note": "pinned host cube copied H2D every step; dataset stays in HBM, metric dict read back"},
        # per step, P = 128: tile_stats_mono_kernel, write_patches_kernel, confusion_kernel;
        # P = 256: big_init / load / sample / pass<0> / median / pass<1> / mad / count / write + confusion_kernel
        "gpu_launches": (3 if w["patch"] == 128 else 10) * args.steps,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_v, "unit": "Gpixel/s", "cores": 1, "kind": "port",
                         "sample": f"{cpu_bl} baselines x 4 pols x {w['channels']}x{w['times']} ({cpu_npix} px) in {cpu_dt:.1f} s, one process"},
        "host_affinity": (f"{len(cpus)} cores near GPU {local}" if cpus else None),
        "clocks": clocks, "wall_s": wall, "metrics_last_step": {k_: float(v) for k_, v in m.items()},
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--baselines", type=int, default=0, help="override the 45 baselines per GPU (debug)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = the bench workload (default)")
    args = ap.parse_args()
    global ACTIVE
    ACTIVE = WORKLOADS[args.workload]
    _reserve_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
