#!/usr/bin/env python
"""bench.py -- waterfall Gpixel/s of create_dataset + evaluate_segmentation (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c5] [--baselines B] [--lookahead L] [--no-extra]

One "step" = one pass of the hot path over one synthetic cube:
    Preprocessor(cube_c64, magnitude=True).create_dataset(patch_size=128, stretch="SQRT",
        flag_sigma=5, use_custom_flags=False)  +  evaluate_segmentation(labels, truth)
on BASELINE.json configs[1] (45 baselines x 4 pols x 1024 ch x 1024 times, complex64,
SQRT stretch, 4-way augmentation).  N > 1 (torchrun, one rank per GPU): every rank owns its
own 45-baseline shard (baselines shard with no data-path exchange; weak scaling) and the
{TP, FP, FN} counts are summed over the ranks inside the counting kernel (NVLink peer memory).

The K timed steps are issued the way a streaming caller issues them (one Preprocessor per
sample / baseline chunk, the reference's own unit of work, synthetic_generator.py:55-107):
`create_dataset_async` keeps `--lookahead` calls in flight, so the host phase of step k (flag
counts D2H -> np.random.permutation -> destination slots H2D) hides behind phase 1 of step
k + 1.  Exactly K complete steps -- pipeline fill and drain included -- lie inside the timed
region; `--lookahead 0` is the plain sequential loop.

JSON keys follow the driver contract; see DESIGN.md "Measurement":
  value     inputs resident in HBM
  e2e       pinned HOST cube copied H2D every step, metric dict read back (dataset stays in HBM)
  e2e_host_result  as e2e, plus the whole dataset (images + labels) downloaded to pinned host
            memory every step -- what the reference's callers hold after `create_dataset`
  roofline  write_patches kernel (algorithmic bytes / CUDA-event time / measured HBM peak) and the
            PATH-level fraction: (s_in + R k 13 [+ 2 B per label px]) * px / step time
  extra     the other BASELINE configs measured in the same run: c3_shard, c5_chunk, c4_sweep
  cpu_baseline  the reference's CPU path on one host core, bounded sample
`--impl reference` times the UNMODIFIED reference package (baseline/_ref, installed by
`__graft_entry__.build()`; the NumPy oracle port only if that is absent) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from collections import deque
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(n_bl=45, n_pol=4, channels=1024, times=1024, patch=128, stretch="SQRT", sigma=5, rot=4)
WORKLOADS = {
    "c2": dict(WORKLOAD, name="configs[1]"),
    # per-GPU shard of the VLA-scale cube (351 baselines over 8 GPUs = 44 per GPU), LOG10 stretch
    "c3": dict(n_bl=44, n_pol=4, channels=4096, times=2048, patch=128, stretch="LOG10", sigma=5, rot=4,
               name="configs[2] (one of 8 baseline shards)"),
    # long-track cube, P = 256 (big-tile path): the per-GPU shard (44 baselines, 153 GB of output) is
    # streamed through create_dataset in baseline chunks; one step = one 8-baseline chunk
    "c5": dict(n_bl=8, n_pol=4, channels=1024, times=16384, patch=256, stretch=None, sigma=3, rot=4,
               name="configs[4] (8-baseline chunk of one of 8 baseline shards)"),
}
METRIC = "waterfall Gpixel/s (create_dataset+metrics)"
_JSON_FD = None  # the real stdout; fd 1 itself is pointed at stderr while the benchmark runs
METRICS_SIDE = True  # --metrics-stream main switches the counting kernel back to the caller's stream (A/B)
ACTIVE = WORKLOAD  # set from --workload in main(); inherited by the forked CPU-baseline workers


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its
    version banner / NCCL_DEBUG output to stdout, the reference package prints debug lines on
    import): keep a private handle on the real stdout and send everything else written to fd 1 to
    stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def _workload_text(w, per_gpu=True):
    return (f"{w.get('name', 'configs[1]')} {w['n_bl']}bl x {w['n_pol']}pol x {w['channels']}ch x {w['times']}t "
            f"complex64{' per GPU' if per_gpu else ''}, magnitude fused, {w['stretch']}, MAD sigma={w['sigma']}, "
            f"R={w['rot']}, P={w['patch']}")


# ------------------------------------------------------------------------------------- CPU side
def _reference_api():
    """(create_dataset(data, **kw) -> dataset with .labels, evaluate_segmentation, kind).
    The unmodified reference package when baseline/_ref holds it, else the NumPy oracle port."""
    ref = ROOT / "baseline" / "_ref"
    if (ref / "rfi_toolbox" / "__init__.py").exists():
        if str(ref) not in sys.path:
            sys.path.append(str(ref))  # at the END: it also ships a `tests` package
        os.environ.setdefault("CI", "1")  # the reference then skips its own process pools (preprocessor.py:491)
        from rfi_toolbox.evaluation import evaluate_segmentation as ref_eval
        from rfi_toolbox.preprocessing import Preprocessor as RefPre

        def create(data, **kw):
            return RefPre(data, None).create_dataset(num_workers=0, **kw)

        return create, (lambda ds: ds.labels.numpy()), ref_eval, "reference"
    import oracle

    def create(data, **kw):
        return oracle.create_dataset(data, None, num_workers=0, **kw)

    return create, (lambda ds: ds.labels), oracle.evaluate_segmentation, "port"


def _cpu_cube(n_bl, seed):
    from tests.cubes import make_cube
    return make_cube(n_bl=n_bl, n_pol=ACTIVE["n_pol"], channels=ACTIVE["channels"],
                     times=ACTIVE["times"], seed=seed, dtype=np.complex64)


def _cpu_step(args):
    """Reference path on one baseline slice: np.abs -> create_dataset -> evaluate_segmentation.
    Returns (pixels, seconds inside the two reference calls): the ground-truth mask and its random
    numbers are made OUTSIDE the timed segments, as on the GPU arm."""
    cube, truth_seed = args
    create, labels_of, evaluate, _ = _reference_api()
    np.random.seed(truth_seed)
    t0 = time.perf_counter()
    ds = create(np.abs(cube), patch_size=ACTIVE["patch"], stretch=ACTIVE["stretch"],
                flag_sigma=ACTIVE["sigma"], use_custom_flags=False)
    t1 = time.perf_counter()
    labels = labels_of(ds)
    truth = labels ^ (np.random.default_rng(truth_seed).random(labels.shape) < 0.01)
    t2 = time.perf_counter()
    evaluate(labels, truth)
    t3 = time.perf_counter()
    return cube.size, (t1 - t0) + (t3 - t2)


def cpu_baseline(n_bl=1, procs=1):
    """Times the reference's CPU path on `procs` processes, each over `n_bl` baselines.
    -> (Gpixel/s, pixels, seconds = the slowest process's time inside the reference calls)."""
    cubes = [_cpu_cube(n_bl, 100 + i)[0] for i in range(procs)]
    if procs == 1:
        npix, dt = _cpu_step((cubes[0], 1))
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_step, [(c, i) for i, c in enumerate(cubes)])
        npix, dt = sum(r[0] for r in res), max(r[1] for r in res)
    return npix / dt / 1e9, npix, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind = _reference_api()[3]
    procs = os.cpu_count() or 1
    for _ in range(1 if args.warmup >= 1 else 0):
        cpu_baseline(1, procs)
    times, npix = [], 0
    for _ in range(args.steps):
        v, npix, dt = cpu_baseline(1, procs)
        times.append(dt)
    total = sum(times)
    value = npix * len(times) / total / 1e9
    sample = (f"{procs} processes x 1 baseline x 4 pols x {ACTIVE['channels']}x{ACTIVE['times']} per step "
              f"(one Preprocessor per process, num_workers=0), "
              f"{'unmodified reference package from baseline/_ref' if kind == 'reference' else 'NumPy oracle port'}")
    line = {
        "metric": METRIC, "value": value, "unit": "Gpixel/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload_text(ACTIVE, per_gpu=False), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": procs, "kind": kind, "sample": sample},
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
class Runner:
    """K pipelined steps of create_dataset + evaluate_segmentation over one input (device or pinned
    host cube), timed with CUDA events on the current stream; max over ranks."""

    def __init__(self, torch, dist, world, dev, kw, group, lookahead):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev
        self.kw, self.group, self.lookahead = kw, group, lookahead
        self.metrics_side = METRICS_SIDE

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, seconds):
        if self.world > 1:
            t = self.torch.tensor([seconds], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return seconds

    def submit(self, data, profile=False):
        from rfi_toolbox_b200 import Preprocessor
        pre = Preprocessor(data, None, magnitude=True, pin=True)
        pre.profile = profile
        return pre, pre.create_dataset_async(**self.kw)

    def steps(self, data, truth, n, profile=False, sink=None):
        """n complete steps; returns (last metric dict, per-step kernel events, n_kept of the last step).
        Issue order of step i: result(i) [host phase + writer] -> evaluate_segmentation_async(i) [counting
        kernel + 24-byte download] -> create_dataset_async(i + lookahead + 1) -> read the metrics of step
        i - 1: every step's metric dict reaches the host, but never by draining the queue."""
        from rfi_toolbox_b200 import evaluate_segmentation, evaluate_segmentation_async
        pending, evs, m, kept, metric = deque(), [], None, 0, None
        issued = 0
        while issued < min(n, self.lookahead + (1 if self.lookahead else 0)):
            pending.append(self.submit(data, profile))
            issued += 1
        for i in range(n):
            if not self.lookahead:
                pending.append(self.submit(data, profile))
                issued += 1
            pre, pd = pending.popleft()
            np.random.seed(0)  # every step draws the same permutation (same dataset every step)
            ds = pd.result()
            if truth is not None:
                if self.lookahead:
                    nxt = evaluate_segmentation_async(ds.labels, truth, group=self.group, side_stream=self.metrics_side)
                else:
                    m = evaluate_segmentation(ds.labels, truth, group=self.group)
            if sink is not None:
                sink(ds)
            if self.lookahead and issued < n:
                pending.append(self.submit(data, profile))
                issued += 1
            if truth is not None and self.lookahead:
                if metric is not None:
                    m = metric.result()
                metric = nxt
            if profile:
                evs.append(pre.events)
            kept = len(ds)
            del ds, pre, pd  # one dataset resident at a time (configs[2]: 76 GB of patches per step)
        if metric is not None:
            m = metric.result()
        return m, evs, kept

    def timed(self, data, truth, warmup, n, profile=False, sink=None, drain=None):
        torch = self.torch
        self.steps(data, truth, warmup, sink=sink)
        if drain:
            drain()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        m, evs, kept = self.steps(data, truth, n, profile=profile, sink=sink)
        if drain:
            drain()
        e1.record()
        self.barrier()
        t1 = time.perf_counter()
        dev_s = e0.elapsed_time(e1) / 1e3
        return dict(seconds=self.max_over_ranks(dev_s), dev_ms=dev_s * 1e3, wall=(t0, t1), metrics=m, events=evs,
                    kept=kept)


def _kernel_ms(evs):
    """Mean CUDA-event time of the launches of a step: {"fused"} on the single-launch path,
    {"stats", "write"} on the two-launch path (None for what did not run)."""
    out = {"stats": [], "write": [], "fused": []}
    for ev in evs:
        for k in out:
            if k in ev:
                out[k].append(ev[k][0].elapsed_time(ev[k][1]))
    return {k: float(np.mean(v)) if v else None for k, v in out.items()}


def _path_roofline(npix, s_in, rot, k, label_px, seconds_per_step, peak):
    """Path-level HBM fraction of one step: the cube read once + every kept output written once
    (SURVEY section 8d), with and without evaluate_segmentation's 2 B per label pixel."""
    create = npix * (s_in + rot * k * 13.0)
    full = create + 2.0 * label_px
    return {"bytes_create_dataset": create, "bytes_with_metrics": full,
            "achieved_gbs_create_dataset": create / seconds_per_step / 1e9,
            "achieved_gbs_with_metrics": full / seconds_per_step / 1e9,
            "frac_create_dataset": create / seconds_per_step / 1e9 / peak,
            "frac_with_metrics": full / seconds_per_step / 1e9 / peak}


def _measure_workload(R, w, rank, peak, warmup, steps, torch):
    """One BASELINE config on this rank's GPU, inputs resident -> dict for `extra`."""
    from rfi_toolbox_b200.utils.synth import device_cube
    cube, _ = device_cube(w["n_bl"], w["n_pol"], w["channels"], w["times"], seed=1234, device=R.dev,
                          first_baseline=rank * w["n_bl"])
    npix = cube.numel()
    kw = dict(patch_size=w["patch"], stretch=w["stretch"], flag_sigma=w["sigma"], use_custom_flags=False,
              augmentation_rotations=w["rot"])
    R2 = Runner(torch, R.dist, R.world, R.dev, kw, R.group, R.lookahead)
    # ground truth of the metric = this configuration's own labels with 1 % of pixels toggled
    np.random.seed(0)
    from rfi_toolbox_b200 import Preprocessor
    ds = Preprocessor(cube, None, magnitude=True).create_dataset(**kw)
    truth = ds.labels ^ (torch.rand(ds.labels.shape, device=R.dev) < 0.01).to(torch.uint8)
    del ds
    # configs[2] holds 76 GB of patches per step: the temporaries above (23 GB of random floats, two masks) must not
    # stay cached, or a timed step runs into the allocator's free-everything-and-retry path (tens of ms, and only
    # when the side stream still holds the previous step's label block)
    torch.cuda.empty_cache()
    res = R2.timed(cube, truth, warmup, steps, profile=True)
    sec = res["seconds"] / steps
    n_tiles = npix // (w["patch"] ** 2)
    k = res["kept"] / (n_tiles * w["rot"])
    km = _kernel_ms(res["events"])
    out = {"workload": _workload_text(w), "gpixel_per_s_per_gpu": npix / sec / 1e9,
           "gpixel_per_s_all_gpus": R.world * npix / sec / 1e9, "ms_per_step": sec * 1e3, "steps": steps,
           "warmup": warmup, "patches_kept": res["kept"], "kept_fraction": k,
           "phase1_ms": km["stats"], "writer_ms": km["write"], "fused_ms": km["fused"],
           "path": _path_roofline(npix, 8, w["rot"], k, truth.numel(), sec, peak)}
    if km["fused"]:   # single launch: cube read once, every kept patch written once
        out["fused_frac_of_peak"] = npix * (8 + w["rot"] * k * 13.0) / (km["fused"] * 1e-3) / 1e9 / peak
    if km["write"]:
        out["writer_frac_of_peak"] = npix * (4 + w["rot"] * k * 13.0) / (km["write"] * 1e-3) / 1e9 / peak
    del truth, cube
    torch.cuda.empty_cache()
    return out


def _measure_c4(R, peak, torch, n_pairs, warmup, steps):
    """BASELINE configs[3]: compute_ffi + MAD reduction + IoU / F1 over n_pairs 128 x 128 pairs."""
    from rfi_toolbox_b200 import evaluate_pairs
    dev = R.dev
    g = torch.Generator(device=dev).manual_seed(7)
    true = torch.rand((n_pairs, 128, 128), generator=g, device=dev) < 0.10
    pred = true ^ (torch.rand((n_pairs, 128, 128), generator=g, device=dev) < 0.02)
    data = torch.view_as_complex(torch.randn((n_pairs, 128, 128, 2), generator=g, device=dev))
    data = (data * (1.0 + 99.0 * true)).contiguous()
    npix = true.numel()

    def one():
        r = evaluate_pairs(data, pred, true, errors="nan")   # ONE launch (rfi_pair_sweep), results read back
        return r, r

    for _ in range(warmup):
        one()
    R.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        f, m = one()
    e1.record()
    R.barrier()
    sec = R.max_over_ranks(e0.elapsed_time(e1) / 1e3) / steps
    alg = npix * 10.0  # complex64 + flags + truth, each read once
    out = {"workload": f"configs[3] {n_pairs} pairs x 128 x 128 complex64: evaluate_pairs = one rfi_pair_sweep launch "
                       "(per-pair FFI, MAD / std reduction, IoU, precision, recall, F1, dice), results downloaded every sweep",
           "gpixel_per_s_per_gpu": npix / sec / 1e9, "ms_per_sweep": sec * 1e3, "steps": steps, "warmup": warmup,
           "algorithmic_bytes": alg, "achieved_gbs": alg / sec / 1e9, "frac_of_peak": alg / sec / 1e9 / peak,
           "mean_ffi": float(np.nanmean(f["ffi"])), "mean_iou": float(np.mean(m["iou"]))}
    del true, pred, data
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

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

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")

    w = dict(WORKLOADS[args.workload])
    if args.baselines:
        w["n_bl"] = args.baselines
    # Philox seed 1234, one subsequence per baseline of the sharded cube
    cube, mask = device_cube(w["n_bl"], w["n_pol"], w["channels"], w["times"], seed=1234, device=dev,
                             first_baseline=rank * w["n_bl"])
    del mask
    npix = cube.numel()
    kw = dict(patch_size=w["patch"], stretch=w["stretch"], flag_sigma=w["sigma"], use_custom_flags=False,
              augmentation_rotations=w["rot"])
    R = Runner(torch, dist, world, dev, kw, group, args.lookahead)

    # ground truth for the metric: the labels of a first pass with 1 % of pixels toggled
    from rfi_toolbox_b200 import Preprocessor
    np.random.seed(0)
    ds = Preprocessor(cube, None, magnitude=True).create_dataset(**kw)
    truth = ds.labels ^ (torch.rand(ds.labels.shape, device=dev) < 0.01).to(torch.uint8)
    n_kept = len(ds)
    del ds

    # ---- resident-input arm
    sampler = ClockSampler(local)
    sampler.start()
    res = R.timed(cube, truth, args.warmup, args.steps, profile=True)
    sampler.stop()
    clocks = sampler.window(*res["wall"])
    elapsed = res["seconds"]
    value = world * npix * args.steps / elapsed / 1e9
    km = _kernel_ms(res["events"])
    m = res["metrics"]

    # ---- end-to-end arm: host (pinned) cube -> H2D -> create_dataset -> metrics -> D2H of the metric
    host = torch.empty(cube.shape, dtype=cube.dtype, pin_memory=True)
    host.copy_(cube)
    e2e_steps = max(1, min(args.steps, 5))
    e2e = R.timed(host, truth, max(1, min(args.warmup, 2)), e2e_steps)
    e2e_value = world * npix * e2e_steps / e2e["seconds"] / 1e9
    n_tiles = npix // (w["patch"] ** 2)
    h2d = cube.numel() * cube.element_size() + n_tiles * w["rot"] * 8
    d2h = n_tiles * 4 + 24

    # ---- end-to-end with the RESULT on the host: images + labels downloaded to pinned memory every step
    e2e_host = None
    if not args.no_extra:
        out_bytes = n_kept * w["patch"] ** 2 * 13
        try:
            himg = torch.empty((n_kept, w["patch"], w["patch"], 3), dtype=torch.float32, pin_memory=True)
            hlab = torch.empty((n_kept, w["patch"], w["patch"]), dtype=torch.uint8, pin_memory=True)
            side = torch.cuda.Stream(device=dev)
            last = []

            def sink(ds):
                for ev in last:       # one download in flight: the pinned buffers are reused
                    ev.synchronize()
                last.clear()
                _, ev = ds.to_host_async(himg, hlab, stream=side)
                last.append(ev)

            def drain():
                for ev in last:
                    ev.synchronize()
                last.clear()

            hsteps = max(1, min(args.steps, 3))
            eh = R.timed(host, truth, 1, hsteps, sink=sink, drain=drain)
            e2e_host = {"value": world * npix * hsteps / eh["seconds"] / 1e9, "unit": "Gpixel/s",
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + out_bytes), "steps": hsteps,
                        "ms_per_step": eh["seconds"] / hsteps * 1e3,
                        "note": "as e2e, plus the whole dataset (images float32 + labels uint8) copied to pinned host "
                                "memory on a side stream (TorchDataset.to_host_async), overlapped with the next step"}
            del himg, hlab
        except Exception as exc:  # e.g. the host cannot pin that much memory
            e2e_host = {"unavailable": str(exc)[:200]}
    del host

    # ---- the other BASELINE configs, same run (inputs resident)
    extra = {}
    if not args.no_extra and args.workload == "c2" and not args.baselines:
        del cube, truth
        torch.cuda.empty_cache()
        for key, name, (wu, st) in (("c3_shard", "c3", (3, 5)), ("c5_chunk", "c5", (3, 8))):
            try:
                extra[key] = _measure_workload(R, WORKLOADS[name], rank, peak, wu, st, torch)
            except Exception as exc:
                extra[key] = {"unavailable": str(exc)[:300]}
        try:
            extra["c4_sweep"] = _measure_c4(R, peak, torch, 100_000, 3, 5)
        except Exception as exc:
            extra["c4_sweep"] = {"unavailable": str(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    k = n_kept / (n_tiles * w["rot"])
    sec_step = elapsed / args.steps
    # DRAM bytes of one launch from the committed `ncu --set full` capture of this very workload
    # (profiles/traffic.json, written by scripts/ncu_summary.py); null for any other workload
    def committed_traffic(kname):
        if args.baselines or args.workload not in ("c2", "c5"):
            return None
        for name in ("r02c_traffic.json", "r02_traffic.json", "traffic.json"):   # the newest committed capture that holds this kernel
            tj = ROOT / "profiles" / name
            if tj.exists():
                v = json.loads(tj.read_text()).get(kname, {}).get("dram_bytes_per_launch")
                if v is not None:
                    return v
        return None
    if km["fused"]:
        # single launch per step (tile_fused_kernel): the cube is read once (8 B / px), every kept patch
        # is written once (13 B per output px), nothing else moves: the path minimum of SURVEY 8d
        kname = "tile_fused_kernel"
        alg_bytes = npix * (8 + w["rot"] * k * 13.0)
        kernel_ms = km["fused"]
        bpp = {"fused": 8 + w["rot"] * k * 13.0, "path_minimum": 8 + w["rot"] * k * 13.0}
    else:
        # two launches; complex64 through the real branch: the writer reads the exact float32 magnitudes
        # phase 1 left in its scratch (4 B / px), the cube itself is read once, by phase 1
        kname = "write_patches_kernel" if w["patch"] == 128 else "big_write_kernel"
        alg_bytes = npix * (4 + w["rot"] * k * 13.0)
        kernel_ms = km["write"]
        bpp = {"phase1": 8 + 4, "writer": 4 + w["rot"] * k * 13.0, "path_minimum": 8 + w["rot"] * k * 13.0}
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": committed_traffic(kname),
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                "stats_kernel_ms": km["stats"], "bytes_per_pixel": bpp,
                "step_ms_device": res["dev_ms"] / args.steps,
                "path": _path_roofline(npix, 8, w["rot"], k, n_kept * w["patch"] ** 2, sec_step, peak)}
    if km["stats"]:
        roofline["stats_kernel_gbs"] = npix * (8 + 4) / (km["stats"] * 1e-3) / 1e9
    launches_per_step = (2 if km["fused"] else 3) if w["patch"] == 128 else 10

    cpu_bl = 2 if w["channels"] * w["times"] <= 1 << 20 else 1
    cpu_v, cpu_npix, cpu_dt = cpu_baseline(cpu_bl, 1)
    line = {
        "metric": METRIC, "value": value, "unit": "Gpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload_text(w),
                   "pixels_per_step_per_gpu": npix, "patches_kept": n_kept,
                   "launches": ("single launch per create_dataset (tile_fused_kernel: statistics + patches of a tile in one CTA; "
                                "destination slots drawn ahead, flag counts checked in result())") if km["fused"] else
                               "two launches per create_dataset (statistics, host shuffle, writer)",
                   "issue": (f"streaming: create_dataset_async with {args.lookahead} call(s) in flight (the host side of step k "
                             "behind the kernels of step k+1; metrics of step k read after step k+1 is enqueued); K complete steps incl. "
                             "fill and drain inside the timed region")
                            if args.lookahead else "sequential: create_dataset, then evaluate_segmentation, per step",
                   "l2": f"input cube {npix * 8 / 1e9:.1f} GB and {n_kept * w['patch'] ** 2 * 13 / 1e9:.1f} GB of output "
                         "per step exceed the 126 MB L2; no flush needed",
                   "parallelism": (f"baseline-sharded x{world}; TP/FP/FN summed inside the counting kernel over NVLink peer "
                                   "memory (NCCL only for rendezvous / barriers)") if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": "Gpixel/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "note": "pinned host cube copied H2D every step; dataset stays in HBM, metric dict read back"},
        "e2e_host_result": e2e_host,
        # per step, P = 128: tile_fused_kernel + confusion_kernel (two-launch path: tile_stats_mono_kernel,
        # write_patches_kernel, confusion_kernel);
        # P = 256: big_init / load / sample / pass<0> / median / pass<1> / mad / count / write + confusion_kernel
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "extra": extra,
        "cpu_baseline": {"value": cpu_v, "unit": "Gpixel/s", "cores": 1, "kind": _reference_api()[3],
                         "sample": f"{cpu_bl} baselines x 4 pols x {w['channels']}x{w['times']} ({cpu_npix} px) in {cpu_dt:.1f} s, one process"},
        "host_affinity": (f"{len(cpus)} cores near GPU {local}" if cpus else None),
        "clocks": clocks, "wall_s": res["wall"][1] - res["wall"][0],
        "metrics_last_step": {k_: float(v) for k_, v in m.items()},
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
    ap.add_argument("--lookahead", type=int, default=2, help="create_dataset_async calls kept in flight (0 = sequential)")
    ap.add_argument("--metrics-stream", default="side", choices=["side", "main"],
                    help="side: the counting kernel of step k runs on a side stream, next to the statistics kernel of step k+2")
    ap.add_argument("--single-launch", action="store_true",
                    help="A/B: opt into the single-launch path (Preprocessor.speculate = True; falls back per call when a tile is blank)")
    ap.add_argument("--no-extra", action="store_true", help="skip e2e_host_result and the c3 / c5 / c4 measurements")
    args = ap.parse_args()
    global ACTIVE, METRICS_SIDE
    ACTIVE = WORKLOADS[args.workload]
    METRICS_SIDE = args.metrics_stream == "side"
    if args.single_launch:
        from rfi_toolbox_b200 import Preprocessor
        Preprocessor.speculate = True
    _reserve_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
