#!/usr/bin/env python
"""bench.py -- headline benchmark of the raster-filter hot path (BASELINE.json).

Workload (config 4): a 16384 x 16384 synthetic 25 m DEM (1 GiB, larger than the 126 MB L2), one
"step" = the multi-scale sweep TPI + STD + gradient/slope/aspect at 100 m ... 20 km, i.e. disc
diameters / Gaussian radii {5, 9, 13, 21, 41, 81, 161, 241, 401, 801} px = 30 descriptor calls.
Metric: DEM Mpixel/s per descriptor call = calls * ny * nx / time.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference's CPU algorithm (oracle port)

N > 1: launched under torchrun, one rank per GPU; the DEM is split in row bands, halos travel over
NVLink (torch.distributed P2P / NCCL), strong scaling (fixed DEM).  Rank 0 prints ONE JSON line.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCALES_M = [100, 200, 300, 500, 1000, 2000, 4000, 6000, 10000, 20000]
RES_M = 25.0
CPU_CROP = 2048  # the CPU arm runs the same sweep on a CPU_CROP^2 crop (bounded sample)


def sizes_for(scales, res):
    from topo_descriptors_b200.helpers import round_up_to_odd

    return [int(s) for s in round_up_to_odd(np.array(scales) / res)]


def make_dem_rows(ny, nx, r0, r1, seed=2):
    """Rows [r0, r1) of the deterministic integer-valued (SRTM-like) synthetic DEM."""
    from topo_descriptors_b200.synth import tiled_fractal_dem

    return tiled_fractal_dem(ny, nx, seed=seed, tile=2048, integer=True, rows=(r0, r1))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms during the timed region."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.index = index
        self.samples, self.mask = [], 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        reasons = [name for bit, name in self.REASONS.items() if self.mask & bit]
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle/oracle.py *_literal = the reference's own scipy/numpy call
# sequence; /root/reference itself cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------
_CPU_DEM = None
_CPU_RES = None


def _cpu_call(task):
    from oracle import oracle as O

    kind, size = task
    t = time.perf_counter()
    if kind == "tpi":
        O.tpi_literal(_CPU_DEM, size)
    elif kind == "std":
        O.std_literal(_CPU_DEM, size)
    else:
        O.gradient_literal(_CPU_DEM, size / 4.0, _CPU_RES)
    return kind, size, time.perf_counter() - t


def cpu_sweep(sizes, crop, workers, repeats=1):
    """Time the reference algorithm for the same sweep on a crop x crop window of the same DEM.
    Independent descriptor calls are spread over `workers` processes (the reference itself is
    single-threaded on these paths; this is the most host parallelism it can use).  Returns
    (Mpixel/s per descriptor call, seconds per sweep, per-call seconds)."""
    import multiprocessing as mp

    global _CPU_DEM, _CPU_RES
    _CPU_DEM = make_dem_rows(crop, crop, 0, crop)
    _CPU_RES = {"x": np.full(crop, RES_M), "y": np.full(crop, -RES_M)}
    tasks = [(k, s) for s in sorted(sizes, reverse=True) for k in ("gradient", "std", "tpi")]
    best = None
    per_call = {}
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(repeats):
            t = time.perf_counter()
            res = pool.map(_cpu_call, tasks, chunksize=1)
            dt = time.perf_counter() - t
            if best is None or dt < best:
                best = dt
                per_call = {f"{k}_{s}": round(sec, 4) for k, s, sec in res}
    mpix = len(tasks) * crop * crop / best / 1e6
    return mpix, best, per_call


def run_reference(args, rank, world):
    if rank != 0:
        return
    sizes = sizes_for(SCALES_M, RES_M)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 3 * len(sizes)))
    # warm-up + timed "steps": each step is one sweep over the bounded sample
    times = []
    for i in range(args.warmup + args.steps):
        mpix, sec, per_call = cpu_sweep(sizes, CPU_CROP, workers)
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    value = 3 * len(sizes) * CPU_CROP * CPU_CROP / sec / 1e6
    sample = f"same sweep on a {CPU_CROP}x{CPU_CROP} crop of the same DEM, {workers} worker processes"
    line = {
        "impl": "reference", "metric": "DEM Mpixel/s per descriptor call (TPI+STD+gradient multi-scale sweep)",
        "value": round(value, 3), "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 in, f64 accumulate", "data": "synthetic",
        "config": workload_config(16384, 16384, sizes, extra={"cpu_sample": sample}),
        "cpu_baseline": {"value": round(value, 3), "unit": "Mpixel/s", "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(ny, nx, sizes, extra=None):
    cfg = {
        "workload": f"config 4: {ny}x{nx} {RES_M:g} m DEM, multi-scale TPI/STD/gradient sweep 100 m-20 km",
        "scales_m": SCALES_M, "sizes_px": sizes, "calls_per_step": 3 * len(sizes),
        "dem": "synthetic fractal, integer-valued metres (SRTM-like), float32",
        "l2": f"input {ny * nx * 4 / 2**20:.0f} MiB > 126 MB L2, every call re-reads it (no flush needed)",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from topo_descriptors_b200 import _lib, bands, device as dev

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_bound = bands.bind_to_gpu_numa(local_rank) if world > 1 else False
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    ny = nx = args.size
    sizes = sizes_for(SCALES_M, RES_M)
    sigmas = [s / 4.0 for s in sizes]
    n_calls = 3 * len(sizes)
    ctx = bands.BandContext(ny, nx, rank, world)

    host = torch.from_numpy(make_dem_rows(ny, nx, ctx.r0, ctx.r1)).pin_memory()
    core = host.to(device, non_blocking=True)
    res_x = (dev._Res(np.full(nx, RES_M), device), 0)
    res_y = (dev._Res(np.full(ny, -RES_M), device), 0)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return bands.sweep(core, ctx, sizes, sigmas, res_x, res_y)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = (_lib.launch_count() - launches0) // max(args.steps, 1)
    ms_step = ms_total / args.steps
    value = n_calls * ny * nx / (ms_step * 1e-3) / 1e6

    # ---- per-kernel attribution of one more step (CUDA events around every launch, on its stream)
    _lib.profile_enable(True)
    step()
    torch.cuda.synchronize()
    prof = _lib.profile_dump()
    _lib.profile_enable(False)
    barrier()

    # ---- e2e: host (pinned) -> HBM -> sweep -> every output back to the host, per step.  The D2H copies run
    # on a second stream into a ring of pinned buffers so that PCIe traffic overlaps the kernels; the compute
    # stream is throttled to at most `kInFlight` outputs waiting for their copy.
    kInFlight = 6
    pinned_ring = [torch.empty((ctx.rows, nx), dtype=torch.float32).pin_memory() for _ in range(3)]
    copy_stream = torch.cuda.Stream(device=device)
    d2h = [0]
    done = []

    def sink(name, i, t):
        cur = torch.cuda.current_stream()
        ready = cur.record_event()
        copy_stream.wait_event(ready)
        with torch.cuda.stream(copy_stream):
            pinned_ring[len(done) % len(pinned_ring)].copy_(t, non_blocking=True)
            done.append(copy_stream.record_event())
        t.record_stream(copy_stream)
        d2h[0] += t.numel() * 4
        if len(done) > kInFlight:
            cur.wait_event(done[len(done) - 1 - kInFlight])

    def e2e_step():
        c = host.to(device, non_blocking=True)
        n = bands.sweep(c, ctx, sizes, sigmas, res_x, res_y, sink=sink)
        torch.cuda.current_stream().wait_stream(copy_stream)
        return n

    e2e_steps = max(1, min(args.steps, 2))
    e2e_step()
    barrier()
    d2h[0] = 0
    done.clear()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_calls * ny * nx / (float(e2e_ms.item()) / e2e_steps * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    band_px = ctx.rows * nx
    total_kernel_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    dom_name, dom_v = dom
    avg_ms = dom_v["ms"] / dom_v["launches"]

    def alg_bytes_px(name):
        """Algorithmic HBM bytes per output pixel of one launch."""
        if name.startswith(("grad_from_smooth", "sobel_gradient")):
            return 20  # 4 B read + 4 outputs
        if name.startswith("stats"):
            return 4
        return 8  # every other kernel on this path: 4 B in, 4 B out

    achieved = alg_bytes_px(dom_name) * band_px / (avg_ms * 1e-3) / 1e9
    # DRAM bytes per algorithmic byte from the committed `ncu --set full` capture (profiles/r01_ncu_full_summary.csv,
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch on an 8192^2 DEM / its algorithmic bytes)
    ncu_dram_ratio = {"gauss_axis0": 1.11, "grad_from_smooth": 0.96, "disc_hybrid": 21.0}
    ratio = next((r for k, r in ncu_dram_ratio.items() if dom_name.startswith(k)), None)
    traffic = None if ratio is None else int(ratio * alg_bytes_px(dom_name) * band_px)
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic,
        "traffic_source": "algorithmic bytes of one launch x the DRAM/algorithmic ratio of the ncu --set full capture in "
                          "profiles/r01_ncu_full_summary.csv" if traffic is not None else None,
        "peak_source": peak_src,
        "share_of_step": round(dom_v["ms"] / total_kernel_ms, 3),
        "note": "the kernels that dominate config 4 are the wide-radius ones (float64 Gaussian taps, disc span gathers): "
                "FP64-pipe / L1-wavefront bound by construction, so their HBM fraction is small; the HBM-bound kernels of "
                "the path are listed in `memory_bound` (DESIGN.md section 4)",
    }
    if dom_name.startswith("gauss_axis0"):
        # float64 FMA roofline of the Gaussian: taps walked per pixel and launch (K + 2*lw rounded up to K = 16)
        steps = [((16 + 2 * int(4.0 * sg + 0.5) + 15) // 16) * 16 for sg in sigmas]
        n_launch = {True: 2, False: 1}
        flops = sum(2.0 * band_px * st * n_launch[int(4.0 * sg + 0.5) > 64] for st, sg in zip(steps, sigmas))
        roofline["fp64"] = {"achieved_tflops": round(flops / (dom_v["ms"] * 1e-3) / 1e12, 2),
                            "peak_tflops": 29.7, "peak_source": "profiles/micro/pipes2.cu: 51 DFMA/clk/SM x 148 SM x 1.965 GHz"}
    kernels = {
        k: {"launches": v["launches"], "ms": round(v["ms"], 3), "avg_ms": round(v["ms"] / v["launches"], 4),
            "max_ms": round(v["max_ms"], 3),
            "avg_GBps_alg": round(alg_bytes_px(k) * band_px / (v["ms"] / v["launches"] * 1e-3) / 1e9, 1)}
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    }
    memory_bound = {
        k: {"GBps_alg": v["avg_GBps_alg"], "frac_of_hbm_peak": round(v["avg_GBps_alg"] / peak, 3)}
        for k, v in kernels.items()
        if k.startswith(("stats_partial", "grad_from_smooth", "sobel_gradient", "disc_tiny", "disc_prefix", "transpose"))
    }

    line = {
        "metric": "DEM Mpixel/s per descriptor call (TPI+STD+gradient multi-scale sweep)",
        "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 in/out; u32 fixed-point + i64 sums (tpi/std), f64 accumulate (gaussian)", "data": "synthetic",
        "config": workload_config(ny, nx, sizes, extra={"parallelism": f"row bands x{world}, halo exchange over NVLink",
                                                            "numa_bound": bool(numa_bound)}),
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": int(ctx.rows * nx * 4),
                "d2h_bytes_per_step": int(d2h[0] // e2e_steps), "steps": e2e_steps,
                "path": "pinned host DEM -> HBM -> bands.sweep -> every output band back to pinned host memory (D2H on a second stream, overlapped)"},
        "gpu_launches": int(launches * args.steps),
        "roofline": roofline,
        "memory_bound": memory_bound,
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, n_calls))
        mpix, sec, per_call = cpu_sweep(sizes, CPU_CROP, workers)
        line["cpu_baseline"] = {
            "value": round(mpix, 3), "unit": "Mpixel/s", "cores": workers, "kind": "port",
            "sample": f"same sweep (reference call sequence on scipy/numpy, oracle/*_literal) on a {CPU_CROP}x{CPU_CROP} crop, "
                      f"{sec:.1f} s, {workers} worker processes of {cores} cores",
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="DEM edge in pixels (default: config 4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
