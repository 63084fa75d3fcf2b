#!/usr/bin/env python
"""bench.py -- headline benchmark of the raster-filter hot path (BASELINE.json).

Workload (config 4): a 16384 x 16384 synthetic 25 m DEM (1 GiB, larger than the 126 MB L2), one
"step" = the multi-scale sweep TPI + STD + gradient/slope/aspect at 100 m ... 20 km, i.e. disc
diameters / Gaussian radii {5, 9, 13, 21, 41, 81, 161, 241, 401, 801} px = 30 descriptor calls.
Metric: DEM Mpixel/s per descriptor call = calls * ny * nx / time.  The headline runs on the FLOAT DEM (the
reference's general case: 3 planes for STD, fixed-point TPI); the same sweep on the integer-valued (SRTM-like) DEM,
config 3 (Sx, 72 azimuths, dealt over the GPUs) and config 5 (valley/ridge size 41 + Sx 10 km, row bands) ride along
in `extra`, each with its own roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # the reference itself on the host cores

N > 1: launched under torchrun, one rank per GPU; the DEM is split in row bands, halos travel over
NVLink (torch.distributed P2P / NCCL), strong scaling (fixed DEM).  Rank 0 prints ONE JSON line.
"""

import argparse
import csv
import glob
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
C3_SIZE, C3_RES, C3_RADIUS = 4096, 30.0, 500.0
C5_SIZE, C5_KSIZE, C5_RADIUS = 8192, 41, 10000.0
METRIC = "DEM Mpixel/s per descriptor call (TPI+STD+gradient multi-scale sweep)"


def sizes_for(scales, res):
    from topo_descriptors_b200.helpers import round_up_to_odd

    return [int(s) for s in round_up_to_odd(np.array(scales) / res)]


def make_dem_rows(ny, nx, r0, r1, seed=2, integer=True):
    """Rows [r0, r1) of the deterministic synthetic DEM (integer-valued = SRTM-like, or float)."""
    from topo_descriptors_b200.synth import tiled_fractal_dem

    return tiled_fractal_dem(ny, nx, seed=seed, tile=2048, integer=integer, rows=(r0, r1))


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
# CPU arm.  kind "reference": the UNMODIFIED reference package (oracle/ref_runner: /root/reference here, the
# staged archive oracle/_ref on the GPU box) called through its own public functions; kind "port": the oracle's
# restatement of the same scipy/numpy call sequence (oracle/*_literal), only when the archive is missing.
# ---------------------------------------------------------------------------------------------
_CPU_DEM = None
_CPU_RES = None
_CPU_REF = None


def cpu_kind():
    from oracle import ref_runner

    return "reference" if ref_runner.available() else "port"


def _cpu_init():
    global _CPU_REF
    if _CPU_REF is None:
        from oracle import ref_runner

        if ref_runner.available():
            import logging

            logging.disable(logging.INFO)
            _CPU_REF = ref_runner.load()[0]
        else:
            _CPU_REF = False
    return _CPU_REF


def _cpu_call(task):
    kind, size = task
    ref = _cpu_init()
    t = time.perf_counter()
    if ref:
        if kind == "tpi":
            ref.tpi(_CPU_DEM, size)
        elif kind == "std":
            ref.std(_CPU_DEM, size)
        else:
            ref.gradient(_CPU_DEM, size / 4.0, _CPU_RES)
    else:
        from oracle import oracle as O

        if kind == "tpi":
            O.tpi_literal(_CPU_DEM, size)
        elif kind == "std":
            O.std_literal(_CPU_DEM, size)
        else:
            O.gradient_literal(_CPU_DEM, size / 4.0, _CPU_RES)
    return kind, size, time.perf_counter() - t


def cpu_sweep(sizes, crop, workers, integer=False):
    """Time the reference for the same sweep on a crop x crop window of the same DEM.  Independent descriptor calls
    are spread over `workers` processes (the reference itself is single-threaded on these paths; this is the most
    host parallelism it can use).  Returns (Mpixel/s per descriptor call, seconds per sweep, per-call seconds)."""
    import multiprocessing as mp

    global _CPU_DEM, _CPU_RES
    _CPU_DEM = make_dem_rows(crop, crop, 0, crop, integer=integer)
    _CPU_RES = {"x": np.full(crop, RES_M), "y": np.full(crop, -RES_M)}
    tasks = [(k, s) for s in sorted(sizes, reverse=True) for k in ("gradient", "std", "tpi")]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        t = time.perf_counter()
        res = pool.map(_cpu_call, tasks, chunksize=1)
        dt = time.perf_counter() - t
    per_call = {f"{k}_{s}": round(sec, 4) for k, s, sec in res}
    return len(tasks) * crop * crop / dt / 1e6, dt, per_call


def cpu_sx(n_az=4, crop=1024):
    """Reference topo.sx (numba prange over rows: all host cores) for `n_az` of config 3's azimuths on a crop."""
    from oracle import ref_runner
    from topo_descriptors_b200.synth import fractal_dem

    ref = _cpu_init()
    z = fractal_dem(crop, crop, seed=1)
    x = 2600000.0 + C3_RES * np.arange(crop)
    y = 1200000.0 - C3_RES * np.arange(crop)
    if ref:
        ds = ref_runner.fake_dataset(z, x, y)
        ref.sx(ds, 0.0, C3_RADIUS)  # numba JIT, excluded
        t = time.perf_counter()
        for az in np.arange(n_az) * 5.0:
            ref.sx(ds, az, C3_RADIUS)
    else:
        from oracle import oracle as O

        t = time.perf_counter()
        for az in np.arange(n_az) * 5.0:
            O.sx_exact(z, x, y, az, C3_RADIUS)
    dt = time.perf_counter() - t
    return n_az * crop * crop / dt / 1e6, dt


def cpu_valley(crop=320):
    """Reference valley_ridge size 41 (180 rotations + 3-D float32 FFT convolutions, one thread) on a crop."""
    from topo_descriptors_b200.synth import fractal_dem

    ref = _cpu_init()
    z = fractal_dem(crop, crop, seed=3)
    t = time.perf_counter()
    if ref:
        ref.valley_ridge(z, C5_KSIZE, "valley")
    else:
        from oracle import oracle as O

        O.valley_ridge_literal(z, C5_KSIZE, "valley")
    dt = time.perf_counter() - t
    return crop * crop / dt / 1e6, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    sizes = sizes_for(SCALES_M, RES_M)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 3 * len(sizes)))
    times = []
    for i in range(args.warmup + args.steps):  # each step = one sweep over the bounded sample
        mpix, sec, per_call = cpu_sweep(sizes, CPU_CROP, workers, integer=False)
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    value = 3 * len(sizes) * CPU_CROP * CPU_CROP / sec / 1e6
    kind = cpu_kind()
    sample = (f"same sweep on a {CPU_CROP}x{CPU_CROP} crop of the same float DEM through the "
              f"{'unmodified reference package (topo_descriptors.topo.tpi/std/gradient)' if kind == 'reference' else 'oracle port of its call sequence'}"
              f", {workers} worker processes; Mpixel/s is per-pixel extrapolated to the full DEM")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": round(value, 3), "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 in, f32 FFT / f64 accumulate (scipy)", "data": "synthetic",
        "config": workload_config(16384, 16384, sizes, "float"),
        "cpu_baseline": {"value": round(value, 3), "unit": "Mpixel/s", "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(ny, nx, sizes, dem_kind):
    return {
        "workload": f"config 4: {ny}x{nx} {RES_M:g} m DEM, multi-scale TPI/STD/gradient sweep 100 m-20 km",
        "scales_m": SCALES_M, "sizes_px": sizes, "calls_per_step": 3 * len(sizes),
        "dem": ("synthetic fractal, float32 metres (the reference's general case)" if dem_kind == "float"
                else "synthetic fractal, integer-valued metres (SRTM-like), float32"),
        "l2": f"input {ny * nx * 4 / 2**20:.0f} MiB > 126 MB L2, every call re-reads it (no flush needed)",
    }


# ---------------------------------------------------------------------------------------------
# ncu evidence: DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summaries
# ---------------------------------------------------------------------------------------------
def ncu_dram_ratio(kernel_name):
    """(DRAM bytes / algorithmic bytes, source file) of the kernel whose bench label is `kernel_name`, from the newest
    profiles/r*_ncu_*_summary.csv that holds it.  The summaries carry a `#pixels=` comment line or a sidecar
    .meta.json with the pixel count of the captured launch."""
    base = kernel_name.split("<")[0]
    sass_name = {"disc_hybrid": "disc_span_kernel", "disc_span": "disc_span_kernel", "grad_from_smooth": "gradient_kernel<0>",
                 "sobel_gradient": "gradient_kernel<1>", "valley_bank": "valley_kernel", "disc_fft_inv": "fft2d_inv_product_kernel",
                 "disc_fft_finish": "dfft_store_kernel", "disc_fft_store": "dfft_store_kernel",
                 "disc_fft_transpose": "fft2d_transpose_kernel", "disc_fft_planes": "dfft_fwd_planes_kernel",
                 "disc_fft_fwd": "fft2d_fwd_cplx_kernel", "gauss_fft": "fft_conv_kernel"}.get(base, base + "_kernel")
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_*summary.csv")), reverse=True):
        meta = path.replace(".csv", ".meta.json")
        if not os.path.exists(meta):
            continue
        m = json.load(open(meta))
        best = None
        with open(path) as f:
            rows = list(csv.reader(f))
        head = rows[0]
        try:
            k, rd, wr, tm = head.index("kernel"), head.index("dram_read [Gbyte]"), head.index("dram_write [Gbyte]"), head.index("time [ms]")
        except ValueError:
            continue
        for r in rows[1:]:
            if len(r) <= max(k, rd, wr) or sass_name not in r[k]:
                continue
            if best is None or float(r[tm]) > float(best[tm]):
                best = r
        if best is not None:
            dram = (float(best[rd]) + float(best[wr])) * 1e9
            return dram / (float(m["alg_bytes_px"].get(base, 8)) * float(m["pixels"])), os.path.relpath(path, ROOT)
    return None, None


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import ctypes

    import torch
    import torch.distributed as dist

    from topo_descriptors_b200 import _lib, _xr, bands, device as dev, topo
    from topo_descriptors_b200.device import DeviceDEM
    from topo_descriptors_b200.synth import fractal_dem

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
    res_x = (dev._Res(np.full(nx, RES_M), device), 0)
    res_y = (dev._Res(np.full(ny, -RES_M), device), 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn, warmup, steps, clocks_index=None):
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(clocks_index) if clocks_index is not None else None
        if sampler:
            sampler.__enter__()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        if sampler:
            sampler.__exit__()
        return max_over_ranks(ev0.elapsed_time(ev1)) / steps, sampler

    # ---------------- headline: float DEM sweep, inputs resident in HBM ---------------------------------
    host = torch.from_numpy(make_dem_rows(ny, nx, ctx.r0, ctx.r1, integer=False)).pin_memory()
    core = host.to(device, non_blocking=True)
    torch.cuda.synchronize()

    # N > 1: thin bands make the ~360 launches of a step latency-visible, so the static kernel sequence is captured in
    # a CUDA graph (halo exchange over NCCL and the statistics all-reduce stay outside it, every step)
    # (N = 1 too: the kernels of a step sum to the same time either way, but eager launches leave up to 10 ms of host-side
    # gaps per step on a slow host -- measured 163.8 vs 174.4 ms on two boxes with identical kernel times)
    graph = bands.SweepGraph() if not args.no_graph else None

    def step():
        if graph is not None:
            return graph.run(core, ctx, sizes, sigmas, res_x, res_y)
        return bands.sweep(core, ctx, sizes, sigmas, res_x, res_y)

    graph_note = None
    if graph is not None and world == 1:
        # insurance on one GPU (whole-image workspaces inside the graph's memory pool): a capture that fails on this box
        # must not cost the whole run -- the eager sweep does the same work
        try:
            step()
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            graph_note = f"graph capture failed ({type(exc).__name__}: {str(exc)[:120]}): eager launches"
            graph = None
            torch.cuda.empty_cache()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _lib.launch_count()
    ms_step, clocks = timed_steps(step, 0, args.steps, clocks_index=local_rank)
    launches = _lib.launch_count() - launches0 + (graph.launches * args.steps if graph is not None else 0)
    used_graph = graph is not None
    graph = None  # (frees the graph's memory pool: tens of GB of workspaces on one GPU)
    torch.cuda.empty_cache()
    value = n_calls * ny * nx / (ms_step * 1e-3) / 1e6

    # ---- per-kernel attribution of one more step (CUDA events around every launch, on its stream)
    # (always eager: the launches inside a replayed graph are invisible to the per-launch events)
    _lib.profile_enable(True)
    bands.sweep(core, ctx, sizes, sigmas, res_x, res_y)
    torch.cuda.synchronize()
    prof = _lib.profile_dump()
    _lib.profile_enable(False)
    barrier()

    # ---- e2e: host (pinned) -> HBM -> sweep -> every output back to the host, per step.  The D2H copies run
    # on a second stream into a ring of pinned buffers so that PCIe traffic overlaps the kernels; the compute
    # stream is throttled to at most `kInFlight` outputs waiting for their copy.
    # N > 1: this leg is bound by getting 60 GiB of results off the GPUs, and the ranks do not drain at the same rate
    # when all copy at once (measured on this box class: 12.1 GB/s for GPUs 0-3, 18.1 GB/s for GPUs 4-7, 121 GB/s in
    # total, against 56 GB/s for one GPU alone: profiles/r02_d2h_probe_n8.json) -- so the bands of this leg are sized in
    # proportion to each rank's concurrent device -> host rate, probed here with a 256 MiB copy
    ctx_dev, host_dev = ctx, host
    d2h_rates = None
    if world > 1:
        probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        probe_h.copy_(probe_d, non_blocking=True)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(4):
            probe_h.copy_(probe_d, non_blocking=True)
        q1.record()
        torch.cuda.synchronize()
        rate = torch.tensor([4 * (256 << 20) / (q0.elapsed_time(q1) * 1e-3) / 1e9], dtype=torch.float64, device=device)
        rates = [torch.zeros_like(rate) for _ in range(world)]
        dist.all_gather(rates, rate)
        d2h_rates = [round(float(r.item()), 1) for r in rates]
        del probe_d, probe_h
        ctx = bands.BandContext(ny, nx, rank, world, weights=d2h_rates)
        host = torch.from_numpy(make_dem_rows(ny, nx, ctx.r0, ctx.r1, integer=False)).pin_memory()
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
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    e2e_value = n_calls * ny * nx / (e2e_ms / e2e_steps * 1e-3) / 1e6
    e2e_d2h = torch.tensor([float(d2h[0] // e2e_steps), float(ctx.rows * nx * 4)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_d2h)  # bytes of the whole job
    e2e_d2h, e2e_h2d = int(e2e_d2h[0].item()), int(e2e_d2h[1].item())
    e2e_rows = [p[1] - p[0] for p in ctx.parts]
    del pinned_ring
    ctx, host = ctx_dev, host_dev

    # ---- N > 1: the bands of this very run against an independent whole-image-coordinates strip (bit for bit)
    band_check = None
    if world > 1:
        band_check = band_spot_check(core, ctx, sizes, sigmas, res_x, res_y, rank, device)
    del core, host
    torch.cuda.empty_cache()

    extra = {}
    # ---------------- extra: the same sweep on the integer-valued DEM -------------------------------------
    if not args.no_extra:
        core_i = torch.from_numpy(make_dem_rows(ny, nx, ctx.r0, ctx.r1, integer=True)).to(device)
        graph_i = bands.SweepGraph() if used_graph else None
        ms_i, _ = timed_steps((lambda: graph_i.run(core_i, ctx, sizes, sigmas, res_x, res_y)) if graph_i is not None else
                              (lambda: bands.sweep(core_i, ctx, sizes, sigmas, res_x, res_y)), 2, max(2, min(args.steps, 3)))
        graph_i = None
        extra["sweep_integer_dem"] = {
            "config": workload_config(ny, nx, sizes, "integer"), "ms_per_step": round(ms_i, 3),
            "value": round(n_calls * ny * nx / (ms_i * 1e-3) / 1e6, 1), "unit": "Mpixel/s",
            "note": "STD_I / TPI_I: exact integer planes, tpi(size) and std(size) share the T-plane sums",
        }
        del core_i
        torch.cuda.empty_cache()

        # ------------ extra: config 3, Sx radius 500 m, azimuths 0..355 step 5 on 4096^2, azimuths dealt over the GPUs
        z3 = fractal_dem(C3_SIZE, C3_SIZE, seed=1)
        x3 = 2600000.0 + C3_RES * np.arange(C3_SIZE, dtype=np.float64)
        y3 = 1200000.0 - C3_RES * np.arange(C3_SIZE, dtype=np.float64)
        grid3 = _xr.Dataset({"alti": (("y", "x"), np.zeros((1, 1), np.float32))}, coords={"x": x3, "y": y3},
                            attrs={"crs": "epsg:2056"})
        azs = list(range(0, 360, 5))
        mine = bands.azimuth_share(azs, ctx)
        plan3 = topo._sx_plan(grid3, mine, C3_RADIUS, 10.0, 15, 0.0)
        d3 = DeviceDEM(torch.from_numpy(z3).to(device))
        ms3, _ = timed_steps(lambda: topo._sx_device(d3, plan3, 10.0), 2, 5)
        samples = torch.tensor([float(plan3[2][-1])], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(samples)
        interior = float(C3_SIZE - 2 * plan3[3]) ** 2
        clk = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
        rate = float(samples.item()) * interior / (ms3 * 1e-3) / (148.0 * world) / clk
        extra["config3_sx"] = {
            "workload": f"config 3: topo.sx radius {C3_RADIUS:g} m, azimuths 0-355 step 5 on a {C3_SIZE}x{C3_SIZE} {C3_RES:g} m DEM, "
                        f"azimuths dealt round-robin over {world} GPU(s), no communication",
            "ms_per_step": round(ms3, 3), "value": round(len(azs) * C3_SIZE * C3_SIZE / (ms3 * 1e-3) / 1e6, 1),
            "unit": "Mpixel/s per azimuth call", "scaling": "strong",
            "roofline": {"bound": "issue", "achieved": round(rate, 2), "peak": 24.0, "unit": "ray samples / clk / SM",
                         "frac": round(rate / 24.0, 3),
                         "note": "per sample and 8 rows: 1 LDS.128 + 1 IADD + 8 x (LDS + FADD + FFMA + FMNMX) = 34 issue slots -> "
                                 "4 schedulers x 8 rows / 34 x ... ~24 samples/clk/SM; HBM traffic is (4 + 4A)/A B/px, irrelevant here"},
        }
        del d3
        torch.cuda.empty_cache()

        # ------------ extra: config 5 kernels on a C5_SIZE^2 DEM in row bands
        n5 = C5_SIZE
        ctx5 = bands.BandContext(n5, n5, rank, world)
        core5 = torch.from_numpy(make_dem_rows(n5, n5, ctx5.r0, ctx5.r1, seed=3, integer=False)).to(device)
        x5 = 2600000.0 + RES_M * np.arange(n5, dtype=np.float64)
        y5 = 1200000.0 - RES_M * np.arange(n5, dtype=np.float64)
        grid5 = _xr.Dataset({"alti": (("y", "x"), np.zeros((1, 1), np.float32))}, coords={"x": x5, "y": y5},
                            attrs={"crs": "epsg:2056"})
        plan5 = topo._sx_plan(grid5, [270.0], C5_RADIUS, 10.0, 15, 0.0)
        flats = [0, 0.15, 0.3]
        bank = topo._device_bank(C5_KSIZE, "valley", flats, device)  # bank build is setup, cached per (size, mode, flats)
        ms_v, _ = timed_steps(lambda: bands.valley_ridge_band(core5, ctx5, C5_KSIZE, "valley", flats), 1, 1)
        ms_s, _ = timed_steps(lambda: bands.sx_band(core5, ctx5, plan5, 10.0), 1, 3)
        macs = float(sum(int(h) * int(w) for h, w, _a in bank["plain"]["hw"]))  # taps of all 180 x F mixed kernels
        tfma = macs * n5 * n5 / (ms_v * 1e-3) / 1e12
        fft_route = max(int(bank["hmax"]), int(bank["wmax"])) >= dev.VALLEY_FFT_MIN_EXTENT
        if fft_route:
            # FFT route: per kernel pair and 2048^2 tile the product + inverse pass reads two spectra and writes one plane,
            # the transpose moves it once more, the fold reads it: 6 planes of 16 B x T^2
            T = 2048
            v_out = T - int(bank["hmax"]) + 1
            planes = (-(-ctx5.rows // v_out)) * (-(-n5 // (T - int(bank["wmax"]) + 1)))
            pairs = (len(bank["plain"]["hw"]) + 1) // 2
            gbs = pairs * planes * 6 * 16.0 * T * T / (ms_v * 1e-3) / 1e9
            v_roof = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak(), "unit": "GB/s", "frac": round(gbs / hbm_peak(), 3),
                      "note": f"2-D overlap-save FFT route: {pairs} kernel pairs x {planes} tiles of {T}^2 float64 complex, 6 plane "
                              f"transfers each; dense-equivalent bank rate {tfma:.1f} TFMA/s (fp32 FMA peak {128.0 * 148 * world * clk / 1e12:.1f})"}
        else:
            fma_peak = 128.0 * 148 * world * clk / 1e12
            v_roof = {"bound": "fp32 fma", "achieved": round(tfma, 2), "peak": round(fma_peak, 1),
                      "unit": "TFMA/s (dense-equivalent bank taps)", "frac": round(tfma / fma_peak, 3)}
        extra["config5"] = {
            "workload": f"config 5 kernels on a {n5}x{n5} {RES_M:g} m DEM in {world} row band(s) (the full 32768^2 runs through bench_c5.py): "
                        f"valley_ridge size {C5_KSIZE} (180 angles x 3 flats) and Sx radius {C5_RADIUS:g} m (window {plan5[3]} px, "
                        f"{int(plan5[2][-1])} samples)",
            "valley_ridge": {"ms": round(ms_v, 2), "value": round(n5 * n5 / (ms_v * 1e-3) / 1e6, 2), "unit": "Mpixel/s",
                             "route": "fft" if fft_route else "direct", "roofline": v_roof},
            "sx_10km": {"ms": round(ms_s, 2), "value": round(n5 * n5 / (ms_s * 1e-3) / 1e6, 1), "unit": "Mpixel/s"},
            "scaling": "strong",
        }
        del core5
        torch.cuda.empty_cache()

    # ---- FP64 pipe rate, measured now (8 DFMA chains per thread)
    scratch = torch.zeros(1, dtype=torch.float64, device=device)
    flops = ctypes.c_double(0.0)

    def probe():
        _lib.call("topo_probe_dfma", 20000, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(flops),
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

    probe()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    probe()
    p1.record()
    torch.cuda.synchronize()
    fp64_peak = flops.value / (p0.elapsed_time(p1) * 1e-3) / 1e12

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
    dom_name, dom_v = max(prof.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = dom_v["ms"] / dom_v["launches"]

    def alg_bytes_px(name):
        """Algorithmic HBM bytes per output pixel of one launch."""
        if name.startswith(("grad_from_smooth", "sobel_gradient", "gauss_grad_fused")):
            return 20  # 4 B read + 4 outputs
        if name.startswith("stats"):
            return 4
        return 8  # every other kernel on this path: 4 B in, 4 B out

    achieved = alg_bytes_px(dom_name) * band_px / (avg_ms * 1e-3) / 1e9
    ratio, ratio_src = ncu_dram_ratio(dom_name)
    traffic = None if ratio is None else int(ratio * alg_bytes_px(dom_name) * band_px)
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic,
        "traffic_source": (f"dram__bytes_read.sum + dram__bytes_write.sum of this kernel in {ratio_src} (ncu --set full), "
                           "scaled by pixels per launch") if traffic is not None else None,
        "peak_source": peak_src,
        "share_of_step": round(dom_v["ms"] / total_kernel_ms, 3),
        "note": "`achieved` counts the descriptor's algorithmic bytes (8 B/px) per launch of this kernel; the wide-radius work "
                "of config 4 runs as float64 FFT passes whose own traffic (tile spectra, 16 B per complex sample and pass) is "
                "several times that -- `memory_bound` lists every HBM-class kernel with the bytes it really moves "
                "(DESIGN.md section 4)",
        "fp64_pipe_peak_tflops": round(fp64_peak, 2),
        "fp64_pipe_peak_source": "topo_probe_dfma timed in this run with CUDA events (8 independent DFMA chains per thread)",
    }
    kernels = {
        k: {"launches": v["launches"], "ms": round(v["ms"], 3), "avg_ms": round(v["ms"] / v["launches"], 4),
            "max_ms": round(v["max_ms"], 3),
            "avg_GBps_alg": round(alg_bytes_px(k) * band_px / (v["ms"] / v["launches"] * 1e-3) / 1e9, 1)}
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    }
    # HBM-class kernels: algorithmic bytes (what the descriptor needs) and, for the helper passes of the FFT routes, the
    # bytes the pass itself moves per DEM pixel (tile spectra are tiles x T^2 x 16 B: `spec` bytes per pixel)
    tile_T, tile_H = 4096, max(sizes) // 2
    tile_V = tile_T - 2 * tile_H
    spec = (-(-ctx.rows // tile_V)) * (-(-nx // tile_V)) * tile_T * tile_T * 16.0 / band_px
    keep = tile_V / tile_T  # share of a first-pass line the second pass needs (the rest is neither stored nor transposed)
    # a plane that travels alone is laid out as twin tiles: those transforms move half the planes
    n_inv = max(1, kernels.get("disc_fft_inv", {}).get("launches", 1))
    n_twin = sum(v["launches"] for k, v in kernels.items() if k.startswith("disc_fft_finish") and ",twin>" in k)
    mix = (n_inv - 0.5 * min(n_twin, n_inv)) / n_inv
    moved = {"disc_fft_inv": (1 + keep) * spec * mix, "disc_fft_store": keep * spec + 16, "disc_fft_planes": 4 + spec * (0.75 if n_twin else 1.0),
             "disc_fft_finish<STD_F,twin>": keep * spec / 2 + 20, "disc_fft_finish<STD_I,twin>": keep * spec / 2 + 12,
             # (tpi of a tpi + std pair also leaves its two plane sums for the std: 16 B/px)
             "disc_fft_finish<TPI_X>": keep * spec + 20, "disc_fft_finish<STD_F>": keep * spec + 20,
             "disc_fft_finish<TPI_I>": keep * spec + 20, "disc_fft_finish<STD_I>": keep * spec + 12,
             "disc_finish<STD_F>": 28, "disc_finish<TPI_X>": 24, "disc_finish<STD_I>": 20, "disc_finish<TPI_I>": 16,
             "gauss_fft": 8, "transpose": 8}
    memory_bound = {}
    for k, v in kernels.items():
        if not k.startswith(("stats_partial", "grad_from_smooth", "sobel_gradient", "gauss_grad_fused", "disc_tiny", "disc_prefix",
                             "transpose", "disc_finish", "disc_fft_inv", "disc_fft_store", "disc_fft_finish", "disc_fft_planes",
                             "gauss_fft")) or \
                k.startswith(("gauss_fft_", )):
            continue
        e = {"GBps_alg": v["avg_GBps_alg"], "frac_of_hbm_peak": round(v["avg_GBps_alg"] / peak, 3)}
        if k in moved:
            gb = moved[k] * band_px / (v["avg_ms"] * 1e-3) / 1e9
            e.update({"moved_B_per_px": round(moved[k], 1), "GBps_moved": round(gb, 1), "frac_moved_of_hbm_peak": round(gb / peak, 3)})
        memory_bound[k] = e

    line = {
        "metric": METRIC,
        "value": round(value, 1), "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 in/out; u32 fixed-point + i64 sums (tpi/std), f64 accumulate (gaussian)", "data": "synthetic",
        "config": dict(workload_config(ny, nx, sizes, "float"),
                       parallelism=f"row bands x{world}, halo exchange over NVLink", numa_bound=bool(numa_bound),
                       launch="CUDA graph replay of the sweep's kernel sequence" if used_graph else (graph_note or "eager")),
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": e2e_h2d,
                "d2h_bytes_per_step": e2e_d2h, "steps": e2e_steps, "band_rows": e2e_rows, "d2h_GBps_all_ranks_at_once": d2h_rates,
                "path": "pinned host DEM -> HBM -> bands.sweep -> every output band back to pinned host memory (D2H on a second stream, overlapped)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "memory_bound": memory_bound,
        "kernels": kernels,
        "extra": extra,
    }
    if band_check is not None:
        line["band_check"] = band_check
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, n_calls))
        kind = cpu_kind()
        mpix, sec, per_call = cpu_sweep(sizes, CPU_CROP, workers, integer=False)
        what = "unmodified reference package" if kind == "reference" else "oracle port of the reference call sequence"
        line["cpu_baseline"] = {
            "value": round(mpix, 3), "unit": "Mpixel/s", "cores": workers, "kind": kind,
            "sample": f"same sweep ({what}) on a {CPU_CROP}x{CPU_CROP} crop of the float DEM, {sec:.1f} s, "
                      f"{workers} worker processes of {cores} cores",
        }
        if not args.no_extra:
            m3, s3 = cpu_sx()
            extra["config3_sx"]["cpu_baseline"] = {
                "value": round(m3, 3), "unit": "Mpixel/s per azimuth call", "cores": cores, "kind": kind,
                "sample": f"{what}: topo.sx for 4 azimuths on a 1024^2 crop, {s3:.1f} s (numba prange, JIT call excluded)"}
            m5, s5 = cpu_valley()
            extra["config5"]["valley_ridge"]["cpu_baseline"] = {
                "value": round(m5, 4), "unit": "Mpixel/s", "cores": 1, "kind": kind,
                "sample": f"{what}: valley_ridge size {C5_KSIZE} on a 320^2 crop, {s5:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(path))["hbm_gbs"]) if os.path.exists(path) else 6650.0


def band_spot_check(core, ctx, sizes, sigmas, res_x, res_y, rank, device):
    """Rows straddling the boundary between ranks 0 and 1, computed (a) by the two ranks as part of their bands (halo
    exchange over NCCL) and (b) by rank 0 alone from a band it builds directly from the DEM generator, in global
    coordinates.  Bit-for-bit equality of tpi / std at the largest size and of the gradient at a small sigma."""
    import torch
    import torch.distributed as dist

    from topo_descriptors_b200 import bands, device as dev
    from topo_descriptors_b200.device import DeviceDEM

    H = 32
    edge = ctx.parts[0][1]
    size, si = sizes[-1], len(sizes) - 1
    gi = 2
    kept = {}

    def sink(name, i, t):
        if (name in ("tpi", "std") and i == si) or (name in ("dx", "aspect") and i == gi):
            kept[name] = t[-H:].clone() if rank == 0 else t[:H].clone()

    stats = bands.global_stats(dev.dem_stats(core), ctx, device=device)
    bands.sweep(core, ctx, sizes, sigmas, res_x, res_y, sink=sink, stats=stats)
    names = ("tpi", "std", "dx", "aspect")
    if rank == 1:
        for n in names:
            dist.send(kept[n].contiguous(), 0)
    ok = None
    if rank == 0:
        got = {}
        for n in names:
            other = torch.empty_like(kept[n])
            dist.recv(other, 1)
            got[n] = torch.cat([kept[n], other], dim=0)
        halo = max(size // 2, dev.gauss_radius(sigmas[gi]) + 1)
        a, b = max(0, edge - H - halo), min(ctx.gny, edge + H + halo)
        band = torch.from_numpy(make_dem_rows(ctx.gny, ctx.nx, a, b, integer=False)).to(device)
        d = DeviceDEM(band, gny=ctx.gny, gy0=a, stats=stats).share_disc_planes(size)
        want = {"tpi": dev.tpi(d, size, edge - H, 2 * H, pair_std=True), "std": dev.std(d, size, edge - H, 2 * H)}
        d.release_disc_planes()
        g0 = edge - H - 1
        g = DeviceDEM(dev.gauss(d, sigmas[gi], sigmas[gi], g0, 2 * H + 2), gny=ctx.gny, gy0=g0, stats=stats)
        outs = dev.gradient_from_smooth(g, g, res_x[0], 0, res_y[0], 0, edge - H, 2 * H)
        want["dx"], want["aspect"] = outs[0], outs[3]
        bad = [n for n in names if not torch.equal(got[n], want[n])]
        ok = {"rows": [edge - H, edge + H], "descriptors": list(names), "bit_identical": not bad, "mismatches": bad}
    dist.barrier()
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="DEM edge in pixels (default: config 4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the integer-DEM / config 3 / config 5 extras")
    ap.add_argument("--no-graph", action="store_true", help="launch the sweep eagerly instead of replaying a CUDA graph")
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
