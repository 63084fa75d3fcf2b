#!/usr/bin/env python
"""Secondary measurements for the other BASELINE configs (bench.py carries the headline, config 4):

  C1  topo.tpi 500 m (size 17, and 35) on the 900x1440 DEM
  C2  gradient (sigma 1.75 / 16.75), std / tpi (sizes 7 / 67) on the same DEM
  C3  topo.sx radius 500 m, azimuths 0..355 step 5 on a 4096^2 DEM (one launch for all 72 sectors)
  C5  valley_ridge size 41 and sx radius 10 km (window 400 px) -- on a crop, the kernels are local

Device time by CUDA events around the device-level calls (inputs resident), median of `--reps`.
Small rasters are L2-resident: those numbers are time-per-call, not HBM roofline material.
With --cpu the reference's algorithm (oracle/*_literal, numba-free) is timed next to it where it is quick.

    python bench_extra.py [--reps 10] [--cpu] > profiles/rNN_extra.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--c5-size", type=int, default=4096)
    args = ap.parse_args()

    import torch

    from topo_descriptors_b200 import _lib, device as dev, helpers as hlp, topo
    from topo_descriptors_b200.device import DeviceDEM
    from topo_descriptors_b200.synth import dem_dataset, fractal_dem

    torch.cuda.set_device(0)
    _lib.load()

    def timed(fn, reps=args.reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {"gpu": torch.cuda.get_device_name(0), "reps": args.reps, "results": []}

    def add(cfg, name, px, ms, cpu_ms=None, note=None):
        r = {"config": cfg, "call": name, "ms": round(ms, 4), "mpix_s": round(px / ms / 1e3, 1)}
        if cpu_ms is not None:
            r["cpu_ms"] = round(cpu_ms, 2)
            r["cpu_mpix_s"] = round(px / cpu_ms / 1e3, 2)
        if note:
            r["note"] = note
        out["results"].append(r)
        print(json.dumps(r), file=sys.stderr, flush=True)

    def cpu_time(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t) * 1e3)
        return float(np.median(ts))

    from oracle import oracle as O

    # ---- C1 / C2 -------------------------------------------------------------------------------
    z = fractal_dem(900, 1440, seed=0)
    zi = np.rint(z).astype(np.float32)
    ds = dem_dataset(z, res=30.0)
    d, di = DeviceDEM(dev.to_device(z)), DeviceDEM(dev.to_device(zi))
    _ = d.stats, di.stats
    px = 900 * 1440
    _, res = hlp.scale_to_pixel([200], ds)
    rx, rx2 = dev._res_to_device(res["x"], d.tensor.device)
    ry, ry2 = dev._res_to_device(res["y"], d.tensor.device)
    for size in (17, 35, 7, 67):
        cfg = "C1" if size in (17, 35) else "C2"
        add(cfg, f"tpi size {size}", px, timed(lambda: dev.tpi(d, size)),
            cpu_time(lambda: O.tpi_literal(z, size)) if args.cpu else None)
    for size in (7, 67):
        add("C2", f"std size {size} (integer DEM)", px, timed(lambda: dev.std(di, size)),
            cpu_time(lambda: O.std_literal(zi, size)) if args.cpu else None)
        add("C2", f"std size {size} (float DEM)", px, timed(lambda: dev.std(d, size)))
    for sigma in (0.75, 1.75, 16.75):
        def grad():
            if sigma <= 1:
                return dev.sobel_gradient(d, rx, rx2, ry, ry2)
            g = DeviceDEM(dev.gauss(d, sigma, sigma))
            return dev.gradient_from_smooth(g, g, rx, rx2, ry, ry2)
        add("C2", f"gradient sigma {sigma}", px, timed(grad),
            cpu_time(lambda: O.gradient_literal(z, sigma, res)) if args.cpu else None)
    # end-to-end through the public API (host array in, host arrays out)
    t = cpu_time(lambda: topo.tpi(z, 17), reps=5)
    add("C1", "topo.tpi(z, 17) end to end incl. H2D/D2H (wall)", px, t)

    # ---- C3: Sx radius 500 m, 72 azimuths, 4096^2 -----------------------------------------------------
    n3 = 4096
    z3 = fractal_dem(n3, n3, seed=1)
    ds3 = dem_dataset(z3, res=30.0)
    d3 = DeviceDEM(dev.to_device(z3))
    azs = list(range(0, 360, 5))
    plan = topo._sx_plan(ds3, azs, 500.0, 10.0, 15, 0.0)
    n_samples = int(plan[2][-1])
    ms = timed(lambda: topo._sx_device(d3, plan, 10.0), reps=max(3, args.reps // 2))
    add("C3", f"sx radius 500 m x {len(azs)} azimuths (one launch, {n_samples} unique samples total)", n3 * n3 * len(azs), ms,
        note="Mpix/s counts one azimuth = one descriptor call")
    plan1 = topo._sx_plan(ds3, [225.0], 500.0, 10.0, 15, 0.0)
    add("C3", "sx radius 500 m, single azimuth", n3 * n3, timed(lambda: topo._sx_device(d3, plan1, 10.0)))

    # ---- C5 (crop): valley_ridge size 41, sx radius 10 km ---------------------------------------------------
    n5 = args.c5_size
    z5 = fractal_dem(min(n5, 4096), min(n5, 4096), seed=3)
    if n5 > 4096:
        z5 = np.tile(z5, (n5 // 4096, n5 // 4096))
    d5 = DeviceDEM(dev.to_device(z5))
    st = d5.stats
    mean = st["sum"] / st["n"]
    sd = np.sqrt(max(st["sumsq"] / st["n"] - mean * mean, 0.0))
    normed = dev.zscore(d5, np.float32(mean), np.float32(sd))
    for size in (7, 41):
        t0 = time.perf_counter()
        bank = topo._device_bank(size, "valley", [0, 0.15, 0.3], d5.tensor.device)
        bank_s = time.perf_counter() - t0
        ms = timed(lambda: dev.valley_ridge(normed, bank), reps=3)
        macs = float(sum(int(h) * int(w) for h, w, _a in bank["plain"]["hw"]))
        add("C5", f"valley_ridge size {size}, 180 angles x 3 flats, {n5}^2 crop", n5 * n5, ms,
            note=f"{macs:.0f} MAC/px -> {macs * n5 * n5 / ms / 1e9:.2f} TFMA/s fp32; bank build {bank_s:.2f} s (host, cached)")
    ds5 = dem_dataset(z5, res=25.0)
    plan5 = topo._sx_plan(ds5, [270.0], 10000.0, 10.0, 15, 0.0)
    add("C5", f"sx radius 10 km (window {plan5[3]} px, {int(plan5[2][-1])} unique samples), {n5}^2 crop", n5 * n5,
        timed(lambda: topo._sx_device(d5, plan5, 10.0), reps=3))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
