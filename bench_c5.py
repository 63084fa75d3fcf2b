#!/usr/bin/env python
"""Config 5 of BASELINE.json: continental DEM, valley_ridge kernel bank + Sx radius 10 km, row-band sharded
(bench.py carries the headline, config 4; this is the secondary multi-GPU measurement).

    python bench_c5.py --size 8192                                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        bench_c5.py --size 32768                                         # 8 GPUs, one rank per GPU

Every rank owns ny/N rows (+ halo rows received from its neighbours over NVLink inside the timed region), the
z-score statistics are all-reduced, and the same kernels run with the band's topo_view.  Timing: CUDA events
bracketed by a barrier, max over ranks; one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768, help="DEM edge in pixels")
    ap.add_argument("--ksize", type=int, default=41, help="valley kernel size in pixels (1 km at 25 m)")
    ap.add_argument("--radius", type=float, default=10000.0, help="Sx radius in metres")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bench import RES_M, make_dem_rows
    from topo_descriptors_b200 import _lib, _xr, bands, topo

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    n = args.size
    ctx = bands.BandContext(n, n, rank, world)
    core = torch.from_numpy(make_dem_rows(n, n, ctx.r0, ctx.r1)).to(device)
    x = 2600000.0 + RES_M * np.arange(n, dtype=np.float64)
    y = 1200000.0 - RES_M * np.arange(n, dtype=np.float64)
    grid = _xr.Dataset({"alti": (("y", "x"), np.zeros((1, 1), np.float32))}, coords={"x": x, "y": y},
                       attrs={"crs": "epsg:2056"})  # only the coordinates are used by the Sx geometry
    plan = topo._sx_plan(grid, [270.0], args.radius, 10.0, 15, 0.0)
    flats = [0, 0.15, 0.3]
    bank = topo._device_bank(args.ksize, "valley", flats, device)  # host-side bank build, cached: outside the timing

    def timed(fn):
        best = None
        for _ in range(args.reps + 1):  # first pass = warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = t.item() if best is None else min(best, t.item())
        return best

    ms_valley = timed(lambda: bands.valley_ridge_band(core, ctx, args.ksize, "valley", flats))
    ms_sx = timed(lambda: bands.sx_band(core, ctx, plan, 10.0))
    if rank == 0:
        px = float(n) * n
        macs = float(sum(int(h) * int(w) for h, w, _a in bank["plain"]["hw"]))
        print(json.dumps({
            "config": f"config 5: {n}x{n} 25 m DEM, valley_ridge size {args.ksize} (180 angles x 3 flats) + Sx radius "
                      f"{args.radius:.0f} m (window {plan[3]} px, {int(plan[2][-1])} samples), row bands x{world}",
            "n_gpus": world, "reps": args.reps, "scaling": "strong", "data": "synthetic",
            "valley_ridge": {"ms": round(ms_valley, 2), "mpix_s": round(px / ms_valley / 1e3, 1),
                             "dense_equivalent_tfma_s": round(macs * px / ms_valley / 1e9, 1)},
            "sx": {"ms": round(ms_sx, 2), "mpix_s": round(px / ms_sx / 1e3, 1)},
            "timing": "CUDA events after a barrier, max over ranks, best of reps; halo exchange + z-score all-reduce inside",
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
