#!/bin/bash
# rectangular FFT windows (3072-point transform along y for thin bands): full GPU suite + band timing on one GPU
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02_pytest35.log 2>&1; tail -8 $O/r02_pytest35.log
python - <<'PY'
# one rank's band of an 8- and a 4-GPU run, emulated on one GPU: sweep time of the band
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from bench import make_dem_rows, sizes_for, SCALES_M, RES_M
from topo_descriptors_b200 import bands, device as dev, _lib
from topo_descriptors_b200.device import DeviceDEM
n = 16384
sizes = sizes_for(SCALES_M, RES_M); sigmas = [s / 4.0 for s in sizes]
dvc = torch.device('cuda', 0)
res_x = (dev._Res(np.full(n, RES_M), dvc), 0); res_y = (dev._Res(np.full(n, -RES_M), dvc), 0)
for world, rank in ((8, 3), (4, 1)):
    ctx = bands.BandContext(n, n, rank, world)
    halo = bands.sweep_halo(sizes, sigmas)
    a, b = ctx.halo_extent(halo)
    band = torch.from_numpy(make_dem_rows(n, n, a, b, integer=False)).to(dvc)
    whole_stats = {"min": 200.0, "max": 3400.0, "nonfinite": 0, "nonint": 1, "sum": 0.0, "sumsq": 0.0, "n": n * n}
    def step():
        d = DeviceDEM(band, gny=n, gy0=a, stats=whole_stats)
        bands._sweep_body(d, ctx, sizes, sigmas, res_x, res_y, ("tpi", "std", "gradient"), None)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    print(f"band of rank {rank}/{world}: rows {ctx.rows}, {e0.elapsed_time(e1)/3:.2f} ms per sweep (eager)")
PY
