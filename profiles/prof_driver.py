"""Small driver for ncu captures: one call of selected descriptors on a 16384^2 DEM (config 4).

    python profiles/prof_driver.py tpi:801 std:801 grad:801 tpi:5 grad:5
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import RES_M, make_dem_rows  # noqa: E402
from topo_descriptors_b200 import device as dev  # noqa: E402
from topo_descriptors_b200.device import DeviceDEM  # noqa: E402

n = int(os.environ.get("PROF_SIZE", 16384))
core = torch.from_numpy(make_dem_rows(n, n, 0, n)).cuda()
d = DeviceDEM(core)
_ = d.stats
rx = torch.full((n,), RES_M, dtype=torch.float64, device="cuda")
ry = torch.full((n,), -RES_M, dtype=torch.float64, device="cuda")
for spec in sys.argv[1:]:
    kind, size = spec.split(":")
    size = int(size)
    if kind == "tpi":
        dev.tpi(d, size)
    elif kind == "std":
        dev.std(d, size)
    elif kind == "gauss":
        dev.gauss(d, size / 4.0, size / 4.0)
    elif kind == "grad":
        g = DeviceDEM(dev.gauss(d, size / 4.0, size / 4.0))
        dev.gradient_from_smooth(g, g, rx, 0, ry, 0)
    torch.cuda.synchronize()
print("done")
