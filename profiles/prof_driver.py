"""Small driver for ncu captures and quick timings: calls of selected descriptors on a 16384^2 DEM (config 4).

    python profiles/prof_driver.py tpi:801 std:801 grad:801 tpi:5 grad:5          # one call each (for ncu)
    PROF_TIME=1 python profiles/prof_driver.py tpi:801 std:801                      # 3 calls each, prints the last ms + its kernels
    PROF_OFF=gauss_fft,grad_fused ...                                               # execution-shape switches to turn off
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
timing = bool(os.environ.get("PROF_TIME"))
core = torch.from_numpy(make_dem_rows(n, n, 0, n)).cuda()
if os.environ.get("PROF_FLOAT"):
    core = core + 0.25
d = DeviceDEM(core)
_ = d.stats
if os.environ.get("PROF_SHARE"):  # multi-scale sweep mode: plane cache + octagon walk
    d.share_disc_planes(int(os.environ["PROF_SHARE"]))
rx = torch.full((n,), RES_M, dtype=torch.float64, device="cuda")
ry = torch.full((n,), -RES_M, dtype=torch.float64, device="cuda")


def run_one(kind, size):
    if kind == "tpi":
        dev.tpi(d, size)
    elif kind == "std":
        dev.std(d, size)
    elif kind == "gauss":
        dev.gauss(d, size / 4.0, size / 4.0)
    elif kind == "grad":  # the public route: fused kernel for small radii, else smoothing + differences
        dev.gradient(d, size / 4.0, rx, 0, ry, 0)
    elif kind == "grad3":  # always the three-kernel route
        g = DeviceDEM(dev.gauss(d, size / 4.0, size / 4.0))
        dev.gradient_from_smooth(g, g, rx, 0, ry, 0)
    elif kind == "sobel":
        dev.sobel_gradient(d, rx, 0, ry, 0)


from topo_descriptors_b200 import _lib  # noqa: E402

for opt in os.environ.get("PROF_OFF", "").split(","):  # e.g. PROF_OFF=gauss_fft,grad_fused
    if opt:
        _lib.set_option(opt, False)
for spec in sys.argv[1:]:
    kind, size = spec.split(":")
    size = int(size)
    reps = 3 if timing else 1
    for rep in range(reps):
        if timing and rep == reps - 1:
            _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_one(kind, size)
        e1.record()
        torch.cuda.synchronize()
        if timing and rep == reps - 1:
            per = ", ".join(f"{k} {ms:.3f}" for k, ms in _lib.profile_dump(aggregate=False))
            _lib.profile_enable(False)
            print(f"{spec}: {e0.elapsed_time(e1):.3f} ms  [{per}]", flush=True)
print("done")
