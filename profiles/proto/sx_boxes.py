"""Round-2 preparation: can the 10 km Sx scan (window 400 px, ~5400 unique samples per sector) run from TMA-staged
shared memory like the 500 m one?  The samples of a sector are split into groups (recursive bisection along the longer extent) so that the tile
(128 x 16 outputs) plus the group's bounding box fits one TMA box (<= 256 x 256 elements, smem budget); the kernel
would loop over the groups (one box each, double buffered).  This script plans the groups for a few azimuths and
prints the box sizes, the shared memory and the L2 -> smem traffic amplification.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from topo_descriptors_b200 import _xr, topo  # noqa: E402

TW, TH = 128, 16


def box_of(g):
    dx0 = int(g[:, 1].min()) & ~3  # the box origin's x must be a multiple of 4 elements (tma_probe)
    w = ((TW + int(g[:, 1].max()) - dx0) + 31) & ~31
    h = TH + int(g[:, 0].max()) - int(g[:, 0].min())
    return w, h


def plan_groups(offsets, budget_bytes=100 * 1024):
    """Recursive bisection of the sample set along its longer extent until tile + bounding box fits one TMA box
    (<= 256 x 256 elements) and the shared-memory budget."""
    out = []

    def rec(g):
        w, h = box_of(g)
        if w <= 256 and h <= 256 and w * h * 4 <= budget_bytes or len(g) == 1:
            out.append((len(g), w, h, w * h * 4))
            return
        axis = 0 if (g[:, 0].max() - g[:, 0].min()) >= (g[:, 1].max() - g[:, 1].min()) else 1
        mid = (int(g[:, axis].min()) + int(g[:, axis].max())) // 2
        lo, hi = g[g[:, axis] <= mid], g[g[:, axis] > mid]
        rec(lo)
        rec(hi)

    rec(np.asarray(offsets))
    return out


if __name__ == "__main__":
    n = 4096
    x = 2600000.0 + 25.0 * np.arange(n)
    y = 1200000.0 - 25.0 * np.arange(n)
    grid = _xr.Dataset({"alti": (("y", "x"), np.zeros((1, 1), np.float32))}, coords={"x": x, "y": y}, attrs={"crs": "epsg:2056"})
    for az in (270.0, 225.0, 200.0, 0.0):
        offsets, inv, begin, window = topo._sx_plan(grid, [az], 10000.0, 10.0, 15, 0.0)
        groups = plan_groups(np.asarray(offsets))
        smem = max(g[3] for g in groups)
        traffic = sum(g[3] for g in groups)
        print(f"azimuth {az:5.1f}: {len(offsets)} samples, window {window}: {len(groups)} boxes, largest {smem / 1024:.0f} KB, "
              f"{traffic / 1024:.0f} KB staged per 128 x 16 tile = {traffic / (TW * TH * 4):.0f}x the tile "
              f"({traffic / (TW * TH) * n * n / 1e9:.1f} GB of L2 reads for a {n}^2 DEM)")
        print("   boxes (samples, w, h):", [(g[0], g[1], g[2]) for g in groups])
