"""ncu driver: one valley_ridge launch (size 41, 180 angles x 3 flats) on a PROF_SIZE^2 crop (default 1024)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from topo_descriptors_b200 import device as dev, topo  # noqa: E402
from topo_descriptors_b200.device import DeviceDEM  # noqa: E402
from topo_descriptors_b200.synth import fractal_dem  # noqa: E402

n = int(os.environ.get("PROF_SIZE", 1024))
size = int(os.environ.get("PROF_KSIZE", 41))
z = fractal_dem(n, n, seed=3)
d = DeviceDEM(dev.to_device(z))
st = d.stats
mean = st["sum"] / st["n"]
sd = np.sqrt(max(st["sumsq"] / st["n"] - mean * mean, 0.0))
normed = dev.zscore(d, np.float32(mean), np.float32(sd))
bank = topo._device_bank(size, "valley", [0, 0.15, 0.3], d.tensor.device)
for rep in range(int(os.environ.get("PROF_REPS", 1))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.valley_ridge(normed, bank)
    e1.record()
    torch.cuda.synchronize()
    print(f"valley_ridge size {size} on {n}^2: {e0.elapsed_time(e1):.2f} ms", flush=True)
