"""Small FFT-route workload for compute-sanitizer (memcheck / racecheck): disc route (single call, cached sweep with the
std look-ahead), Gaussian FFT pass, valley/ridge FFT route -- sizes chosen so that every kernel runs on one or two tiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from topo_descriptors_b200 import device as dev, topo  # noqa: E402
from topo_descriptors_b200.device import DeviceDEM  # noqa: E402
from topo_descriptors_b200.synth import fractal_dem  # noqa: E402

z = fractal_dem(300, 360, seed=3)
d = DeviceDEM(dev.to_device(z))
a = dev.tpi(d, 129).cpu().numpy()
s = DeviceDEM(dev.to_device(z)).share_disc_planes(161)
outs = []
for k, size in enumerate((161, 65, 33)):
    outs.append(dev.tpi(s, size, pair_std=True).cpu().numpy())
    outs.append(dev.std(s, size, next_size=(65, 33, 0)[k]).cpu().numpy())
g = dev.gauss(d, 12.0, 12.0).cpu().numpy()
v = topo.valley_ridge(z[:200, :220], 61, "valley")
torch.cuda.synchronize()
print("ok", float(np.abs(a).max()), float(np.abs(outs[1]).max()), float(g.mean()), float(np.asarray(v[0]).max()))
