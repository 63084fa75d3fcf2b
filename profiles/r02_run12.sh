#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gauss or gradient or valley or fft" > $O/r02_pytest12.log 2>&1; tail -3 $O/r02_pytest12.log
PROF_TIME=1 PROF_FLOAT=1 python profiles/prof_driver.py gauss:81 gauss:161 gauss:241 gauss:401 gauss:801 > $O/r02_prof12.log 2>&1
cat $O/r02_prof12.log
timeout 300 python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from topo_descriptors_b200 import device as dev, topo
from topo_descriptors_b200.device import DeviceDEM
from topo_descriptors_b200.synth import fractal_dem
z = fractal_dem(2048, 2048, seed=3)
d = DeviceDEM(dev.to_device(z)); st = d.stats
mean = st["sum"]/st["n"]; sd = np.sqrt(st["sumsq"]/st["n"]-mean*mean)
normed = dev.zscore(d, np.float32(mean), np.float32(sd))
bank = topo._device_bank(41, "valley", [0, 0.15, 0.3], d.tensor.device)
dev.valley_ridge(normed, bank); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dev.valley_ridge(normed, bank); e1.record(); torch.cuda.synchronize()
print(f"valley 41 fft: {e0.elapsed_time(e1):.1f} ms per 2048^2")
PY
