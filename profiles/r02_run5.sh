#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q -k "valley or rotated" > $O/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest5.log
tail -25 $O/r02_pytest5.log
timeout 600 python - > $O/r02_valley_sizes.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from topo_descriptors_b200 import device as dev, topo
from topo_descriptors_b200.device import DeviceDEM
from topo_descriptors_b200.synth import fractal_dem
z = fractal_dem(2048, 2048, seed=3)
d = DeviceDEM(dev.to_device(z)); st = d.stats
mean = st["sum"]/st["n"]; sd = np.sqrt(st["sumsq"]/st["n"]-mean*mean)
normed = dev.zscore(d, np.float32(mean), np.float32(sd))
for size in (21, 41, 61, 81, 161, 401):
    for route, thr in (("fft", 1), ("direct", 10**6)):
        if route == "direct" and size > 81: continue
        dev.VALLEY_FFT_MIN_EXTENT = thr
        topo._BANK_CACHE.clear()
        t0 = time.perf_counter(); bank = topo._device_bank(size, "valley", [0, 0.15, 0.3], d.tensor.device); torch.cuda.synchronize(); tb = time.perf_counter() - t0
        dev.valley_ridge(normed, bank); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dev.valley_ridge(normed, bank); e1.record(); torch.cuda.synchronize()
        print(f"valley size {size} {route}: {e0.elapsed_time(e1):.1f} ms per 2048^2 (hmax {bank['hmax']}, bank build {tb:.2f} s)", flush=True)
PY
cat $O/r02_valley_sizes.log
