#!/bin/bash
# group-level barriers between FFT stages: timing + FFT-route tests (disc, valley, Gaussian)
O=gpurun_out
PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 grad:801 grad:161 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q -k "fft or valley_ridge or cached_sweep or 401_801 or next_size or sweep_graph or gaussian or wide_radii" > $O/r02_pytest24.log 2>&1; tail -4 $O/r02_pytest24.log
