#!/bin/bash
# mask line by eights (product inside the first inverse stage) + 4 CTAs/SM for the 2048-point Gaussian FFT
O=gpurun_out
PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 grad:161 grad:401 2>&1 | tail -5
PROF_SIZE=8192 PROF_TIME=1 python profiles/prof_valley.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "fft or valley_ridge or cached_sweep or 401_801 or next_size or sweep_graph or gaussian or wide_radii or stats or nan or edge" > $O/r02_pytest26.log 2>&1; tail -4 $O/r02_pytest26.log
