#!/bin/bash
# grouped epilogue of the FFT store pass: timing + FFT-route tests
O=gpurun_out
PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 2>&1 | tail -3
PROF_TIME=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q -k "disc_fft or cached_sweep or 401_801 or sweep_graph or wide_range or tiler or edge" > $O/r02_pytest30.log 2>&1; tail -4 $O/r02_pytest30.log
