#!/bin/bash
# A/B: transposed store of the first inverse pass (fft_tstore) vs the transpose pass; bit-identity tests of both FFT routes
O=gpurun_out
export PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801
echo "--- tstore on"; python profiles/prof_driver.py tpi:801 std:801 tpi:161 std:161 2>&1 | tail -5
echo "--- tstore off"; PROF_OFF=fft_tstore python profiles/prof_driver.py tpi:801 std:801 2>&1 | tail -3
unset PROF_TIME PROF_FLOAT PROF_SHARE
timeout 900 python -m pytest tests -m gpu -x -q -k "disc_fft or valley_ridge or cached_sweep or 401_801" > $O/r02_pytest16.log 2>&1; tail -4 $O/r02_pytest16.log
