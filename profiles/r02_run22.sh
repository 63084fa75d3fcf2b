#!/bin/bash
# padded pitch of the transposed planes; tstore on / off with the async-staged passes
O=gpurun_out
echo "--- tstore on (padded pitch)"; PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 2>&1 | tail -3
echo "--- tstore off"; PROF_OFF=fft_tstore PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py tpi:801 std:801 2>&1 | tail -3
PROF_SIZE=8192 PROF_TIME=1 python profiles/prof_valley.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "disc_fft or valley_ridge or cached_sweep or 401_801 or next_size or sweep_graph" > $O/r02_pytest22.log 2>&1; tail -4 $O/r02_pytest22.log
