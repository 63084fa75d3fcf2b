#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tpi or std or disc or bands or cached" > $O/r02_pytest9.log 2>&1; tail -3 $O/r02_pytest9.log
PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py std:41 std:81 std:161 std:241 std:401 std:801 > $O/r02_prof9.log 2>&1
cat $O/r02_prof9.log
