#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q -k "fused" > $O/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest3.log
tail -3 $O/r02_pytest3.log
PROF_TIME=1 PROF_FLOAT=1 python profiles/prof_driver.py grad:5 grad:9 grad:13 grad:21 std:5 std:9 std:13 std:21 tpi:21 > $O/r02_prof3.log 2>&1
PROF_TIME=1 PROF_FLOAT=1 PROF_SHARE=801 python profiles/prof_driver.py std:41 std:81 std:161 std:241 std:401 std:801 tpi:801 >> $O/r02_prof3.log 2>&1
cat $O/r02_prof3.log
