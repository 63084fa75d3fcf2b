#!/bin/bash
# packed-word tiny kernel for the float std of sizes 5..13: tests + timing
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tiny or std or tpi or edge or compute" > $O/r02_pytest31.log 2>&1; tail -4 $O/r02_pytest31.log
PROF_TIME=1 PROF_FLOAT=1 python profiles/prof_driver.py std:5 std:9 std:13 tpi:13 2>&1 | tail -5
PROF_OFF=tiny PROF_TIME=1 PROF_FLOAT=1 python profiles/prof_driver.py std:5 std:9 std:13 2>&1 | tail -4
