#!/bin/bash
O=gpurun_out
PROF_SIZE=8192 PROF_FLOAT=1 ncu --set full --import-source on --clock-control none -k regex:gauss_grad_fused -c 2 -f -o /tmp/r02_fused python profiles/prof_driver.py grad:5 grad:9 > $O/r02_prof4.log 2>&1
tail -3 $O/r02_prof4.log
ncu -i /tmp/r02_fused.ncu-rep --page raw --csv > $O/r02_fused_raw.csv 2>/dev/null
ncu -i /tmp/r02_fused.ncu-rep --page source --csv --print-source sass > $O/r02_fused_source.csv 2>/dev/null
ls -la $O/r02_fused_source.csv
