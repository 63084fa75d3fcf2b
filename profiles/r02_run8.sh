#!/bin/bash
O=gpurun_out
PROF_SIZE=8192 PROF_FLOAT=1 PROF_SHARE=801 ncu --set full --clock-control none -k regex:disc_span -c 8 -f -o /tmp/r02_disc python profiles/prof_driver.py std:41 std:161 std:801 > $O/r02_prof8.log 2>&1
tail -2 $O/r02_prof8.log
python profiles/ncu_summary.py /tmp/r02_disc.ncu-rep > $O/r02_ncu_disc_summary.csv
ncu -i /tmp/r02_disc.ncu-rep --page raw --csv > $O/r02_disc_raw.csv 2>/dev/null
