#!/bin/bash
# ncu --set full of the FFT disc route's kernels (8192^2 float DEM, sweep mode) -> condensed CSV
O=gpurun_out
PROF_SIZE=8192 PROF_FLOAT=1 PROF_SHARE=801 ncu --set full --clock-control none --import-source on -k regex:"dfft|fft2d" -c 14 -f -o /tmp/r02_prof_dfft \
    python profiles/prof_driver.py tpi:801 std:801 > $O/r02_prof_dfft.log 2>&1
python profiles/ncu_summary.py /tmp/r02_prof_dfft.ncu-rep > $O/r02_ncu_dfft_summary.csv
ncu -i /tmp/r02_prof_dfft.ncu-rep --page details --csv > $O/r02_ncu_dfft_details.csv 2>/dev/null
ls -la /tmp/r02_prof_dfft.ncu-rep; cp /tmp/r02_prof_dfft.ncu-rep $O/ 2>/dev/null; du -sh $O
cat $O/r02_ncu_dfft_summary.csv | cut -c1-400
