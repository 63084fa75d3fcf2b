#!/bin/bash
# Round-2 measurement + profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/, which is
# limited to 64 MiB: the big ncu reports are condensed to CSV on the box and dropped).
set -x
O=gpurun_out
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r02_ncu_launches.log 2>&1
# full-set captures, one call per descriptor class in sweep mode on an 8192^2 float DEM (plane cache + octagon walk,
# FFT Gaussian, fused gradient), condensed to CSV
PROF_SIZE=8192 PROF_FLOAT=1 PROF_SHARE=801 ncu --set full --clock-control none -c 80 -f -o /tmp/r02_prof \
    python profiles/prof_driver.py tpi:801 std:801 grad:801 grad:161 tpi:21 tpi:5 grad:5 grad:21 sobel:0 > $O/r02_prof.log 2>&1
python profiles/ncu_summary.py /tmp/r02_prof.ncu-rep > $O/r02_ncu_full_summary.csv
# valley/ridge FFT route (size 41, 2048^2) and Sx (config 3 through bench_extra)
PROF_SIZE=2048 ncu --set full --clock-control none -k regex:"vfft|rotate" -c 12 -f -o /tmp/r02_prof_valley \
    python profiles/prof_valley.py > $O/r02_prof_valley.log 2>&1
python profiles/ncu_summary.py /tmp/r02_prof_valley.ncu-rep > $O/r02_ncu_valley_summary.csv
python bench_extra.py --reps 6 > $O/r02_extra.json 2> $O/r02_extra.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
du -sh $O; ls -la $O | tail -20
