#!/bin/bash
# Round-2 measurement + profile capture, second pass (after the FFT disc route, the async-staged inverse passes and the
# std look-ahead).  Run on the GPU box through gpurun; outputs land in gpurun_out/ (64 MiB limit: the ncu reports are
# condensed to CSV on the box and dropped).
set -x
O=gpurun_out
python bench.py > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02b_bench_reference.json 2> $O/r02b_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r02b_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r02b_ncu_launches.log 2>&1
# full-set captures, one call per descriptor class in sweep mode on an 8192^2 float DEM (plane spectra cached, FFT disc
# route, FFT Gaussian, fused gradient, fused / tiny small discs), condensed to CSV
PROF_SIZE=8192 PROF_FLOAT=1 PROF_SHARE=801 ncu --set full --clock-control none -c 110 -f -o /tmp/r02b_prof \
    python profiles/prof_driver.py tpi:801 std:801 tpi:401 std:401 grad:801 grad:161 grad:5 grad:21 sobel:0 tpi:21 std:21 tpi:5 std:5 std:13 > $O/r02b_prof.log 2>&1
python profiles/ncu_summary.py /tmp/r02b_prof.ncu-rep > $O/r02b_ncu_full_summary.csv
# valley/ridge FFT route (size 41, 2048^2)
PROF_SIZE=2048 ncu --set full --clock-control none -k regex:"vfft|rotate|fft2d" -c 14 -f -o /tmp/r02b_prof_valley \
    python profiles/prof_valley.py > $O/r02b_prof_valley.log 2>&1
python profiles/ncu_summary.py /tmp/r02b_prof_valley.ncu-rep > $O/r02b_ncu_valley_summary.csv
python bench_extra.py --reps 6 > $O/r02b_extra.json 2> $O/r02b_extra.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1
du -sh $O; ls -la $O | tail -12
