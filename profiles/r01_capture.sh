#!/bin/bash
# Round-1 measurement + profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/,
# which is limited to 64 MiB: the big ncu reports are condensed to CSV on the box and dropped).
set -x
O=gpurun_out
python bench.py > $O/r01_bench_n1.json 2> $O/r01_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r01_bench_reference.json 2> $O/r01_bench_reference.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r01_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/r01_ncu_launches.log 2>&1
# full-set captures of the kernels of config 4: one tpi + std + gradient call per size class on an 8192^2 DEM in
# sweep mode (PROF_SHARE: plane cache + octagon walk), condensed to CSV
PROF_SIZE=8192 PROF_SHARE=801 ncu --set full --clock-control none -c 60 -f -o /tmp/r01_prof \
    python profiles/prof_driver.py tpi:801 std:801 grad:801 tpi:21 tpi:5 grad:5 sobel:0 > $O/r01_prof.log 2>&1
python profiles/ncu_summary.py /tmp/r01_prof.ncu-rep > $O/r01_ncu_full_summary.csv
ncu --set full --clock-control none -k regex:sx_tma -c 1 -f -o /tmp/r01_prof_sx \
    python bench_extra.py --reps 2 > /dev/null 2> $O/r01_prof_sx.log
python profiles/ncu_summary.py /tmp/r01_prof_sx.ncu-rep > $O/r01_ncu_sx_summary.csv
PROF_SIZE=2048 ncu --set full --clock-control none -k regex:valley -c 1 -f -o /tmp/r01_prof_valley \
    python profiles/prof_valley.py > $O/r01_prof_valley.log 2>&1
python profiles/ncu_summary.py /tmp/r01_prof_valley.ncu-rep > $O/r01_ncu_valley_summary.csv
python bench_extra.py --reps 6 > $O/r01_extra.json 2> $O/r01_extra.err
python bench_c5.py --size 4096 --reps 1 > $O/r01_c5_n1_4096.json 2> $O/r01_c5_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r01_smoke.log 2>&1
du -sh $O; ls -la $O
