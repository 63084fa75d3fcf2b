#!/bin/bash
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python profiles/sanitize_fft.py > $O/r02_sanitize_memcheck.log 2>&1; tail -4 $O/r02_sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python profiles/sanitize_fft.py > $O/r02_sanitize_racecheck.log 2>&1; tail -4 $O/r02_sanitize_racecheck.log
