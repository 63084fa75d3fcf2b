#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q > $O/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest2.log
tail -5 $O/r02_pytest2.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_bench_params.py > $O/r02_pytest2b.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest2b.log
tail -5 $O/r02_pytest2b.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $O/r02_bench2.json 2> $O/r02_bench2.err; echo "bench rc=$?"; tail -c 600 $O/r02_bench2.err
