#!/bin/bash
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --size 5792 --steps 5 --warmup 3 --no-cpu > $O/r02_bench_n2_small.json 2> $O/r02_bench_n2_small.err; echo "bench2 rc=$?"
tail -c 400 $O/r02_bench_n2_small.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tests/mgpu_check.py > $O/r02_mgpu2.log 2>&1; echo "mgpu rc=$?"; grep -i "mgpu_check\|MISMATCH" $O/r02_mgpu2.log | head
timeout 600 python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q -k "torchrun or valley_ridge_fft" > $O/r02_pytest_n2.log 2>&1; tail -3 $O/r02_pytest_n2.log
