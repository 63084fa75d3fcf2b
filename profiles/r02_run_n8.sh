#!/bin/bash
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 profiles/d2h_probe.py > $O/r02_d2h_probe_n8.json 2> $O/r02_d2h_probe_n8.err; echo "probe rc=$?"
tail -c 300 $O/r02_d2h_probe_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 3 --warmup 3 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench8 rc=$?"
tail -c 300 $O/r02_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tests/mgpu_check.py > $O/r02_mgpu8.log 2>&1; echo "mgpu rc=$?"; tail -3 $O/r02_mgpu8.log
