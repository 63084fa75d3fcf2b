#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q -k "graph" > $O/r02_pytest10.log 2>&1; tail -15 $O/r02_pytest10.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --size 5792 --steps 5 --warmup 3 --no-cpu --no-extra > $O/r02_bench_n2_graph.json 2> $O/r02_bench_n2_graph.err; echo "bench2 rc=$?"
tail -c 600 $O/r02_bench_n2_graph.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --size 5792 --steps 5 --warmup 3 --no-cpu --no-extra --no-graph > $O/r02_bench_n2_eager.json 2> /dev/null; echo "bench2 eager rc=$?"
grep -o '"ms_per_step": [0-9.]*' $O/r02_bench_n2_graph.json $O/r02_bench_n2_eager.json
