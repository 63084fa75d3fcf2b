#!/bin/bash
# 8-GPU run of the round-2 build: bench as the driver launches it + band == whole check
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 > $O/r02b_bench_n4.json 2> $O/r02b_bench_n4.err; echo "bench4 rc=$?"
tail -c 300 $O/r02b_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 tests/mgpu_check.py > $O/r02b_mgpu4.log 2>&1; echo "mgpu rc=$?"; grep -i "mgpu_check\|MISMATCH" $O/r02b_mgpu4.log | head -3
python - <<'PY'
import json
s=open('gpurun_out/r02b_bench_n4.json').read()
b=json.loads(s[s.index('{'):])
print('value', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], b['config'].get('launch'), 'band_check', b.get('band_check'))
for k,v in list(b['kernels'].items())[:14]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f} avg {v['avg_ms']:.3f}")
PY
