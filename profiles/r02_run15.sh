#!/bin/bash
# N=2 validation of the FFT disc route in row bands (full-size bench as the driver launches it + band==whole checks)
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench15_n2.json 2> $O/r02_bench15_n2.err; echo "bench2 rc=$?"
tail -c 600 $O/r02_bench15_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tests/mgpu_check.py > $O/r02_mgpu15.log 2>&1; echo "mgpu rc=$?"; grep -i "mgpu_check\|MISMATCH" $O/r02_mgpu15.log | head
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench15_n2.json'))
print('value', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launch', b['config'].get('launch'), 'band_check', b.get('band_check'))
for k,v in list(b['kernels'].items())[:12]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f} avg {v['avg_ms']:.3f}")
PY
