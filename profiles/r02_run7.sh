#!/bin/bash
O=gpurun_out
python bench.py --size 5792 --steps 5 --warmup 3 --no-cpu --no-extra > $O/r02_bench7_small.json 2> $O/r02_bench7_small.err; echo "rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench7_small.json'))
print('ms_per_step', b['ms_per_step'], 'sum kernels', sum(v['ms'] for v in b['kernels'].values()), 'launches', b['gpu_launches'])
PY
