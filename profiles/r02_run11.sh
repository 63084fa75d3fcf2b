#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest11.log 2>&1; tail -3 $O/r02_pytest11.log
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02_bench11.json 2> $O/r02_bench11.err; echo "bench rc=$?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench11.json'))
print('float ms', b['ms_per_step'], 'int ms', b['extra']['sweep_integer_dem']['ms_per_step'], 'e2e', b['e2e']['value'])
for k,v in list(b['kernels'].items())[:12]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f}")
PY
