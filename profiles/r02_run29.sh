#!/bin/bash
# twin tiles for lone planes (replaces the std look-ahead): full GPU suite + bench
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest29.log 2>&1; tail -6 $O/r02_pytest29.log
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02_bench29.json 2> $O/r02_bench29.err; echo "bench rc=$?"; tail -c 500 $O/r02_bench29.err
python - <<'PY'
import json
s=open('gpurun_out/r02_bench29.json').read()
b=json.loads(s[s.index('{'):])
print('float ms', b['ms_per_step'], 'int ms', b['extra']['sweep_integer_dem']['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], 'c5 valley', b['extra']['config5']['valley_ridge']['ms'])
for k,v in list(b['kernels'].items())[:26]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f} avg {v['avg_ms']:.3f}")
PY
