#!/bin/bash
O=gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02_bench39.json 2> $O/r02_bench39.err; echo "bench rc=$?"; tail -c 300 $O/r02_bench39.err
python - <<'PY'
import json
s=open('gpurun_out/r02_bench39.json').read()
b=json.loads(s[s.index('{'):])
print('float ms', b['ms_per_step'], 'int ms', b['extra']['sweep_integer_dem']['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], b['config']['launch'], 'c5', b['extra']['config5']['valley_ridge']['ms'], 'c3', b['extra']['config3_sx']['ms_per_step'])
PY
