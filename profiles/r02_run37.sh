#!/bin/bash
# final check of the round-2 build on one GPU: smoke + the default bench line
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.log 2>&1; tail -2 $O/r02c_smoke.log
python bench.py > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err; echo "bench rc=$?"; tail -c 300 $O/r02c_bench_n1.err
python - <<'PY'
import json
s=open('gpurun_out/r02c_bench_n1.json').read()
b=json.loads(s[s.index('{'):])
print('float ms', b['ms_per_step'], 'value', b['value'], 'int ms', b['extra']['sweep_integer_dem']['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], b['clocks'])
print(b['roofline']['kernel'], b['roofline']['frac'], b['roofline']['traffic'], b['cpu_baseline']['value'])
PY
