#!/bin/bash
# mask spectrum shared by the tpi + std pair through the plane cache: tests + bench
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest32.log 2>&1; tail -5 $O/r02_pytest32.log
python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > $O/r02_bench32.json 2> $O/r02_bench32.err; echo "bench rc=$?"; tail -c 400 $O/r02_bench32.err
python - <<'PY'
import json
s=open('gpurun_out/r02_bench32.json').read()
b=json.loads(s[s.index('{'):])
print('float ms', b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'])
for k,v in list(b['kernels'].items())[:12]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f} avg {v['avg_ms']:.3f}")
print({k:v for k,v in b['memory_bound'].items() if 'fft' in k})
PY
