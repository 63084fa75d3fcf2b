#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_params.py -m gpu -x -q -k "disc_fft" > $O/r02_pytest13.log 2>&1; tail -12 $O/r02_pytest13.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest13b.log 2>&1; tail -4 $O/r02_pytest13b.log
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02_bench13.json 2> $O/r02_bench13.err; echo "bench rc=$?"; tail -c 500 $O/r02_bench13.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r02_bench13.json'))
print('float ms', b['ms_per_step'], 'int ms', b['extra']['sweep_integer_dem']['ms_per_step'], 'e2e', b['e2e']['value'])
for k,v in list(b['kernels'].items())[:24]: print(f"{k:34s} {v['launches']:3d} {v['ms']:8.2f} avg {v['avg_ms']:.3f}")
PY
