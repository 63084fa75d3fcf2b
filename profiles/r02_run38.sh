#!/bin/bash
# N = 1 with the CUDA-graph replay of the sweep (as N > 1 already does) vs eager, same box
O=gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > $O/r02_bench38_graph.json 2> $O/r02_bench38_graph.err; echo "graph rc=$?"; tail -c 300 $O/r02_bench38_graph.err
python bench.py --steps 4 --warmup 3 --no-cpu --no-extra --no-graph > $O/r02_bench38_eager.json 2> $O/r02_bench38_eager.err; echo "eager rc=$?"
python - <<'PY'
import json
for f in ('graph','eager'):
    s=open(f'gpurun_out/r02_bench38_{f}.json').read()
    b=json.loads(s[s.index('{'):])
    print(f, 'ms', b['ms_per_step'], 'value', b['value'], 'e2e', b['e2e']['value'], 'launches', b['gpu_launches'], b['config']['launch'])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
