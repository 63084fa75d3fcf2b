#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest6.log
tail -5 $O/r02_pytest6.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke6.log 2>&1; tail -2 $O/r02_smoke6.log
python bench.py --steps 3 --warmup 3 > $O/r02_bench6.json 2> $O/r02_bench6.err; echo "bench rc=$?"; tail -c 800 $O/r02_bench6.err
