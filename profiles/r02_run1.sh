#!/bin/bash
# Round 2, first GPU pass: the whole GPU test-suite, smoke, and the new bench line (short).
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest1.log
tail -5 $O/r02_pytest1.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke1.log 2>&1; tail -2 $O/r02_smoke1.log
python bench.py --steps 3 --warmup 3 > $O/r02_bench1.json 2> $O/r02_bench1.err; echo "bench rc=$?"; tail -c 600 $O/r02_bench1.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/r02_bench1_ref.json 2> $O/r02_bench1_ref.err; echo "ref rc=$?"
