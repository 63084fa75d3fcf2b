#!/bin/bash
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest20.log 2>&1; tail -6 $O/r02_pytest20.log
bash profiles/r02_run19.sh
