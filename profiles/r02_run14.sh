#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest14.log 2>&1; tail -4 $O/r02_pytest14.log
