#!/bin/bash
# one test selection with the full log:  gpu_t1.sh <tag> <-k expression>
TAG=${1:-t1}
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$2" > gpurun_out/${TAG}_test.log 2>&1; echo rc=$?
grep -v "Warning\|warn" gpurun_out/${TAG}_test.log | tail -60 | cut -c1-250
