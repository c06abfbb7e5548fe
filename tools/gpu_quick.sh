#!/bin/bash
# Quick single-GPU check: GPU tests + the bench line with the per-stage table.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --stages ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 1700 gpurun_out/${TAG}_bench_stages.txt
head -c 300 gpurun_out/${TAG}_bench.json; echo
if [ -f cal_b200/libcal_b200_pt.so ]; then
  CAL_B200_LIB=$PWD/cal_b200/libcal_b200_pt.so timeout 200 python tools/phase_timing.py > gpurun_out/${TAG}_phases.txt 2>&1; echo "phase timing rc=$?"
  tail -8 gpurun_out/${TAG}_phases.txt
fi
