#!/bin/bash
# round 2, call F: fused small-graph backward: parity vs the tiled kernels + oracle, full GPU tests, phase cycles, bench
set -u
TAG=${1:-r02_f}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -k fused_small -x -q > gpurun_out/${TAG}_fsg_test.log 2>&1; rc=$?; echo "fsg test rc=$rc"
tail -30 gpurun_out/${TAG}_fsg_test.log | cut -c1-250
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG in the fused kernel: stopping"; exit 1; fi
timeout -k 10 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log | cut -c1-250
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_pt.so timeout -k 10 300 python tools/phase_timing.py > gpurun_out/${TAG}_phases.txt 2>&1; grep -E "fsg" gpurun_out/${TAG}_phases.txt
timeout -k 10 400 python bench.py --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_stages.txt
head -c 300 gpurun_out/${TAG}_bench.json; echo
