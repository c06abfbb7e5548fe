#!/bin/bash
# quick: fused-path tests + phase cycles + bench
set -u
TAG=${1:-r02_m}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -k "fused_small or cfg1 or trainer" -x -q > gpurun_out/${TAG}_fsg_test.log 2>&1; rc=$?; echo "fsg test rc=$rc"
tail -5 gpurun_out/${TAG}_fsg_test.log | cut -c1-250
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 1; fi
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_pt.so timeout -k 10 300 python tools/phase_timing.py > gpurun_out/${TAG}_phases.txt 2>&1; grep -E "fsg|k_ro" gpurun_out/${TAG}_phases.txt
timeout -k 10 400 python bench.py --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
grep launches gpurun_out/${TAG}_bench_stages.txt
head -c 300 gpurun_out/${TAG}_bench.json; echo
