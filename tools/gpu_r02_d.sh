#!/bin/bash
# round 2, call D: resident-tile tensor-core readout: parity, phase cycles, stage timing
set -u
TAG=${1:-r02_d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log | cut -c1-250
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_pt.so timeout 300 python tools/phase_timing.py > gpurun_out/${TAG}_phases.txt 2>&1; grep readout gpurun_out/${TAG}_phases.txt
timeout 400 python bench.py --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_stages.txt
head -c 300 gpurun_out/${TAG}_bench.json; echo
timeout 900 python tools/grad_table.py > gpurun_out/${TAG}_grad_errors.txt 2>&1; echo "grad table rc=$?"
grep -E "FAIL|illcond" gpurun_out/${TAG}_grad_errors.txt | head -30
