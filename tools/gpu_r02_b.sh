#!/bin/bash
# round 2, call B: tensor-core readout parity + timing, second tcgen05 probe, updated grad table
set -u
TAG=${1:-r02_b}
mkdir -p gpurun_out
timeout 300 python tools/umma_probe.py > gpurun_out/${TAG}_umma_probe.txt 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/${TAG}_umma_probe.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -30 gpurun_out/${TAG}_tests.log
timeout 300 python tools/debug_parity.py gcn_add_h32 gcn_cat_h32 gcn_add_h128 > gpurun_out/${TAG}_debug_parity.txt 2>&1; echo "debug rc=$?"
grep -E "=====|logp|grad fc|grad fc1_bn|grad fc2_bn|loss|NONFINITE" gpurun_out/${TAG}_debug_parity.txt | head -80
timeout 400 python bench.py --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_stages.txt
head -c 300 gpurun_out/${TAG}_bench.json; echo
CAL_READOUT=legacy timeout 400 python bench.py --steps 200 --stages --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_legacy.json 2> gpurun_out/${TAG}_bench_legacy_stages.txt; echo "bench legacy rc=$?"
grep readout gpurun_out/${TAG}_bench_legacy_stages.txt
timeout 900 python tools/grad_table.py > gpurun_out/${TAG}_grad_errors.txt 2>&1; echo "grad table rc=$?"
grep -E "^==" gpurun_out/${TAG}_grad_errors.txt | cut -c1-200
grep -E "FAIL|illcond" gpurun_out/${TAG}_grad_errors.txt | head -30
