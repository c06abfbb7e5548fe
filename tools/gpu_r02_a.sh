#!/bin/bash
# round 2, call A: tightened parity tests, tcgen05 probe, gradient error table, sanitizer runs, bench smoke
set -u
TAG=${1:-r02_a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python tools/umma_probe.py > gpurun_out/${TAG}_umma_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/${TAG}_umma_probe.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -40 gpurun_out/${TAG}_tests.log
timeout 900 python tools/grad_table.py --large > gpurun_out/${TAG}_grad_errors.txt 2>&1; echo "grad table rc=$?"
grep "^==" gpurun_out/${TAG}_grad_errors.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/${TAG}_racecheck.log
timeout 400 python bench.py --steps 200 --stages > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_stages.txt
head -c 600 gpurun_out/${TAG}_bench.json; echo
