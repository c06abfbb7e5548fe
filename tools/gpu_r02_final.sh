#!/bin/bash
# round-2 evidence call: smi, smoke, the default bench line (with CPU baseline and e2e), the reference arm
set -u
TAG=${1:-r02_zz}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout -k 10 500 python bench.py --stages > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
grep launches gpurun_out/${TAG}_bench_stages.txt
head -c 3000 gpurun_out/${TAG}_bench.json; echo
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
head -c 900 gpurun_out/${TAG}_bench_reference.json; echo
