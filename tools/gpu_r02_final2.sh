#!/bin/bash
# final 1-GPU evidence, part 1: the whole GPU test suite, smoke, the default bench line (+ stages), the reference arm
set -u
TAG=${1:-r02_fin}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout -k 10 420 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log | cut -c1-200
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout -k 10 400 python bench.py --stages > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
grep launches gpurun_out/${TAG}_bench_stages.txt
head -c 400 gpurun_out/${TAG}_bench.json; echo
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"
head -c 400 gpurun_out/${TAG}_bench_reference.json; echo
bash tools/gpu_tl.sh $TAG > gpurun_out/${TAG}_tl.log 2>&1; tail -28 gpurun_out/${TAG}_timeline_1.txt
