#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list and one full capture of the node kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 600 python bench.py --stages > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.err
cat gpurun_out/bench.json | head -c 6000
NCUARGS="--resident 2 --steps 2 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-stage-timing"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NCUARGS > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_(fwd|bwd)' -s 8 -c 6 -f -o gpurun_out/prof_conv python bench.py $NCUARGS > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
