#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list and one full capture of the node kernels.
# Every step runs under its own timeout; logs land in gpurun_out/.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --stages > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench_stages.txt; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_stages.txt
head -c 1500 gpurun_out/${TAG}_bench.json; echo
NCUARGS="--resident 2 --steps 2 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-stage-timing"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $NCUARGS > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_(fwd|bwd)' -s 8 -c 6 -f -o gpurun_out/${TAG}_prof_conv python bench.py $NCUARGS > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
# other configurations (BASELINE.json configs[2] and configs[4] shapes), single GPU, no CPU leg
timeout 300 python bench.py --workload large --steps 30 --warmup 5 --no-cpu-baseline --stages > gpurun_out/${TAG}_bench_large.json 2> gpurun_out/${TAG}_bench_large_stages.txt; echo "bench large rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_large_stages.txt
timeout 300 python bench.py --workload mutag --model CausalGAT --steps 200 --warmup 20 --no-cpu-baseline --stages > gpurun_out/${TAG}_bench_gat.json 2> gpurun_out/${TAG}_bench_gat_stages.txt; echo "bench gat rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_gat_stages.txt
timeout 300 python bench.py --model CausalGIN --steps 200 --warmup 20 --no-cpu-baseline --stages > gpurun_out/${TAG}_bench_gin.json 2> gpurun_out/${TAG}_bench_gin_stages.txt; echo "bench gin rc=$?"
tail -c 1800 gpurun_out/${TAG}_bench_gin_stages.txt
