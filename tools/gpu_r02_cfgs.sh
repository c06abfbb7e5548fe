#!/bin/bash
# 1-GPU bench lines of the other configurations: cfg 3 (CausalGAT, MUTAG-shaped), cfg 5 shapes (large, fp32 and bf16 readout), CausalGIN
set -u
TAG=${1:-r02_cfgs}
mkdir -p gpurun_out
timeout -k 10 300 python bench.py --model CausalGAT --workload mutag --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench_gat.json 2> gpurun_out/${TAG}_bench_gat_stages.txt; echo "gat rc=$?"
grep launches gpurun_out/${TAG}_bench_gat_stages.txt; head -c 250 gpurun_out/${TAG}_bench_gat.json; echo
timeout -k 10 400 python bench.py --workload large --steps 50 --warmup 5 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench_large.json 2> gpurun_out/${TAG}_bench_large_stages.txt; echo "large rc=$?"
grep launches gpurun_out/${TAG}_bench_large_stages.txt; head -c 250 gpurun_out/${TAG}_bench_large.json; echo
timeout -k 10 400 python bench.py --workload large --readout-bf16 --steps 50 --warmup 5 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench_large_bf16.json 2> gpurun_out/${TAG}_bench_large_bf16_stages.txt; echo "large bf16 rc=$?"
grep "readout" gpurun_out/${TAG}_bench_large_bf16_stages.txt; head -c 250 gpurun_out/${TAG}_bench_large_bf16.json; echo
timeout -k 10 300 python bench.py --model CausalGIN --steps 200 --stages --no-cpu-baseline > gpurun_out/${TAG}_bench_gin.json 2> gpurun_out/${TAG}_bench_gin_stages.txt; echo "gin rc=$?"
head -c 250 gpurun_out/${TAG}_bench_gin.json; echo
