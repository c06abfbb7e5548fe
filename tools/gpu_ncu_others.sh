#!/bin/bash
# ncu --set full of the dominant kernels of the other workloads (cfg 5 shapes, cfg 3 CausalGAT): DRAM traffic for ncu_traffic.json
set -u
TAG=${1:-r02_nco}
mkdir -p gpurun_out
timeout -k 10 400 ncu --set full --clock-control none -k regex:'k_conv_bwd|k_conv_fwd|k_masked_bwd_gemm|k_masked_bwd_gather|k_masked_fwd' -s 10 -c 6 -f -o gpurun_out/${TAG}_large \
  python bench.py --workload large --steps 2 --warmup 1 --resident 4 --no-stage-timing --no-cpu-baseline --no-graph --no-e2e > gpurun_out/${TAG}_large.log 2>&1; echo "large rc=$?"
timeout -k 10 300 ncu --set full --clock-control none -k regex:'k_gat_' -s 16 -c 8 -f -o gpurun_out/${TAG}_gat \
  python bench.py --model CausalGAT --workload mutag --steps 2 --warmup 1 --no-stage-timing --no-cpu-baseline --no-graph --no-e2e > gpurun_out/${TAG}_gat.log 2>&1; echo "gat rc=$?"
ls -la gpurun_out/${TAG}_large.ncu-rep gpurun_out/${TAG}_gat.ncu-rep
