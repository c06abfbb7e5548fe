#!/bin/bash
# 8-GPU call: cfg 4 (weak scaling, batch 128 per GPU, peer exchange and NCCL), the exchange-cost experiment (same batches on
# every rank), cfg 5 (large, batch 512 per GPU, bf16 readout MLP and fp32)
set -u
TAG=${1:-r02_s8}
NG=${2:-8}
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $NG --no-stage-timing --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name rc=$?"; tail -2 gpurun_out/${TAG}_${name}.err | cut -c1-200; head -c 330 gpurun_out/${TAG}_${name}.json; echo
}
run bench_${NG}gpu_peer --steps 400 --warmup 20 --collective peer
run bench_${NG}gpu_nccl --steps 400 --warmup 20 --collective nccl
if [ "${3:-}" = "all" ]; then
run bench_${NG}gpu_peer_same --steps 400 --warmup 20 --collective peer --same-batches --no-e2e
run bench_${NG}gpu_large_bf16 --workload large --readout-bf16 --steps 40 --warmup 5 --no-e2e
fi
run bench_${NG}gpu_large --workload large --steps 40 --warmup 5 --no-e2e
