#!/bin/bash
# N-GPU call: the bench line with each gradient-exchange implementation (weak scaling, batch 128 per GPU).
set -u
TAG=${1:-r01}
NG=${2:-8}
mkdir -p gpurun_out
for COLL in peer nccl; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $NG --steps 400 --warmup 20 --collective $COLL --no-stage-timing > gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.json 2> gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.err
  echo "bench $NG $COLL rc=$?"; tail -2 gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.err | cut -c1-300; head -c 400 gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.json; echo
done
