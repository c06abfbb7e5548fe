#!/bin/bash
# 2-GPU call: data-parallel parity tests (NCCL and peer exchange) + the 2-GPU bench line with each collective.
set -u
TAG=${1:-r01}
NG=${2:-2}
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -x > gpurun_out/${TAG}_dp_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_dp_tests.log
tail -15 gpurun_out/${TAG}_dp_tests.log
for COLL in nccl peer; do
  timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $NG --steps 400 --warmup 20 --collective $COLL --no-stage-timing --no-cpu-baseline > gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.json 2> gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.err
  echo "bench $COLL rc=$?"; tail -3 gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.err; head -c 700 gpurun_out/${TAG}_bench_${NG}gpu_${COLL}.json; echo
done
