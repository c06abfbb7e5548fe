#!/bin/bash
# 2-GPU: exchange timeline + DP parity tests + peer bench line
TAG=${1:-dp2}
bash tools/gpu_tl_dp.sh $TAG 2 2>&1 | grep -v "^$" | tail -44
timeout -k 10 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -x > gpurun_out/${TAG}_dp_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_dp_tests.log
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 400 --warmup 20 --collective peer --no-stage-timing --no-cpu-baseline > gpurun_out/${TAG}_bench_2gpu_peer.json 2> gpurun_out/${TAG}_bench_2gpu_peer.err
echo "bench rc=$?"; head -c 330 gpurun_out/${TAG}_bench_2gpu_peer.json; echo; grep -o '"dp_check": {[^}]*}' gpurun_out/${TAG}_bench_2gpu_peer.json | head -c 600; echo
