#!/bin/bash
# the driver's own N-GPU command (default flags: stage timing + CPU baseline on rank 0, e2e on every rank)
TAG=${1:-dn}
NG=${2:-2}
timeout -k 10 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $NG --steps 400 --warmup 20 > gpurun_out/${TAG}_bench_${NG}gpu_default.json 2> gpurun_out/${TAG}_bench_${NG}gpu_default.err
echo "rc=$?"; tail -3 gpurun_out/${TAG}_bench_${NG}gpu_default.err | cut -c1-200; head -c 500 gpurun_out/${TAG}_bench_${NG}gpu_default.json; echo
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${NG}gpu_default.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["dp_check"]["replicas_bit_identical"], d["launches_per_step"], d["gpu_launches"])
PY
