#!/bin/bash
TAG=${1:-tldp}
NG=${2:-2}
for m in distinct same; do
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_tl.so timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 tools/timeline_dp.py $m > gpurun_out/${TAG}_timeline_dp_$m.txt 2>&1
grep -v "Warn\|warn\|OMP\|\*\*\*" gpurun_out/${TAG}_timeline_dp_$m.txt | tail -40
done
