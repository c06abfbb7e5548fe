#!/bin/bash
TAG=${1:-tl}
for a in 0 1; do
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_tl.so timeout -k 10 120 python tools/timeline.py $a > gpurun_out/${TAG}_timeline_$a.txt 2>&1; tail -45 gpurun_out/${TAG}_timeline_$a.txt
done
