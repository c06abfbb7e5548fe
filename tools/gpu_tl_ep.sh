#!/bin/bash
TAG=${1:-tlep}
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_tl.so timeout -k 10 150 python tools/timeline.py 1 epoch > gpurun_out/${TAG}_timeline_epoch.txt 2>&1; grep -A14 "epoch path" gpurun_out/${TAG}_timeline_epoch.txt | tail -32; tail -3 gpurun_out/${TAG}_timeline_epoch.txt | cut -c1-200
