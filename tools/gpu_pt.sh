#!/bin/bash
TAG=${1:-pt}
CAL_B200_LIB=$PWD/cal_b200/libcal_b200_pt.so timeout -k 10 300 python tools/phase_timing.py > gpurun_out/${TAG}_phases.txt 2>&1; grep -E "fsg|k_ro" gpurun_out/${TAG}_phases.txt
