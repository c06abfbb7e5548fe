#!/bin/bash
# ncu evidence: launch list of a step (device time per launch) + one --set full capture with source of the fused-path kernels
set -u
TAG=${1:-r02_ncu}
mkdir -p gpurun_out
# (bench.py runs its warm-up + timed sequence once untimed first: 2 x 9 steps x 7 kernels; the last 6 steps are captured)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 84 -c 42 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 6 --warmup 3 --no-stage-timing --no-cpu-baseline --no-graph --no-e2e > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
tail -3 gpurun_out/${TAG}_launches.csv | cut -c1-200
if [ "${2:-}" = "list" ]; then exit 0; fi
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'k_ro_|k_fsg_|k_prep_small|k_adam' -s 64 -c 8 -f -o gpurun_out/${TAG}_prof \
  python bench.py --steps 6 --warmup 3 --no-stage-timing --no-cpu-baseline --no-graph --no-e2e > gpurun_out/${TAG}_prof.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_prof.ncu-rep
