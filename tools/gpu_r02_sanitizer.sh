#!/bin/bash
# compute-sanitizer over the fused small-graph path (forward, readouts, backward, gradient reduce): memcheck, then racecheck / synccheck
set -u
TAG=${1:-r02_zz}
mkdir -p gpurun_out
T='tests/test_gpu_parity.py -x -q -p no:cacheprovider'
K='fused_small_graph_kernels_match or (prep_ahead and eager) or per_graph or limits'
for TOOL in memcheck synccheck racecheck; do
  timeout -k 10 420 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest $T -k "$K" > gpurun_out/${TAG}_sanitizer_${TOOL}.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_${TOOL}.log | tail -3
done
