#!/bin/bash
# per-graph structure preparation: tests + cfg 5 bench line with stages
TAG=${1:-t2}
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "beyond or grouped or prep_structure or status" > gpurun_out/${TAG}_test.log 2>&1; echo rc=$?
grep -v "Warning\|warn" gpurun_out/${TAG}_test.log | tail -30 | cut -c1-250
timeout -k 10 300 python bench.py --workload large --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --stages > gpurun_out/${TAG}_bench_large.json 2> gpurun_out/${TAG}_bench_large_stages.txt; echo "bench rc=$?"
grep -E "launches|Error|error" gpurun_out/${TAG}_bench_large_stages.txt | head -30; head -c 400 gpurun_out/${TAG}_bench_large.json; echo
