#!/bin/bash
# quick cycle: the prep-ahead / pipelined-host tests, the live timeline, one bench line
TAG=${1:-r02_p}
timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prep_ahead or pipelined or fused_small or cfg1 or trainer or limits or status or abi or step_many" 2>&1 | tail -5
bash tools/gpu_tl.sh $TAG > gpurun_out/${TAG}_tl.log 2>&1; tail -42 gpurun_out/${TAG}_timeline_1.txt
timeout -k 10 240 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
