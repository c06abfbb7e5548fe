#!/bin/bash
# device-resident epochs with the look-ahead: collate tests (bit-identity with host-collated steps), the reference-loop test, one bench line
TAG=${1:-ep}
timeout -k 10 200 python -m pytest tests/test_gpu_collate.py tests/test_reference_loop.py -x -q -m gpu > gpurun_out/${TAG}_test.log 2>&1; echo rc=$?
grep -v "Warning\|warn" gpurun_out/${TAG}_test.log | tail -25 | cut -c1-250
timeout -k 10 240 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sync", d["e2e"]["sync_every_step"]["value"], "epoch", d["e2e"]["device_resident_epoch"]["value"], d["e2e"]["device_resident_epoch"]["last_epoch_metrics"])
PY
