#!/bin/bash
# final 1-GPU evidence, part 2: the other configurations, compute-sanitizer, ncu launch list + full capture
set -u
TAG=${1:-r02_fin}
bash tools/gpu_r02_cfgs.sh ${TAG}
bash tools/gpu_r02_sanitizer.sh ${TAG}
bash tools/gpu_r02_ncu.sh ${TAG}
