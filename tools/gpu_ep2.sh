#!/bin/bash
TAG=${1:-ep}
bash tools/gpu_ep.sh $TAG
bash tools/gpu_tl_ep.sh $TAG | tail -16
