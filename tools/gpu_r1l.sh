#!/bin/bash
TAG=${1:-r1l}
set -x
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_net.log 2>&1; echo "pytest rc=$?"
grep -E "resnet-|passed|failed|Error|error" gpurun_out/${TAG}_pytest_net.log | tail -20
for pr in fp16 split split_act; do
  timeout 300 python tools/profile_step.py --arch resnet --blocks 10 --precision $pr --playouts 40 --games 4096 2>&1 | tail -1
done
