#!/bin/bash
TAG=${1:-r1o}
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_net.log 2>&1; echo "pytest rc=$?"
grep -E "incep|passed|failed|Error|error|mode" gpurun_out/${TAG}_pytest_net.log | tail -12
timeout 300 python tools/profile_step.py --arch inception --blocks 10 --playouts 40 --games 4096 2>&1 | tr '\n' ' ' | sed 's/phase ms/\nphase ms/' | cut -c1-700; echo
