#!/bin/bash
TAG=${1:-r1c}
set -x
timeout 600 python -m pytest tests/test_gpu_tree.py tests/test_gpu_fullsize.py tests/test_gpu_shims.py -m gpu -x -q > gpurun_out/${TAG}_pytest_tree.log 2>&1; echo "pytest tree rc=$?"
tail -15 gpurun_out/${TAG}_pytest_tree.log
timeout 300 python tools/rollout_stats.py > gpurun_out/${TAG}_rollout_stats.log 2>&1; echo "stats rc=$?"; cat gpurun_out/${TAG}_rollout_stats.log
timeout 300 python tools/pure_ab.py > gpurun_out/${TAG}_pure_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/${TAG}_pure_ab.log
timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/${TAG}_bench_az.json 2> gpurun_out/${TAG}_bench_az.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_az.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pure_run -c 1 -o gpurun_out/${TAG}_pure_full python tools/profile_step.py --games 8192 --playouts 100 --pure 0 > gpurun_out/${TAG}_ncu_pure.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_ncu_pure.log
