#!/bin/bash
# second-session GPU check: rollout tests first, then A/B, full suite, benches, ncu of the pure kernel
TAG=${1:-r1b}
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_tree.py -m gpu -x -q > gpurun_out/${TAG}_pytest_tree.log 2>&1; echo "pytest tree rc=$?"
tail -15 gpurun_out/${TAG}_pytest_tree.log
timeout 300 python tools/pure_ab.py > gpurun_out/${TAG}_pure_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/${TAG}_pure_ab.log
timeout 300 python bench.py --workload pure --no-cpu > gpurun_out/${TAG}_bench_pure.json 2> gpurun_out/${TAG}_bench_pure.err; echo "bench pure rc=$?"; cat gpurun_out/${TAG}_bench_pure.json
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_tree.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench_az.json 2> gpurun_out/${TAG}_bench_az.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_az.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pure_run -c 1 -o gpurun_out/${TAG}_pure_full python tools/profile_step.py --games 8192 --playouts 100 --pure 0 > gpurun_out/${TAG}_ncu_pure.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_ncu_pure.log
