#!/bin/bash
# rebuild the library on the GPU box with different resident-CTA targets for the pure kernel and time each
set -x
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_fullsize.py tests/test_gpu_shims.py -m gpu -x -q > gpurun_out/sweep_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/sweep_pytest.log
for mb in 32 28 24 20; do
  AP_NVCC_EXTRA="-DAP_PURE_MINBLK=$mb" python -m alphapig_b200.build --force > /dev/null 2>&1
  echo "== MINBLK $mb"; timeout 300 python tools/pure_ab.py 2>&1 | grep pure_run
done
python -m alphapig_b200.build --force > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pure_run -c 1 -o gpurun_out/r1g_pure_full python tools/profile_step.py --games 8192 --playouts 100 --pure 0 > gpurun_out/r1g_ncu_pure.log 2>&1; echo "ncu rc=$?"
