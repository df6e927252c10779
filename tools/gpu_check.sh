#!/bin/bash
# GPU-box check used during development: parity tests, smoke, bench, ncu launch list.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-chk}
set -x
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
for m in ${MODES:-2}; do
AP_HEAD_MODE=$m timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench_m$m.json 2> gpurun_out/${TAG}_bench_m$m.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_m$m.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum --print-units base --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --games 4096 --playouts 4 > gpurun_out/${TAG}_prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_prof.log
# configs[2]: fused mcts_pure kernel, A/B of the rollout implementations, launch list of the full-size launch
timeout 300 python tools/pure_ab.py > gpurun_out/${TAG}_pure_ab.log 2>&1; cat gpurun_out/${TAG}_pure_ab.log
timeout 300 python bench.py --workload pure --no-cpu > gpurun_out/${TAG}_bench_pure.json 2> gpurun_out/${TAG}_bench_pure.err; echo "bench pure rc=$?"; cat gpurun_out/${TAG}_bench_pure.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --print-units base --clock-control none -k regex:k_pure_run --csv --log-file gpurun_out/${TAG}_launches_pure.csv python tools/profile_step.py --games 8192 --playouts 1000 --pure 0 > gpurun_out/${TAG}_prof_pure.log 2>&1; echo "ncu pure rc=$?"
