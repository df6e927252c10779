#!/bin/bash
# final evidence of this session: launch lists (az lock-step, pure), full-set captures of the tree kernels, bench lines
TAG=${1:-r1j}
set -x
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --print-units base --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_az.csv python tools/profile_step.py --games 4096 --playouts 4 > gpurun_out/${TAG}_prof_az.log 2>&1; echo "ncu az rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --print-units base --clock-control none -k regex:k_pure_run --csv --log-file gpurun_out/${TAG}_launches_pure.csv python tools/profile_step.py --games 8192 --playouts 1000 --pure 0 > gpurun_out/${TAG}_prof_pure.log 2>&1; echo "ncu pure rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_select|k_expand_backup|k_emit_features|k_head_fc" -s 40 -c 5 -o gpurun_out/${TAG}_tree_full python tools/profile_step.py --games 4096 --playouts 12 > gpurun_out/${TAG}_ncu_tree.log 2>&1; echo "ncu tree rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench_az.json 2> gpurun_out/${TAG}_bench_az.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_az.json
timeout 600 python bench.py --workload pure > gpurun_out/${TAG}_bench_pure.json 2> gpurun_out/${TAG}_bench_pure.err; echo "bench pure rc=$?"; cat gpurun_out/${TAG}_bench_pure.json
timeout 600 python bench.py --workload selfplay > gpurun_out/${TAG}_bench_selfplay.json 2> gpurun_out/${TAG}_bench_selfplay.err; echo "bench selfplay rc=$?"; cat gpurun_out/${TAG}_bench_selfplay.json
