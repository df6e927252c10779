set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r1c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1c_smoke.log
timeout 600 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo "bench rc=$?"; cat gpurun_out/r1c_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1c_launches.csv python tools/profile_step.py --games 4096 --playouts 4 > gpurun_out/r1c_prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r1c_prof.log
