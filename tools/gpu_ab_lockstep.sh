#!/bin/bash
# A/B of the fused lock-step (ticket + features in k_select, FC finish in k_expand_backup) against the unfused one
TAG=${1:-ab}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for ck in 0 1 0 1; do
AP_COMPACT_KERNEL=$ck timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/${TAG}_bench_ck$ck.json 2> gpurun_out/${TAG}_bench_ck$ck.err; echo "unfused=$ck rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_ck$ck.json')); print(d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['phase_ms_per_lockstep'])"
done
