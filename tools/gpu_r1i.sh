#!/bin/bash
TAG=${1:-r1i}
set -x
timeout 1500 python -m pytest tests/test_gpu_net.py tests/test_gpu_shims.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
for ks in 4 1 2 6; do
AP_FC_KSPLIT=$ks timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/${TAG}_bench_az_ks$ks.json 2> gpurun_out/${TAG}_bench_az_ks$ks.err; echo "bench ks=$ks rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_az_ks$ks.json')); print(d['value'], d['e2e']['value'], d['roofline']['phase_ms_per_lockstep'])"
done
