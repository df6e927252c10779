#!/bin/bash
TAG=${1:-r1h}
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/${TAG}_bench_az.json 2> gpurun_out/${TAG}_bench_az.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_az.json')); print(d['value'], d['e2e']['value'], d['roofline']['phase_ms_per_lockstep'])"
AP_COMPACT_KERNEL=1 timeout 600 python bench.py --no-cpu --steps 2 > gpurun_out/${TAG}_bench_az_ck.json 2> gpurun_out/${TAG}_bench_az_ck.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_az_ck.json')); print(d['value'], d['e2e']['value'], d['roofline']['phase_ms_per_lockstep'])"
