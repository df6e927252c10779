#!/bin/bash
# multi-GPU bench lines (torchrun, one rank per GPU): az headline + pure
N=${1:-2}
TAG=${2:-r1k}
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --workload pure > gpurun_out/${TAG}_bench_pure_${N}gpu.json 2> gpurun_out/${TAG}_bench_pure_${N}gpu.err; echo "bench pure rc=$?"; tail -1 gpurun_out/${TAG}_bench_pure_${N}gpu.json
