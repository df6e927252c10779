#!/bin/bash
# development: where does the loop lose time on N GPUs?  normal / no trainer work / no collectives, with per-rank detail
N=${1:-8}; OUT=${2:-gpurun_out/loop_dbg}; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload loop --steps 3 "$@"; }
AP_LOOP_DETAIL=1 run 2>$OUT/normal.err | tail -1 > $OUT/normal.json
AP_LOOP_DETAIL=1 AP_LOOP_NOTRAIN=1 run 2>$OUT/notrain.err | tail -1 > $OUT/notrain.json
AP_LOOP_DETAIL=1 AP_LOOP_NOCOLL=1 run 2>$OUT/nocoll.err | tail -1 > $OUT/nocoll.json
for f in normal notrain nocoll; do echo "== $f: $(python -c "import json; d=json.load(open('$OUT/$f.json')); print(round(d['value']), d['seconds'], d['train_steps'], d['weight_swaps'])")"; grep "^rank" $OUT/$f.err | sort | cut -c1-330; done
