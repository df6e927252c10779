#!/bin/bash
# configs[4] on N GPUs of one box: self-play alone vs the self-play + train loop (trainer_share A/B).
#   gpurun --gpus N -- 'tools/gpu_loop_ab.sh N out_dir'
N=${1:-2}; OUT=${2:-gpurun_out/loop_ab}; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --workload selfplay --device-pick --device-records --steps 4 --warmup 1 2>$OUT/selfplay.err | tail -1 > $OUT/selfplay.json
for share in ${SHARES:-0.0 0.08}; do
  run --workload loop --steps 3 --trainer-share $share 2>$OUT/loop_$share.err | tail -1 > $OUT/loop_$share.json
done
python - $OUT <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/*.json")):
    d = json.load(open(f))
    c = d.get("collectives") or {}
    print("%-28s %10.0f playouts/s  %8.0f moves/s  train_steps %s swaps %s trainer_s %s main-thread exchanges %.3f s gathered %s B bcast %s B" % (
        f.split("/")[-1], d.get("playouts_per_s", d["value"]), d.get("moves_per_s", d["value"]), d.get("train_steps"),
        d.get("weight_swaps"), d.get("trainer_seconds"), c.get("seconds_main_thread_max", 0.0), c.get("bytes_gathered"), c.get("bytes_broadcast")))
PY
