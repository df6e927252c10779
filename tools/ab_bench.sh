#!/bin/bash
# A/B of environment knobs on the headline workload: tools/ab_bench.sh out_dir "VAR=val ..." "VAR2=val ..." ...
# (each configuration: python bench.py --legs none --no-cpu; prints value, ms/step, clock, per-layer phase times)
OUT=$1; shift
mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --legs none --no-cpu > $OUT/ab_$i.json 2> $OUT/ab_$i.err
  python - "$OUT/ab_$i.json" "$cfg" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
r = d["roofline"]; ph = r["phase_ms_per_lockstep"]
print("%-40s %8d playouts/s %7.1f ms/step %5.0f MHz frac %.3f convs %s sel %.4f heads %.4f exp %.4f" % (
    sys.argv[2], d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], r["frac"], [round(x, 4) for x in ph["trunk_convs"]],
    ph["select"], ph["heads"], ph["expand_backup"]))
PY
done
