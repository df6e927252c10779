#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (tools/sanitize_small.py) + the minimal repro of
# the one report class that is a tool limitation (tools/sanitizer_repro/tmem_alloc_pair.cu).  Exit code != 0 when a
# tool reports anything that is not on the allow-list below.
#   usage: tools/gpu_sanitize.sh [out_dir]      (TOOLS="memcheck racecheck synccheck" to pick tools)
# Allow-list (tools/sanitize_allow.py): racecheck hazards between the tcgen05.alloc.cta_group::2 result write (write
# PC outside the kernel, offset 0xffff...) and the read of the TMEM base address after __syncthreads + cluster barrier
# in the CTA-pair kernels - reproduced on a 40-line kernel that contains nothing else.
OUT=${1:-gpurun_out}
mkdir -p $OUT
rc=0
i=0
[ -n "$SKIP_REPRO" ] || for cfg in "s 128 0 2 64 1" "p 128 0 2 64 1" "s 384 2 148 256 3" "p 384 2 148 256 3" "p 384 2 148 512 3"; do
  i=$((i+1))
  compute-sanitizer --tool racecheck --print-limit 5 tools/sanitizer_repro/tmem_alloc_pair $cfg > $OUT/sanitize_repro_racecheck_$i.log 2>&1
  echo "repro racecheck [$cfg]: $(grep -c 'Race reported' $OUT/sanitize_repro_racecheck_$i.log) race reports; $(grep -E 'tmem base|RACECHECK SUMMARY' $OUT/sanitize_repro_racecheck_$i.log | tr '\n' ' ')"
done
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 200 python tools/sanitize_small.py > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"
  python tools/sanitize_allow.py $tool $OUT/sanitize_$tool.log || rc=1
done
exit $rc
