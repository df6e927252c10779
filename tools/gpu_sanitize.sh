#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (tools/sanitize_small.py).
# racecheck reports the tcgen05.alloc shared-memory slot of the CTA-pair kernels (written by the tensor-core unit,
# published by a cluster barrier the tool does not model) - known false positive.  synccheck stops at the fused-head
# epilogue of the conv kernels: the two warps that share a named barrier (bar.sync id, 64) reach it from two inlined
# copies of the same code (column half as a compile-time constant), which the PTX rules allow (alignment is per warp)
# and the tool reports as divergence.  memcheck is clean; racecheck is clean for the board / tree / rollout kernels.
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small ok|Barrier error|Race reported" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
