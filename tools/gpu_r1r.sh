#!/bin/bash
for net in resnet inception; do
timeout 900 python bench.py --net $net --steps 2 --warmup 3 > gpurun_out/r1r_bench_$net.json 2> gpurun_out/r1r_bench_$net.err; echo "bench $net rc=$?"; cat gpurun_out/r1r_bench_$net.json | cut -c1-2500; tail -3 gpurun_out/r1r_bench_$net.err
done
