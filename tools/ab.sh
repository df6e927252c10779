for g in 1 2; do timeout 600 python bench.py --workload selfplay --steps 6 --warmup 2 --groups $g 2>&1 | tail -1; done
