timeout 900 python -m pytest tests/test_gpu_net.py -x -q -k "inception" -s 2>&1 | grep -v "^$" | tail -25
timeout 900 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -3
timeout 300 python tools/profile_step.py --games 4096 --playouts 8 --arch inception --blocks 10
