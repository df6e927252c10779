timeout 900 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -3
timeout 300 python tools/profile_step.py --games 4096 --playouts 8 --arch resnet --blocks 10 --precision split
timeout 300 python tools/profile_step.py --games 4096 --playouts 8 --arch resnet --blocks 10 --precision fp16
timeout 600 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r1i.json 2>gpurun_out/r1i.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r1i.json'));r=d['roofline'];print(d['value'], r['frac'], [round(x,4) for x in r['phase_ms_per_lockstep']['trunk_convs']], r['phase_ms_per_lockstep']['heads'], d['clocks'])"
