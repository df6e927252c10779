AP_CONV4_128=2 timeout 900 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -3
for c in 0 1 2; do
AP_CONV4_128=$c timeout 600 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r1s_c$c.json 2>gpurun_out/r1s_c$c.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r1s_c$c.json'));r=d['roofline'];print($c, d['value'], r['frac'], [round(x,4) for x in r['phase_ms_per_lockstep']['trunk_convs']], d['clocks'])"
done
for c in 0 2; do AP_CONV4_128=$c timeout 300 python tools/profile_step.py --games 4096 --playouts 8 --arch resnet --blocks 10 --precision fp16; done
