timeout 900 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -5
for c4 in 1 2 1 2; do
AP_CONV4=$c4 timeout 600 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r1p_c$c4.json 2>gpurun_out/r1p_c$c4.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r1p_c$c4.json'));r=d['roofline'];print($c4, d['value'], r['frac'], [round(x,4) for x in r['phase_ms_per_lockstep']['trunk_convs']], d['clocks'])"
done
