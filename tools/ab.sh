timeout 600 python -m pytest tests/test_gpu_net.py -x -q 2>&1 | tail -3
for hp in 0 1; do
AP_HEAD_PAIR=$hp timeout 600 python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/r1h_hp$hp.json 2>gpurun_out/r1h_hp$hp.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r1h_hp$hp.json'));r=d['roofline'];print($hp, d['value'], r['frac'], [round(x,4) for x in r['phase_ms_per_lockstep']['trunk_convs']], r['phase_ms_per_lockstep']['heads'], d['clocks'])"
done
