timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r1m.json 2>gpurun_out/r1m.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r1m.json'));r=d['roofline'];print(d['value'], d['e2e']['value'], r['frac'], r['phase_ms_per_lockstep'], r['tree_kernels'], d['clocks'])"
