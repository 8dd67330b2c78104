timeout 300 python -m pytest tests/test_integrate_gpu.py -x -q --timeout 200 2>&1 | tail -8
timeout 400 python bench.py --no-fmm2d > gpurun_out/bench_r01_v7.json 2> gpurun_out/bench_r01_v7.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01_v7.json').read())
print('value %.4g ms/step %.3f e2e %.4g launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print({k:v['avg_ms'] for k,v in d['phases'].items()})
PY
