python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 48 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_v3.json; cat gpurun_out/bench_r01_v3.json | cut -c1-1500
