python -m pytest tests -q -m gpu 2>&1 | tail -4
python tools/fmm_check.py 16777216 3 1 2>&1 | head -4
python bench.py --steps 48 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_v4.json; cut -c1-2200 gpurun_out/bench_r01_v4.json
