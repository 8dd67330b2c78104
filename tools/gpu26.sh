python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_r01_v5.json 2> gpurun_out/bench_r01_v5.err; tail -c 3000 gpurun_out/bench_r01_v5.json; tail -5 gpurun_out/bench_r01_v5.err
