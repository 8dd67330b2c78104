timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py > gpurun_out/bench_r01_final_1gpu.json 2> gpurun_out/bench_final.err; cut -c1-330 gpurun_out/bench_r01_final_1gpu.json; tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 0 2>/dev/null | cut -c1-400
