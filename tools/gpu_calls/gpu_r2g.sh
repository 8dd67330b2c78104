#!/bin/bash
# round 2, GPU call G: full GPU suite, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?" >> gpurun_out/r2g_bench.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2g_pytest.log | tail -15; tail -c 400 gpurun_out/r2g_bench.err; head -c 600 gpurun_out/r2g_bench.json
