#!/bin/bash
# final check of the round: the driver's own commands (GPU tests with -x, smoke, bench)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2x_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
echo "bench rc=$?" >> gpurun_out/r2x_bench.err
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2x_pytest.log | tail -n 5; tail -n 2 gpurun_out/r2x_smoke.log; tail -c 200 gpurun_out/r2x_bench.err; head -c 250 gpurun_out/r2x_bench.json
