#!/bin/bash
# round 2, GPU call B: new kd build (linear-histogram selection): sanitizer on a small case, kd parity tests, timings
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/fmm_check.py 30011 3 1 > gpurun_out/r2b_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2b_sanitizer.log
timeout 300 python tools/fmm_check.py 1048576 3 1 > gpurun_out/r2b_check1m.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --durations=10 > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python tools/ab_phases.py 16777216 3 >> gpurun_out/r2b_ab.log 2>&1
timeout 300 python tools/ab_phases.py 1048576 3 >> gpurun_out/r2b_ab.log 2>&1
tail -30 gpurun_out/r2b_sanitizer.log; cat gpurun_out/r2b_check1m.log; tail -15 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_ab.log
