#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
timeout 300 python tools/ab_phases.py 16777216 3 > gpurun_out/r2j_ab.log 2>&1
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2j_pytest.log | tail; cat gpurun_out/r2j_ab.log
