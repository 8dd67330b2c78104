#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_gpu.py tests/test_fmm_gpu.py -m gpu -q --maxfail=6 > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 300 python tools/ab_phases.py 16777216 3 > gpurun_out/r2k_ab.log 2>&1
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2k_pytest.log | tail; cat gpurun_out/r2k_ab.log
