#!/bin/bash
# round 2, GPU call I: uniform-leaf fast paths + own-range zero-fill: tests, phase times, ncu page of the direct-sum kernel
mkdir -p gpurun_out /tmp/ncu
timeout 1200 python -m pytest tests/test_fmm_gpu.py tests/test_peer_gpu.py tests/test_integrate_gpu.py -m gpu -q --maxfail=8 > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
for cfg in "16777216 3" "1048576 3" "1048576 5"; do timeout 300 python tools/ab_phases.py $cfg >> gpurun_out/r2i_ab.log 2>&1; done
timeout 600 ncu --set full --clock-control none -k regex:"direct3_packed" -c 1 -o /tmp/ncu/dir python tools/direct_sweep.py > gpurun_out/r2i_ncu_dir.log 2>&1
ncu -i /tmp/ncu/dir.ncu-rep --page raw --csv > gpurun_out/r2i_direct_raw.csv 2>/dev/null
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2i_pytest.log | tail; cat gpurun_out/r2i_ab.log; tail -3 gpurun_out/r2i_ncu_dir.log
