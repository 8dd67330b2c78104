#!/bin/bash
# round 2, GPU call F: full GPU suite after the reproducible-mode switch; ncu source page of the new kd_bottom
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 300 python tools/ab_phases.py 16777216 3 > gpurun_out/r2f_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kd_bottom" -c 1 -o /tmp/ncu/kdb python tools/fmm_once.py 16777216 > gpurun_out/r2f_ncu_b.log 2>&1
ncu -i /tmp/ncu/kdb.ncu-rep --page raw --csv > gpurun_out/r2f_kd_bottom_raw.csv 2>/dev/null
ncu -i /tmp/ncu/kdb.ncu-rep --page source --csv > gpurun_out/r2f_kd_bottom_source.csv 2>/dev/null
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2f_pytest.log | tail -15; cat gpurun_out/r2f_ab.log
