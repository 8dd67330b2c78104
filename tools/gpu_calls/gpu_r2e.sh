#!/bin/bash
# round 2, GPU call E: by-target downward pass + leaner kd_bottom: parity tests, A/B phase times, bench
mkdir -p gpurun_out
timeout 300 python tools/fmm_check.py 1048576 3 1 > gpurun_out/r2e_check1m.log 2>&1
NBCO_PAIR_LISTS=1 timeout 300 python tools/fmm_check.py 1048576 3 1 > gpurun_out/r2e_check1m_pairs.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
for cfg in "16777216 3" "1048576 3" "1048576 5"; do
  timeout 300 python tools/ab_phases.py $cfg >> gpurun_out/r2e_ab.log 2>&1
  NBCO_PAIR_LISTS=1 timeout 300 python tools/ab_phases.py $cfg >> gpurun_out/r2e_ab.log 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?" >> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_check1m.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2e_pytest.log | tail -15; cat gpurun_out/r2e_ab.log; tail -c 600 gpurun_out/r2e_bench.err
