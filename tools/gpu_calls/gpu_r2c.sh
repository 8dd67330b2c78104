#!/bin/bash
# round 2, GPU call C: changed-test subset, then ncu on the kd build kernels (full set + launch list)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_direct_gpu.py tests/test_integrate_gpu.py tests/test_fmm_gpu.py tests/test_peer_gpu.py -m gpu -q --maxfail=10 --durations=8 -s > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 300 python tools/ab_phases.py 16777216 3 > gpurun_out/r2c_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kd_bottom|top_|pack_bbox" -o gpurun_out/r2c_kd python tools/fmm_once.py 16777216 > gpurun_out/r2c_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python tools/fmm_once.py 16777216 > gpurun_out/r2c_ncu2.log 2>&1
grep -E "passed|failed|N=2\^24|vs fp64" gpurun_out/r2c_pytest.log | tail -12; cat gpurun_out/r2c_ab.log; tail -3 gpurun_out/r2c_ncu.log; ls -la gpurun_out/r2c_kd.ncu-rep
