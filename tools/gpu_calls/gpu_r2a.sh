#!/bin/bash
# round 2, GPU call A: full GPU test suite, A/B phase times (round-1 build vs current), reference GPU path, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_gpu.txt
free -g | head -2 >> gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -x --durations=15 > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
for cfg in "1048576 5" "1048576 4" "16777216 3"; do
  NBCO_LIB=$PWD/build/r01/libnbco.so timeout 300 python tools/ab_phases.py $cfg >> gpurun_out/r2a_ab.log 2>&1
  timeout 300 python tools/ab_phases.py $cfg >> gpurun_out/r2a_ab.log 2>&1
done
timeout 240 python tools/ref_gpu_baseline.py 1048576 16 > gpurun_out/r2a_refgpu.log 2>&1
echo "refgpu rc=$?" >> gpurun_out/r2a_refgpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_ab.log; cat gpurun_out/r2a_refgpu.log; tail -c 1500 gpurun_out/r2a_bench.err
