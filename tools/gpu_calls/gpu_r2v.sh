#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fmm_gpu.py -m gpu -q -x -k "shallow" > gpurun_out/r2v_shallow.log 2>&1
echo "rc=$?" >> gpurun_out/r2v_shallow.log
timeout 900 python -m pytest tests/test_fmm_gpu.py tests/test_peer_gpu.py -m gpu -q -x -k "not shallow and not headline" > gpurun_out/r2v_fmm.log 2>&1
echo "rc=$?" >> gpurun_out/r2v_fmm.log
python tools/ab_phases.py 16777216 3 > gpurun_out/r2v_ab.json 2> gpurun_out/r2v.err
tail -n 30 gpurun_out/r2v_shallow.log | cut -c1-300; tail -n 3 gpurun_out/r2v_fmm.log; cat gpurun_out/r2v_ab.json
