#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fmm_gpu.py -m gpu -q -x -k "fixture or uniform_leaves or shallow or leaf_level" > gpurun_out/r2w.log 2>&1
echo "rc=$?" >> gpurun_out/r2w.log
tail -n 25 gpurun_out/r2w.log | cut -c1-250
