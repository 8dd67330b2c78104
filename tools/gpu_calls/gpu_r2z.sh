#!/bin/bash
# records of the incremental traversal outgrowing their array in long runs: device-side fallback to the root
mkdir -p gpurun_out
timeout 100 python tools/drift3d.py 1048576 3 200 5e-4 8 > gpurun_out/r2z_drift3d_1m.json 2> gpurun_out/r2z.err
timeout 400 python -m pytest tests/test_fmm_gpu.py tests/test_peer_gpu.py -m gpu -q -x -k "incremental or reuse or peer" > gpurun_out/r2z_fmm.log 2>&1
echo "rc=$?" >> gpurun_out/r2z_fmm.log
python tools/ab_phases.py 16777216 3 > gpurun_out/r2z_ab.json 2>> gpurun_out/r2z.err
cat gpurun_out/r2z_drift3d_1m.json; tail -n 3 gpurun_out/r2z_fmm.log; cat gpurun_out/r2z_ab.json; tail -n 3 gpurun_out/r2z.err
