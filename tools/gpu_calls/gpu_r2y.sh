#!/bin/bash
# config 3 at its stated size: PEFRL energy drift of the 2D fp64 FMM, N = 2^22 (and a short run at 2^18 to size the O(N^2) energy)
mkdir -p gpurun_out
timeout 120 python tools/drift2d.py 262144 5 200 > gpurun_out/r2y_drift_256k.json 2> gpurun_out/r2y.err
timeout 300 python tools/drift2d.py 4194304 5 200 > gpurun_out/r2y_drift_4m.json 2>> gpurun_out/r2y.err
cat gpurun_out/r2y_drift_*.json; tail -n 3 gpurun_out/r2y.err
