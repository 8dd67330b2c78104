#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2; do NBCO_NEAR_MODE=$m python tools/ab_phases.py 16777216 3 > gpurun_out/r2r_ab_mode$m.json 2>> gpurun_out/r2r.err; done
cat gpurun_out/r2r_ab_mode*.json; tail -n 3 gpurun_out/r2r.err
