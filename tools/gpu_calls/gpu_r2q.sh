#!/bin/bash
mkdir -p gpurun_out
for c in 8 12 16 24; do
  NBCO_M2L_CTAS=$c python tools/ab_phases.py 16777216 3 > gpurun_out/r2q_ab_$c.json 2>> gpurun_out/r2q.err
  NBCO_M2L_CTAS=$c python tools/ab_phases.py 1048576 5 > gpurun_out/r2q_ab_p5_$c.json 2>> gpurun_out/r2q.err
done
cat gpurun_out/r2q_ab*.json; tail -n 5 gpurun_out/r2q.err
