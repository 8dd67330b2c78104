#!/bin/bash
# round 2: A/B of the bottom kd kernel's capacity / CTA shape (build/var libraries), then the FMM tests with the default build
mkdir -p gpurun_out
for v in 4096_512 4096_256; do
  NBCO_LIB=$PWD/build/var/libnbco_b$v.so python tools/ab_phases.py 16777216 3 > gpurun_out/r2o_ab_$v.json 2>> gpurun_out/r2o.err
  NBCO_LIB=$PWD/build/var/libnbco_b$v.so python tools/ab_phases.py 1048576 3 > gpurun_out/r2o_ab_${v}_1m.json 2>> gpurun_out/r2o.err
done
python tools/ab_phases.py 16777216 3 > gpurun_out/r2o_ab_default.json 2>> gpurun_out/r2o.err
NBCO_LIB=$PWD/build/var/libnbco_b4096_512.so timeout 600 python -m pytest tests/test_fmm_gpu.py -m gpu -q -x -k "oracle or equal_keys or leaf_level" > gpurun_out/r2o_fmm_var.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_fmm_var.log
timeout 600 python -m pytest tests/test_fmm_gpu.py -m gpu -q -x -k "leaf_level" > gpurun_out/r2o_fmm.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_fmm.log
cat gpurun_out/r2o_ab_*.json; tail -3 gpurun_out/r2o_fmm_var.log gpurun_out/r2o_fmm.log; cat gpurun_out/r2o.err | tail -5
