#!/bin/bash
# round 2: leaf-level fusion A/B + hunt for the emulated-ranks barrier time-out seen only inside the full suite
mkdir -p gpurun_out
python tools/ab_phases.py 16777216 3 > gpurun_out/r2n_ab_fused.json 2> gpurun_out/r2n_ab.err
NBCO_NO_LEAF_FUSION=1 python tools/ab_phases.py 16777216 3 > gpurun_out/r2n_ab_unfused.json 2>> gpurun_out/r2n_ab.err
python tools/ab_phases.py 1048576 3 > gpurun_out/r2n_ab_fused_1m.json 2>> gpurun_out/r2n_ab.err
NBCO_NO_LEAF_FUSION=1 python tools/ab_phases.py 1048576 3 > gpurun_out/r2n_ab_unfused_1m.json 2>> gpurun_out/r2n_ab.err
export NBCO_PEER_TIMEOUT_S=4
NBCO_TEST_POISON=60 NBCO_DEBUG_KD=1 timeout 600 python -m pytest tests/test_peer_gpu.py -m gpu -q -x > gpurun_out/r2n_peer_poison.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_peer_poison.log
NBCO_DEBUG_KD=1 timeout 900 python -m pytest tests/test_integrate_gpu.py tests/test_multigpu_gpu.py tests/test_peer_gpu.py -m gpu -q -x > gpurun_out/r2n_peer_after_integrate.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_peer_after_integrate.log
unset NBCO_PEER_TIMEOUT_S
timeout 900 python -m pytest tests/test_fmm_gpu.py -m gpu -q -x > gpurun_out/r2n_fmm.log 2>&1
echo "rc=$?" >> gpurun_out/r2n_fmm.log
cat gpurun_out/r2n_ab_*.json; tail -5 gpurun_out/r2n_peer_poison.log; tail -5 gpurun_out/r2n_peer_after_integrate.log; tail -5 gpurun_out/r2n_fmm.log
