#!/bin/bash
mkdir -p gpurun_out
python tools/ab_phases.py 16777216 3 > gpurun_out/r2p_ab.json 2> gpurun_out/r2p.err
python tools/ab_phases.py 1048576 3 > gpurun_out/r2p_ab_1m.json 2>> gpurun_out/r2p.err
python tools/ab_phases.py 4194304 2 > gpurun_out/r2p_ab_4m_p2.json 2>> gpurun_out/r2p.err
timeout 900 python -m pytest tests/test_fmm_gpu.py tests/test_peer_gpu.py tests/test_integrate_gpu.py tests/test_dropin_gpu.py -m gpu -q -x > gpurun_out/r2p_fmm.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_fmm.log
cat gpurun_out/r2p_ab*.json; tail -n 4 gpurun_out/r2p_fmm.log; tail -n 5 gpurun_out/r2p.err
