python -m pytest tests/test_fmm_gpu.py -q -x 2>&1 | tail -12
python tools/fmm_check.py 16777216 3 1 2>&1 | head -5
python tools/fmm_check.py 1048576 3 1 2>&1 | head -5
python tools/fmm_once.py 16777216 > gpurun_out/once.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fmm16m_v2.csv python tools/fmm_once.py 16777216 > gpurun_out/ncu_once.log 2>&1
