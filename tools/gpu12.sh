python -m pytest tests -q -m gpu 2>&1 | tail -4
python tools/fmm_check.py 16777216 3 1 2>&1 | head -4
python tools/fmm_check.py 1048576 3 1 2>&1 | head -4
python tools/fmm_once.py 16777216 > gpurun_out/once.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fmm16m_v5.csv python tools/fmm_once.py 16777216 > gpurun_out/ncu_once.log 2>&1
