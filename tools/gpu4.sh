python -m pytest tests -q -m gpu 2>&1 | tail -8
python tools/fmm_once.py 16777216 > gpurun_out/once.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fmm16m.csv python tools/fmm_once.py 16777216 > gpurun_out/ncu_once.log 2>&1
tail -2 gpurun_out/once.log
