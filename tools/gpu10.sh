python -m pytest tests -q -m gpu 2>&1 | tail -4
python tools/fmm_check.py 16777216 3 1 2>&1 | head -4
python tools/fmm_once.py 16777216 > gpurun_out/once.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:kd_bottom -c 1 -o gpurun_out/prof_bottom python tools/fmm_once.py 16777216 > gpurun_out/ncu_bottom.log 2>&1
tail -3 gpurun_out/ncu_bottom.log
