python -m pytest tests -q -m gpu 2>&1 | tail -4
python tools/fmm_check.py 16777216 3 1 2>&1 | head -4
python tools/fmm_check.py 1048576 3 1 2>&1 | head -4
for v in 0 6 7 4; do NBCO_DIRECT_VARIANT=$v python tools/direct_sweep.py 1048576; done
python tools/direct_sweep.py 262144 > gpurun_out/d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:direct3_packed -c 1 -o gpurun_out/prof_direct python tools/direct_sweep.py 262144 > gpurun_out/ncu_direct.log 2>&1
tail -2 gpurun_out/ncu_direct.log
