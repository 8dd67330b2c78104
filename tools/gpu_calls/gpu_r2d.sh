#!/bin/bash
# round 2, GPU call D: ncu on the kd build kernels, exported to CSV on the box (reports stay there: size limit)
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python tools/fmm_once.py 16777216 > gpurun_out/r2d_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kd_bottom" -c 1 -o /tmp/ncu/kdb python tools/fmm_once.py 16777216 > gpurun_out/r2d_ncu_b.log 2>&1
ncu -i /tmp/ncu/kdb.ncu-rep --page raw --csv > gpurun_out/r2d_kd_bottom_raw.csv 2>/dev/null
ncu -i /tmp/ncu/kdb.ncu-rep --page source --csv > gpurun_out/r2d_kd_bottom_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"top_|pack_bbox" -c 9 -o /tmp/ncu/kdt python tools/fmm_once.py 16777216 > gpurun_out/r2d_ncu_t.log 2>&1
ncu -i /tmp/ncu/kdt.ncu-rep --page raw --csv > gpurun_out/r2d_kd_top_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/ncu; du -sh gpurun_out
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q > gpurun_out/r2d_dropin.log 2>&1; tail -5 gpurun_out/r2d_dropin.log
