python tools/fmm2_once.py 4194304 5 kv 2 > gpurun_out/f2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fmm2d_v1.csv python tools/fmm2_once.py 4194304 5 kv 2 > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/f2.log
