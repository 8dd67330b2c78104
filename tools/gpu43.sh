timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_r01_v6.json 2> gpurun_out/bench_r01_v6.err; cut -c1-400 gpurun_out/bench_r01_v6.json; tail -2 gpurun_out/bench_r01_v6.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fmm16m_v6.csv python tools/fmm_once.py 16777216 > gpurun_out/ncu_once6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:near_l2p2 -c 1 -o gpurun_out/prof_near2d python tools/fmm2_once.py 4194304 5 kv 1 > gpurun_out/ncu_near2d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:l2l_level_kernel -s 4 -c 1 -o gpurun_out/prof_l2l python tools/fmm_once.py 16777216 > gpurun_out/ncu_l2l.log 2>&1
ls -la gpurun_out/*.ncu-rep
