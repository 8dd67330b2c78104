#!/bin/bash
# round 2, final single-GPU call: full GPU suite, smoke, bench (both arms), ncu launch list + full pages of the top kernels
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --durations=6 > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2s_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?" >> gpurun_out/r2s_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err
echo "ref rc=$?" >> gpurun_out/r2s_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches.csv python tools/fmm_once.py 16777216 > gpurun_out/r2s_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"m2l_kernel|l2lp_uniform|leaf_p2m_u8|reval_kernel|emit_kernel|p2p_kernel" -c 8 -o /tmp/ncu/ev python tools/fmm_once.py 16777216 > gpurun_out/r2s_ncu2.log 2>&1
ncu -i /tmp/ncu/ev.ncu-rep --page raw --csv > gpurun_out/r2s_eval_kernels_raw.csv 2>/dev/null
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2s_pytest.log | tail; tail -2 gpurun_out/r2s_smoke.log; tail -c 300 gpurun_out/r2s_bench.err; head -c 300 gpurun_out/r2s_bench_ref.json; tail -c 200 gpurun_out/r2s_bench_ref.err; du -sh gpurun_out
