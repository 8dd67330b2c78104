# What a round-end validation on a B200 box runs (under /usr/local/graft/bin/gpurun -- 'bash tools/gpu_validate.sh'):
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; cut -c1-330 gpurun_out/bench_1gpu.json; tail -2 gpurun_out/bench_1gpu.err
# launch list and one full capture of the dominant kernels (profiles/r01_notes.md):
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/fmm_once.py 16777216
#   ncu --set full --clock-control none --import-source on -k regex:near_l2p2 -c 1 -o gpurun_out/prof_near2d python tools/fmm2_once.py 4194304 5 kv 1
# several GPUs (gpurun --gpus N):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py 300001 16777216 48
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus N
