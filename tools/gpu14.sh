python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v Warning | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_2gpu.json; cut -c1-900 gpurun_out/bench_r01_2gpu.json
python bench.py --steps 24 --warmup 3 --no-direct 2>&1 | tail -1 | cut -c1-400
