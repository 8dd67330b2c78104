timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 48 --warmup 3 2> gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_r01_8gpu_v3.json
cut -c1-300 gpurun_out/bench_r01_8gpu_v3.json; grep -v "^W\|OMP\|\*\*\*" gpurun_out/bench_8gpu.err | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 48 --warmup 3 --no-direct 2> /dev/null | tail -1 > gpurun_out/bench_r01_4gpu_v3.json
cut -c1-300 gpurun_out/bench_r01_4gpu_v3.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 48 --warmup 3 --no-direct 2> /dev/null | tail -1 > gpurun_out/bench_r01_2gpu_v3.json
cut -c1-300 gpurun_out/bench_r01_2gpu_v3.json
