timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 16 --warmup 3 2> gpurun_out/bench_2gpu.err | tail -1 > gpurun_out/bench_r01_2gpu_v2.json
cut -c1-1500 gpurun_out/bench_r01_2gpu_v2.json; tail -3 gpurun_out/bench_2gpu.err
