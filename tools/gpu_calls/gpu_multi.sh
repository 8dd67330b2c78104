#!/bin/bash
# round 2, multi-GPU call: peer parity against the oracle (default traversal) + bench at $1 GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2i_gpus_$N.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/r2i_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2i_bench_$N.json 2> gpurun_out/r2i_bench_$N.err
echo "bench rc=$?" >> gpurun_out/r2i_bench_$N.err
tail -4 gpurun_out/r2i_pytest_$N.log; tail -c 600 gpurun_out/r2i_bench_$N.err; head -c 400 gpurun_out/r2i_bench_$N.json
