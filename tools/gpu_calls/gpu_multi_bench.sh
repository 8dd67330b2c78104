#!/bin/bash
# bench only, at every rank count given (one box with >= max GPUs)
mkdir -p gpurun_out
for N in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2u_bench_$N.json 2> gpurun_out/r2u_bench_$N.err
  echo "bench rc=$?" >> gpurun_out/r2u_bench_$N.err
  tail -n 1 gpurun_out/r2u_bench_$N.err; tail -n 1 gpurun_out/r2u_bench_$N.json | head -c 330; echo
done
