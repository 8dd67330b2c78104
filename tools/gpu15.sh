python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/bench2.log 2>&1
grep -v "Warning\|^\*\*\*\|OMP_NUM" gpurun_out/bench2.log | tail -25 | cut -c1-1800
