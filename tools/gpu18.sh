python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 16 --warmup 3 > gpurun_out/bench8.log 2>&1
grep -v "Warning\|^\*\*\*\|OMP_NUM" gpurun_out/bench8.log | tail -4 | cut -c1-2600
