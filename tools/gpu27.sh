timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py 300001 16777216 16 2>&1 | grep -v "^W\|^\[W\|warn" | tail -30
