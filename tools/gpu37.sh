timeout 300 python -m pytest tests/test_peer_gpu.py -x -q --timeout 120 --durations=5 2>&1 | tail -14
