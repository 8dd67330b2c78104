timeout 400 python -m pytest tests/test_fmm_gpu.py tests/test_peer_gpu.py -x -q --timeout 100 2>&1 | tail -3
timeout 90 python tools/fmm_check.py 16777216 3 1 2>&1 | sed -n 3,3p
timeout 90 python tools/fmm_check.py 1000003 3 1 2>&1 | sed -n 3,3p
