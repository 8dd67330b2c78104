timeout 60 python tools/fmm_check.py 1000 3 1 2>&1 | head -5; echo "rc $?"
timeout 60 python tools/fmm_check.py 100000 3 0 2>&1 | head -5; echo "rc $?"
timeout 90 python tools/fmm_check.py 16777216 3 1 2>&1 | head -5
timeout 250 python -m pytest tests/test_fmm_gpu.py -x -q --timeout 40 2>&1 | tail -5
