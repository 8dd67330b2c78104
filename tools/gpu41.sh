timeout 300 python -m pytest tests/test_fmm_gpu.py -x -q --timeout 100 -k "incremental or reuse" 2>&1 | tail -12
timeout 90 python tools/fmm_check.py 16777216 3 1 2>&1 | sed -n 2,4p
