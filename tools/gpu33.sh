timeout 300 python -m pytest tests/test_fmm_gpu.py tests/test_integrate_gpu.py tests/test_cli_gpu.py -x -q --timeout 60 2>&1 | tail -4
timeout 90 python tools/fmm_check.py 16777216 3 1 2>&1 | head -5
