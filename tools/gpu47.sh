timeout 600 python -m pytest tests/test_cli2d_gpu.py tests/test_cli_gpu.py -x -q --timeout 300 2>&1 | tail -12
