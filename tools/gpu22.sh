python -m pytest tests/test_fmm2_gpu.py -x -q 2>&1 | tail -30
