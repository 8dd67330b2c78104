python -m pytest tests/test_fmm_gpu.py -q 2>&1 | tail -4
python tools/fmm_check.py 1048576 5 1 2>&1 | head -4
