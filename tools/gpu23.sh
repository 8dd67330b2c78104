python -m pytest tests/test_fmm2_gpu.py -x -q 2>&1 | tail -5
python tools/fmm2_once.py 4194304 5 kv
python tools/fmm2_once.py 4194304 5 ga
python tools/fmm2_once.py 1048576 3 kv
