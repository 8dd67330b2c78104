python -m pytest tests/test_fmm2_gpu.py -x -q 2>&1 | tail -5
python tools/fmm2_once.py 4194304 5 kv 3 2>&1 | tail -4
python tools/fmm2_once.py 16777216 5 kv 2 2>&1 | tail -3
./build/dfma_peak
