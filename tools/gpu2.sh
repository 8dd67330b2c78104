python tools/fmm_check.py 8192 3 0
python tools/fmm_check.py 20000 2 1
python tools/fmm_check.py 100003 3 1
python tools/fmm_check.py 50000 4 0
python tools/fmm_check.py 65536 5 1 cube
python tools/fmm_check.py 1048576 3 1
python tools/fmm_check.py 1048576 3 0
python tools/fmm_check.py 16777216 3 1
