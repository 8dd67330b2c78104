timeout 120 python tools/fmm_check.py 16777216 3 1 2>&1 | head -5
