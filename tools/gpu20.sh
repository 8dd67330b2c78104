python -m pytest tests -q -m gpu 2>&1 | tail -4
python tools/fmm_check.py 1048576 5 1 2>&1 | head -4
python tools/fmm_check.py 1048576 4 1 2>&1 | head -4
python tools/fmm_check.py 1048576 6 1 2>&1 | head -4
