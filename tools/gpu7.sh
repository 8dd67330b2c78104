python -m pytest tests -q -m gpu 2>&1 | tail -8
python tools/fmm_check.py 16777216 3 1 2>&1 | head -5
