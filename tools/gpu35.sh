timeout 100 python tools/trav_ab.py 2>&1 | tail -4
NBCO_TRAVERSE=queue timeout 250 python -m pytest tests/test_fmm_gpu.py -x -q --timeout 60 2>&1 | tail -3
NBCO_TRAVERSE=queue timeout 90 python tools/fmm_check.py 16777216 3 1 2>&1 | sed -n 3,3p
NBCO_TRAVERSE=queue timeout 90 python tools/fmm_check.py 1048576 3 1 2>&1 | sed -n 3,3p
NBCO_TRAVERSE=rounds timeout 90 python tools/fmm_check.py 1048576 3 1 2>&1 | sed -n 3,3p
