NBCO_DEBUG_TRAV=1 timeout 120 python tools/trav_dbg.py 16777216 2>&1 | tail -6
timeout 300 python -m pytest tests/test_fmm_gpu.py -x -q --timeout 100 -k "incremental or reuse" 2>&1 | tail -3
