NBCO_DEBUG_TRAV=1 timeout 120 python tools/trav_dbg.py 16777216 2>&1 | tail -22
