timeout 100 python tools/trav_ab.py
NBCO_TRAVERSE=rounds timeout 100 python tools/trav_ab.py
