python -m pytest tests/test_fmm_gpu.py -q 2>&1 | tail -4
for r in 44 24 20 16 12; do echo "NBCO_TRAV_ROUNDS=$r"; NBCO_TRAV_ROUNDS=$r python tools/fmm_check.py 16777216 3 1 2>&1 | sed -n 3,3p; done
echo default; python tools/fmm_check.py 16777216 3 1 2>&1 | head -4
for r in 40 18 14 10; do echo "1M NBCO_TRAV_ROUNDS=$r"; NBCO_TRAV_ROUNDS=$r python tools/fmm_check.py 1048576 3 1 2>&1 | sed -n 3,3p; done
