python -m pytest tests/test_cli_gpu.py -q 2>&1 | tail -3
for v in 0 8 9 10 11 12 13; do NBCO_DIRECT_VARIANT=$v python tools/direct_sweep.py 1048576; done
