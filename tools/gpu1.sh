set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; free -g | head -2
for v in 0 1 2 3 4 5; do NBCO_DIRECT_VARIANT=$v python tools/direct_sweep.py 32768; done
for v in 0 2 3 4; do NBCO_DIRECT_VARIANT=$v python tools/direct_sweep.py 524288; done
