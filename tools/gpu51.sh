timeout 300 ncu --set full --clock-control none --import-source on -k regex:"m2l_kernel|l2p_kernel|leaf_p2m_kernel" -s 3 -c 3 -o gpurun_out/prof_m2l_l2p python tools/fmm_once.py 16777216 > gpurun_out/ncu_m2l.log 2>&1
ls -la gpurun_out/prof_m2l_l2p.ncu-rep
