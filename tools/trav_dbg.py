import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
st = nb.init_ga(n)
buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda"); buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
par = torch.from_numpy(nb.default_param(n)).cuda()
ctx = nb.Context(order=3, unsort=0, tree_steps=8)
ctx.compute_force(nb.EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, par.data_ptr())
for s in range(9):
    ctx.integrate(nb.LEAPFROG, nb.EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, par.data_ptr(), 5e-4, 1)
    print("step", s, "traverse ms", round(ctx.fmm_phase_ms()["traverse"], 3), flush=True)
