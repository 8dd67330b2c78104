"""A/B of the two traversal kernels (NBCO_TRAVERSE=rounds|queue): force error statistics and time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import coulomb_oscillators_b200 as nb
from refs import Oracle, mean_rel_err
n, order = 65536, 5
st = nb.init_test_cube(n); par = nb.default_param(n)
orc = Oracle(order=order, unsort=0, m2l_first=1)
opos = st[0].copy(); oacc = orc.fmm3_kd(opos, st[1].copy(), par)
for rep in range(4):
    ctx = nb.Context(order=order, unsort=0, m2l_first=1)
    pos, vel = st[0].copy(), st[1].copy()
    acc = ctx.eval_host(nb.EVAL_FMM3_KD, pos, vel, par)
    P, M = ctx.fmm_lists(); OP, OM = orc.lists()
    print(os.environ.get("NBCO_TRAVERSE", "queue"), rep, "lists equal", np.array_equal(P, OP) and np.array_equal(M, OM), "err", mean_rel_err(acc, oacc),
          "traverse ms", round(ctx.fmm_phase_ms()["traverse"], 3))
