"""One 2D FMM evaluation + phase times (profiling helper): python tools/fmm2_once.py [n] [order] [kv|ga] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
order = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dist = sys.argv[3] if len(sys.argv) > 3 else "kv"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
st = nb.init_kv2(n) if dist == "kv" else nb.init_ga2(n)
par = torch.from_numpy(nb.default_param2(n)).cuda()
buf = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))])).cuda()
ctx = nb.Context(order=order)
for r in range(reps):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    ctx.compute_force2(nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, par.data_ptr())
    t1.record(); torch.cuda.synchronize()
    print(f"n={n} p={order} {dist} L={ctx.fmm2_info().levels} eval {t0.elapsed_time(t1):.3f} ms phases", {k: round(v, 3) for k, v in ctx.fmm2_phase_ms().items()})
T = ctx.fmm2_tree()
L = T["levels"]; m = T["mult"][(4 ** L - 1) // 3:]
side = 1 << L
mm = m.reshape(side, side).astype(np.int64)
pad = np.pad(mm, 1)
nb9 = sum(pad[1 + di:1 + di + side, 1 + dj:1 + dj + side] for di in (-1, 0, 1) for dj in (-1, 0, 1))
inter = int((mm * nb9).sum())
print("leaf mult: max", m.max(), "mean nonempty", m[m > 0].mean(), "empty frac", (m == 0).mean(), "near interactions", inter, "per particle", inter / n)
ms = ctx.fmm2_phase_ms()["near_l2p"]
print(f"near field: {inter / ms / 1e6:.1f} G interactions/s")
