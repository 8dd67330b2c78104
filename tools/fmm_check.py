"""Scratch driver: one FMM evaluation on the GPU vs the oracle (and the compiled reference when present)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import coulomb_oscillators_b200 as nb
from refs import Oracle, Ref, mean_rel_err

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m2l_first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dist = sys.argv[4] if len(sys.argv) > 4 else "ga"
st = nb.init_ga(n) if dist == "ga" else nb.init_test_cube(n)
par = nb.default_param(n)
ctx = nb.Context(order=p, unsort=0, m2l_first=m2l_first)
pos = st[0].copy(); vel = st[1].copy()
t0 = time.time(); acc = ctx.eval_host(nb.EVAL_FMM3_KD, pos, vel, par); t1 = time.time()
info = ctx.fmm_info()
print(f"n={n} p={p} m2l_first={m2l_first} L={info.levels} p2p={info.p2p_pairs} m2l={info.m2l_pairs} first {t1-t0:.3f}s launches {info.kernel_launches}")
print("  rebuild phases ms:", {k: round(v, 3) for k, v in ctx.fmm_phase_ms().items()})
T = ctx.fmm_tree(); P, M = ctx.fmm_lists()
# timing of a full tree_steps cycle on the device (state already in tree order): 1 rebuild + 7 reuse
import torch
buf = torch.from_numpy(np.concatenate([pos, vel, acc]).copy()).cuda()
dpar = torch.from_numpy(par).cuda()
ctx2 = nb.Context(order=p, unsort=0, m2l_first=m2l_first)
stt = torch.cuda.ExternalStream(ctx2.stream)
ctx2.compute_force(nb.EVAL_FMM3_KD, buf.data_ptr(), n, dpar.data_ptr())
for rep in range(2):
    tot = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stt)
    for k in range(8):
        ctx2.compute_force(nb.EVAL_FMM3_KD, buf.data_ptr(), n, dpar.data_ptr())
        ph = ctx2.fmm_phase_ms()
        for kk, v in ph.items(): tot[kk] = tot.get(kk, 0) + v
        if k == 1: reuse = dict(ph)
    e1.record(stt); e1.synchronize()
ms = e0.elapsed_time(e1) / 8
print(f"  8-eval cycle: {ms:.3f} ms/eval -> {n/ms/1e3:.1f} M particle-evals/s; per-eval phase avg ms:", {k: round(v/8, 3) for k, v in tot.items()})
print("  reuse-eval phases ms:", {k: round(v, 3) for k, v in reuse.items()})
if n <= (1 << 21):
    orc = Oracle(order=p, unsort=0, m2l_first=m2l_first)
    opos = st[0].copy(); ovel = st[1].copy()
    t0 = time.time(); oacc = orc.fmm3_kd(opos, ovel, par); t1 = time.time()
    OT = orc.tree(); OP, OM = orc.lists()
    print(f"  oracle {t1-t0:.2f}s lists {len(OP)},{len(OM)}")
    print("  perm", np.array_equal(T["perm"], OT["perm"]), "pos", np.array_equal(pos, opos), "vel", np.array_equal(vel, ovel))
    for k in ("lbound", "rbound", "center", "mult", "index", "splitdim"):
        eq = np.array_equal(T[k], OT[k])
        print("  ", k, eq, "" if eq else f"first diff at {np.argwhere(T[k] != OT[k])[:3].tolist()}")
    print("   p2p", np.array_equal(P, OP), "m2l", np.array_equal(M, OM))
    for k in ("mpole", "local"):
        print("  ", k, "max|diff|/max", float(np.abs(T[k] - OT[k]).max() / max(np.abs(OT[k]).max(), 1e-30)))
    print("   acc vs oracle: mean %.3e max %.3e" % mean_rel_err(acc, oacc))
    if dist == "ga" and n <= 65536:
        d = Oracle().direct3(st[0], par)
        dd = d[T["perm"]]
        print("   FMM vs direct: ours %.4f oracle %.4f" % (mean_rel_err(acc, dd)[0], mean_rel_err(oacc, dd)[0]))
