"""Scratch: two ranks emulated on one device (host threads + nbco_peer_attach_local), one evaluation, stage reports
of the distributed kd build (NBCO_DEBUG_KD=1)."""
import os, sys, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ["NBCO_TRAVERSE"] = "launches"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100003
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
st = nb.init_ga(n)
par = torch.from_numpy(nb.default_param(n)).cuda()
EV = nb.EVAL_COULOMB_FMM3_KD
def state():
    b = torch.zeros(9 * n, dtype=torch.float32, device="cuda"); b[:6 * n] = torch.from_numpy(st.ravel()).cuda(); return b
c1 = nb.Context(order=3, unsort=0, tree_steps=4); b1 = state()
c1.compute_force(EV, b1.data_ptr(), n, par.data_ptr())
want = b1.cpu().numpy().reshape(3, n, 3)
ctxs = [nb.Context(order=3, unsort=0, tree_steps=4, rank=r, world=world) for r in range(world)]
bufs = [state() for _ in range(world)]
torch.cuda.synchronize()
for c in ctxs: c.peer_export(n)
for r, c in enumerate(ctxs):
    for q in range(world):
        if q != r: c.peer_attach_local(q, ctxs[q])
    c.peer_commit()
errs = [None] * world
def main(r):
    try:
        torch.cuda.set_device(0)
        ctxs[r].compute_force(EV, bufs[r].data_ptr(), n, par.data_ptr())
        ctxs[r].peer_gather(bufs[r].data_ptr(), n)
    except Exception as e:
        errs[r] = e
th = [threading.Thread(target=main, args=(r,)) for r in range(world)]
[t.start() for t in th]; [t.join() for t in th]
print("errors:", errs)
T1 = c1.fmm_tree()
L = T1["levels"]
for r in range(world):
    try:
        T = ctxs[r].fmm_tree()
    except Exception as e:
        print("rank", r, "tree unreadable:", e); continue
    lo, hi = nb.shard_range(n, r, world)
    print(f"rank {r}: perm own range equal {np.array_equal(T['perm'][lo:hi], T1['perm'][lo:hi])}  (first 5 ours {T['perm'][lo:lo+5]} want {T1['perm'][lo:lo+5]})")
    for l in range(0, min(L, 7) + 1):
        b, e = (1 << l) - 1, (1 << (l + 1)) - 1
        msg = []
        for k in ("lbound", "rbound", "splitdim"):
            eq = (T[k][b:e] == T1[k][b:e])
            eq = eq.reshape(e - b, -1).all(1)
            msg.append(f"{k} {int(eq.sum())}/{e - b}")
        print(f"   level {l}: " + "  ".join(msg))
if not any(errs):
    for r in range(world):
        got = bufs[r].cpu().numpy().reshape(3, n, 3)
        print(r, "pos equal", np.array_equal(got[0], want[0]), "vel equal", np.array_equal(got[1], want[1]), "acc max rel", float(np.abs(got[2] - want[2]).max() / np.abs(want[2]).max()))
