"""Run under torchrun (one process per GPU): the peer-memory FMM leapfrog (csrc/peer.cu) with the DEFAULT traversal
(cooperative + incremental) on WORLD_SIZE GPUs against the ORACLE (oracle/nbco_oracle.c, the CPU restatement pinned to
the reference) on the same inputs:  final positions / velocities, and the union of the per-rank interaction lists of
the last evaluation against the oracle's lists (bit-exact sets).  Prints PEER_PARITY OK|FAIL on rank 0; exit code 1 on FAIL.
   torchrun --nproc-per-node W tools/peer_parity.py [n] [steps] [tree_steps] [order]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import coulomb_oscillators_b200 as nb
from coulomb_oscillators_b200.parallel import peer_setup, fmm_leapfrog_peer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tree_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
order = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ev = nb.EVAL_COULOMB_FMM3_KD
st = nb.init_ga(n)
par = nb.default_param(n)
buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
dpar = torch.from_numpy(par).cuda()
ctx = nb.Context(device=local, order=order, unsort=0, tree_steps=tree_steps, m2l_first=1, rank=rank, world=world)
peer_setup(ctx, n)
ctx.compute_force(ev, buf.data_ptr(), n, dpar.data_ptr())
fmm_leapfrog_peer(ctx, buf, n, dpar.data_ptr(), 5e-4, steps)          # gathers the full state at the end
P, M = ctx.fmm_lists()                                                # this rank's part of the last evaluation's lists
every = [None] * world
dist.all_gather_object(every, (P, M))
ok = True
if rank == 0:
    from refs import Oracle
    orc = Oracle(order=order, unsort=0, tree_steps=tree_steps, m2l_first=1)
    ob = np.zeros(9 * n, np.float32)
    ob[:6 * n] = st.ravel()
    orc.eval(3, ob, n, par)
    orc.integrate(1, 3, ob, n, par, 5e-4, steps)
    o = ob.reshape(3, n, 3)
    g = buf.cpu().numpy().reshape(3, n, 3)
    dp = float(np.abs(g[0] - o[0]).max() / np.abs(o[0]).max())
    dv = float(np.abs(g[1] - o[1]).max() / np.abs(o[1]).max())
    OP, OM = orc.lists()
    up = np.unique(np.concatenate([e[0] for e in every]), axis=0)
    um = np.unique(np.concatenate([e[1] for e in every]), axis=0)
    lists_ok = bool(np.array_equal(up, OP) and np.array_equal(um, OM))
    ok = dp <= 1e-5 and dv <= 1e-4 and lists_ok
    print(f"PEER_PARITY {'OK' if ok else 'FAIL'} world {world} n {n} steps {steps} tree_steps {tree_steps}: pos {dp:.2e} vel {dv:.2e} "
          f"lists {'equal' if lists_ok else 'DIFFER'} ({len(up)} p2p, {len(um)} m2l; oracle {len(OP)}, {len(OM)})", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
ctx.peer_detach()
dist.barrier()
del ctx
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
