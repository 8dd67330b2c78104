"""Run under torchrun: peer-memory FMM leapfrog (csrc/peer.cu) on WORLD_SIZE GPUs vs the single-GPU integrator,
then a timed run.   torchrun --nproc-per-node 2 tools/peer_check.py [n_check] [n_time] [steps_time]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import coulomb_oscillators_b200 as nb
from coulomb_oscillators_b200.parallel import peer_setup, fmm_leapfrog_peer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_check = int(sys.argv[1]) if len(sys.argv) > 1 else 300001
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
steps_time = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ev = nb.EVAL_COULOMB_FMM3_KD


def state(n):
    st = nb.init_ga(n)
    buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
    buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
    return buf, torch.from_numpy(nb.default_param(n)).cuda()


# ---- parity against one GPU (10 steps, rebuild every 4: two rebuilds fetch the remote ranges) ----
n, steps = n_check, 10
buf, par = state(n)
ctx = nb.Context(device=local, order=3, unsort=0, tree_steps=4, rank=rank, world=world)
peer_setup(ctx, n)
ctx.compute_force(ev, buf.data_ptr(), n, par.data_ptr())
fmm_leapfrog_peer(ctx, buf, n, par.data_ptr(), 5e-4, steps)
got = buf.cpu().numpy().reshape(3, n, 3)
ok = True
if rank == 0:
    c1 = nb.Context(device=local, order=3, unsort=0, tree_steps=4)
    b1, _ = state(n)
    c1.compute_force(ev, b1.data_ptr(), n, par.data_ptr())
    c1.integrate(nb.LEAPFROG, ev, b1.data_ptr(), n, par.data_ptr(), 5e-4, steps)
    want = b1.cpu().numpy().reshape(3, n, 3)
    for k, name in enumerate(("pos", "vel", "acc")):
        print(name, "max rel diff", float(np.abs(got[k] - want[k]).max() / np.abs(want[k]).max()), flush=True)
    ok = np.abs(got[0] - want[0]).max() <= 1e-6 * np.abs(want[0]).max() and np.abs(got[1] - want[1]).max() <= 1e-5 * np.abs(want[1]).max()
    print("PEER_CHECK", "OK" if ok else "FAIL", "world", world, flush=True)
    del c1, b1
ctx.peer_detach()
del ctx, buf
dist.barrier()

# ---- timing ----
if n_time > 0:
    n = n_time
    buf, par = state(n)
    ctx = nb.Context(device=local, order=3, unsort=0, tree_steps=8, rank=rank, world=world)
    peer_setup(ctx, n)
    ctx.compute_force(ev, buf.data_ptr(), n, par.data_ptr())
    fmm_leapfrog_peer(ctx, buf, n, par.data_ptr(), 5e-4, 8, gather_final=False)
    ctx.fmm_phase_totals(reset=True)
    stream = torch.cuda.ExternalStream(ctx.stream)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fmm_leapfrog_peer(ctx, buf, n, par.data_ptr(), 5e-4, steps_time, gather_final=False)
    e1.record(stream); e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot, evals, rebuilds = ctx.fmm_phase_totals(reset=True)
    if rank == 0:
        print(f"PEER_TIME world {world} n {n}: {ms.item() / steps_time:.3f} ms/step -> {n * steps_time / ms.item() / 1e6:.3f} G particle-steps/s", flush=True)
        print("  per-eval phase ms (rank 0):", {k: round(v / max(rebuilds if k in ('kd_top', 'kd_bottom', 'permute') else evals, 1), 3) for k, v in tot.items()},
              "evals", evals, "rebuilds", rebuilds, flush=True)
    ctx.peer_detach()
dist.barrier()
dist.destroy_process_group()
