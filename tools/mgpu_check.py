"""Run under torchrun: sharded FMM leapfrog on WORLD_SIZE GPUs vs the single-GPU integrator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import coulomb_oscillators_b200 as nb
from coulomb_oscillators_b200.parallel import fmm_leapfrog_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, steps = 300001, 10
st = nb.init_ga(n)
par = torch.from_numpy(nb.default_param(n)).cuda()
ctx = nb.Context(device=local, order=3, unsort=0, tree_steps=4, rank=rank, world=world)
buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
ctx.compute_force(nb.EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, par.data_ptr())
fmm_leapfrog_sharded(ctx, buf, n, par.data_ptr(), 5e-4, steps)
got = buf.cpu().numpy().reshape(3, n, 3)
if rank == 0:
    c1 = nb.Context(device=local, order=3, unsort=0, tree_steps=4)
    b1 = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
    b1[:6 * n] = torch.from_numpy(st.ravel()).cuda()
    c1.compute_force(nb.EVAL_COULOMB_FMM3_KD, b1.data_ptr(), n, par.data_ptr())
    c1.integrate(nb.LEAPFROG, nb.EVAL_COULOMB_FMM3_KD, b1.data_ptr(), n, par.data_ptr(), 5e-4, steps)
    want = b1.cpu().numpy().reshape(3, n, 3)
    for k, name in enumerate(("pos", "vel", "acc")):
        print(name, "max rel diff", float(np.abs(got[k] - want[k]).max() / np.abs(want[k]).max()))
    ok = np.abs(got[0] - want[0]).max() <= 1e-6 * np.abs(want[0]).max() and np.abs(got[1] - want[1]).max() <= 1e-5 * np.abs(want[1]).max()
    print("MGPU_CHECK", "OK" if ok else "FAIL", "world", world)
dist.barrier()
dist.destroy_process_group()
