"""Energy drift of the 3D kd-tree FMM under leapfrog (the pair energy is an O(N^2) direct sum on the GPU):
   python tools/drift3d.py [n] [order] [steps] [dt] [tree_steps]   ->   one JSON line
The reference's own drift on the same initial state comes from its CPU path (coulombOscillatorFMMKD3_cpu under leapfrog,
integrator.cuh:68-96) run in the build container; both are recorded in profiles/r02_drift3d.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dt = float(sys.argv[4]) if len(sys.argv) > 4 else 5e-4
tree_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 8
st = nb.init_ga(n)
par = torch.from_numpy(nb.default_param(n)).cuda()
buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
ctx = nb.Context(order=order, unsort=0, tree_steps=tree_steps)
ev = nb.EVAL_COULOMB_FMM3_KD
e0 = np.array(ctx.energy(buf.data_ptr(), n, par.data_ptr()), dtype=np.float64)
ctx.compute_force(ev, buf.data_ptr(), n, par.data_ptr())
ctx.fmm_phase_totals(reset=True)
t0 = time.time()
ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, par.data_ptr(), dt, steps)
torch.cuda.synchronize()
t_i = time.time() - t0
tot, ne, nr = ctx.fmm_phase_totals(reset=True)
reb = ("kd_top", "kd_bottom", "permute")
phases = {k: round(v / max(nr if k in reb else ne, 1), 4) for k, v in tot.items()}
info = ctx.fmm_info()
e1 = np.array(ctx.energy(buf.data_ptr(), n, par.data_ptr()), dtype=np.float64)
print(json.dumps({"n": n, "order": order, "scheme": "leapfrog", "steps": steps, "dt": dt, "tree_steps": tree_steps, "H0_terms": e0.tolist(),
                  "H1_terms": e1.tolist(), "H0": float(e0.sum()), "H1": float(e1.sum()),
                  "rel_drift": float(abs(e1.sum() - e0.sum()) / abs(e0.sum())), "ms_per_step": round(1e3 * t_i / steps, 3), "evals": int(ne), "rebuilds": int(nr), "phases_ms": phases,
                  "p2p_pairs": int(info.p2p_pairs), "m2l_pairs": int(info.m2l_pairs), "levels": int(info.levels)}))
