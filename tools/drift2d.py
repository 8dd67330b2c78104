"""Energy drift of the 2D fp64 FMM under PEFRL at BASELINE config 3's size (the pair energy is an O(N^2) fp64 direct sum on the GPU):
   python tools/drift2d.py [n] [order] [steps] [dt]   ->   one JSON line {n, order, steps, dt, H0, H1, rel_drift, seconds ...}
The reference's CPU path needs ~2.4 s per evaluation at N = 2^22 (800 evaluations for 200 PEFRL steps), so its drift is quoted from
the same experiment at the size the CPU finishes (tests/test_fmm2_gpu.py::test_energy2_and_drift, tests/test_oracle2d.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
order = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dt = float(sys.argv[4]) if len(sys.argv) > 4 else 5e-4
st = nb.init_kv2(n)
par = torch.from_numpy(nb.default_param2(n)).cuda()
buf = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))])).cuda()
ctx = nb.Context(order=order)
t0 = time.time()
e0 = np.array(ctx.energy2(buf.data_ptr(), n, par.data_ptr()))
t_e = time.time() - t0
ctx.compute_force2(nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, par.data_ptr())
t0 = time.time()
ctx.integrate2(nb.PEFRL, nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, par.data_ptr(), dt, steps)
torch.cuda.synchronize()
t_i = time.time() - t0
e1 = np.array(ctx.energy2(buf.data_ptr(), n, par.data_ptr()))
print(json.dumps({"n": n, "order": order, "scheme": "PEFRL", "steps": steps, "dt": dt, "H0_terms": e0.tolist(), "H1_terms": e1.tolist(),
                  "H0": float(e0.sum()), "H1": float(e1.sum()), "rel_drift": float(abs(e1.sum() - e0.sum()) / abs(e0.sum())),
                  "seconds_energy": round(t_e, 2), "seconds_run": round(t_i, 2), "ms_per_step": round(1e3 * t_i / steps, 3)}))
