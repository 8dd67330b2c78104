"""Per-phase CUDA-event times of the FMM evaluation for one build of the library (NBCO_LIB selects it):
   python tools/ab_phases.py N ORDER [EVALS]   ->  one JSON line {lib, n, order, L, phases: {name: avg ms}, ms_per_eval}
Run once per build (separate processes) to compare an older build with the current one on the same inputs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coulomb_oscillators_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
evals = int(sys.argv[3]) if len(sys.argv) > 3 else 16
st = nb.init_ga(n)
buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
par = torch.from_numpy(nb.default_param(n)).cuda()
ctx = nb.Context(order=p, unsort=0, tree_steps=8)
ev = nb.EVAL_COULOMB_FMM3_KD
ctx.compute_force(ev, buf.data_ptr(), n, par.data_ptr())
ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, par.data_ptr(), 5e-4, 3)
ctx.fmm_phase_totals(reset=True)
s = torch.cuda.ExternalStream(ctx.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(s)
ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, par.data_ptr(), 5e-4, evals)
e1.record(s); e1.synchronize()
tot, ne, nr = ctx.fmm_phase_totals(reset=True)
reb = ("kd_top", "kd_bottom", "permute")
ph = {k: round(v / max(nr if k in reb else ne, 1), 4) for k, v in tot.items()}
info = ctx.fmm_info()
print(json.dumps({"lib": os.path.basename(os.path.dirname(nb.lib_path)) + "/" + os.path.basename(nb.lib_path), "n": n, "order": p,
                  "L": int(info.levels), "evals": int(ne), "rebuilds": int(nr), "phases_ms": ph,
                  "ms_per_step": round(e0.elapsed_time(e1) / evals, 4)}))
