"""The reference's own GPU path (generic SIMT source compiled for sm_100 into oracle/_ref/libnbco_ref.so) timed on
device 0: coulombOscillatorFMMKD3 under the reference's leapfrog (tree_steps = 8), plus its direct3 kernel.
Prints one JSON line.  Run in its own process (bench.py does): the reference exit()s on CUDA errors.
   python tools/ref_gpu_baseline.py [N] [STEPS]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from refs import Ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
if not Ref.available() or not hasattr(Ref.lib(), "ref_fmm3_gpu_step_seconds"):
    print(json.dumps({"unavailable": "oracle/_ref/libnbco_ref.so without the GPU hooks"})); sys.exit(0)
L = Ref.lib()
ref = Ref(order=3, unsort=0, tree_steps=8)
st = Ref.init_ga(n)
import coulomb_oscillators_b200 as nb
par = nb.default_param(n)
buf = np.zeros(9 * n, np.float32)
buf[:6 * n] = st.reshape(-1)
ref.apply()
t0 = time.perf_counter()
sec = L.ref_fmm3_gpu_step_seconds(buf, n, par, 5e-4, 3, steps)
out = {"kind": "reference GPU path (fmm_cart3_kdtree + add_elastic under leapfrog, -arch=sm_100, unmodified)", "n": n, "order": 3,
       "steps": steps, "tree_steps": 8, "s_per_step": sec, "particle_steps_per_s": (n / sec) if sec > 0 else None,
       "finite": bool(np.isfinite(buf[:6 * n]).all()), "wall_s": round(time.perf_counter() - t0, 2)}
nd = min(n, 1 << 18)
pos = st[0][:nd].copy()
acc = np.zeros((nd, 3), np.float32)
pard = nb.default_param(nd)
sd = L.ref_direct3_gpu_seconds(pos.reshape(-1), acc.reshape(-1), nd, pard, 2)
out["direct3"] = {"n": nd, "s_per_eval": sd, "Ginteractions_per_s": (nd * nd / sd / 1e9) if sd > 0 else None}
print(json.dumps(out))
