"""Scratch: time the direct-sum kernel variants on the GPU (CUDA events on the context stream)."""
import os, sys, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import coulomb_oscillators_b200 as nb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
variant = os.environ.get("NBCO_DIRECT_VARIANT", "0")
state = nb.init_ga(n)
pos = torch.from_numpy(state[0].copy()).cuda()
acc = torch.empty_like(pos)
par = torch.from_numpy(nb.default_param(n)).cuda()
ctx = nb.Context()
st = torch.cuda.ExternalStream(ctx.stream)
ctx.force_direct3(pos.data_ptr(), acc.data_ptr(), n, par.data_ptr())
best = 1e9
for r in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); ctx.force_direct3(pos.data_ptr(), acc.data_ptr(), n, par.data_ptr()); e1.record(st); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
inter = n * n / (best * 1e-3)
print(f"variant {variant} n={n} {best:.3f} ms  {inter/1e12:.3f} Tinteractions/s  frac_fma_peak(18flop)={inter*18/74.5e12:.3f}")
if n <= 65536:
    L = ctypes.CDLL("oracle/_ref/libnbco_ref.so")
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
    L.ref_config.argtypes=[ctypes.c_int,ctypes.c_float,ctypes.c_float,ctypes.c_float]+[ctypes.c_int]*4
    L.ref_config(3, 1.0, 1e-18, 1.0, os.cpu_count(), 1, 1, 8)
    L.ref_eval.argtypes=[ctypes.c_int,f32p,ctypes.c_int,f32p]
    buf = np.zeros(9*n, np.float32); buf[:6*n] = state.ravel()
    t = time.time(); L.ref_eval(0, buf, n, nb.default_param(n)); t = time.time() - t
    ref = buf[6*n:].reshape(n, 3)
    a = acc.cpu().numpy()
    rel = np.linalg.norm(a - ref, axis=1) / np.sqrt((ref**2).sum(1) + 1e-18)
    print(f"  vs reference direct3_cpu ({t:.2f}s, {n*n/t/1e9:.2f} Ginter/s on {os.cpu_count()} threads): mean rel {rel.mean():.3e} max {rel.max():.3e}")
