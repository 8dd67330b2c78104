"""Generate tests/golden/fmm2d_*.npz from the UNMODIFIED 2D reference (oracle/_ref/libnbco_ref2d.so:
fmm_cart_cpu, direct2_cpu and the integrators compiled with SCAL = double, DIM = 2).
    python tools/make_golden2d.py
The reference's CPU sort is unstable inside a grid cell, so everything is stored in INPUT order
(particles are matched back by their coordinates, which are distinct)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from refs2d import Ref2, by_position
import coulomb_oscillators_b200._lib as nb

OUT = os.path.join(ROOT, "tests", "golden")


def to_input_order(pos_in, pos_out, *arrays):
    ki, ko = by_position(pos_in), by_position(pos_out)
    assert np.array_equal(pos_in[ki], pos_out[ko])
    inv = np.empty_like(ki); inv[ki] = np.arange(len(ki))
    return [a[ko][inv] for a in arrays]


def fixture(name, n, order, dist, radius=1):
    st = nb.init_kv2(n) if dist == "kv" else nb.init_ga2(n)
    par = nb.default_param2(n)
    ref = Ref2(order=order, radius=radius, threads=4)
    d = {"pos": st[0].copy(), "vel": st[1].copy(), "param": par, "order": np.int32(order), "radius": np.int32(radius),
         "levels": np.int32(ref.levels(n))}
    for which, key in ((1, "acc_fmm"), (3, "acc_osc_fmm"), (0, "acc_direct")):
        buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
        ref.eval(which, buf, n, par)
        (d[key],) = to_input_order(st[0], buf[:n], buf[2 * n:])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "L", d["levels"])


def trajectory(name, n, order, scheme, steps):
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    ref = Ref2(order=order, threads=4)
    buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    ref.eval(3, buf, n, par)
    ref.integrate(scheme, 3, buf, n, par, 5e-4, steps)
    o = by_position(buf[:n])   # final state in canonical (position-sorted) order
    np.savez_compressed(os.path.join(OUT, name + ".npz"), state0=st, param=par, final=buf.reshape(3, n, 2)[:, o],
                        scheme=np.int32(scheme), order=np.int32(order), steps=np.int32(steps))
    print(name, "done")


if __name__ == "__main__":
    assert Ref2.available(), "build oracle/_ref first (make -f oracle/Makefile)"
    fixture("fmm2d_kv_n6000_p5", 6000, 5, "kv")
    fixture("fmm2d_ga_n5000_p3", 5000, 3, "ga")
    fixture("fmm2d_kv_n4000_p8_r2", 4000, 8, "kv", radius=2)
    trajectory("traj2d_fmm_pefrl_n3000", 3000, 4, 3, 4)
    trajectory("traj2d_fmm_fr_n3000", 3000, 4, 2, 4)
