"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libnbco_ref.so, CPU path).
Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tools/make_golden.py
Each fixture holds the inputs and everything the reference computed for them: sorted positions,
permutation, tree arrays, sorted interaction lists (both traversal orders), multipoles, locals and
accelerations in tree order, plus direct3_cpu accelerations in input order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from refs import Ref
import coulomb_oscillators_b200._lib as nb

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

def fixture(name, n, order, dist):
    st = nb.init_ga(n) if dist == "ga" else nb.init_test_cube(n)
    par = nb.default_param(n)
    ref = Ref(order=order, threads=4)
    d = {"pos": st[0].copy(), "vel": st[1].copy(), "param": par, "order": np.int32(order)}
    for mf in (0, 1):
        R = ref.fmm3_phases(st[0], par, mf)
        if mf == 0:
            for k in ("perm", "lbound", "rbound", "center", "mult", "index", "splitdim", "mpole", "pos_sorted"):
                d[k] = R[k]
            d["levels"] = np.int32(R["levels"])
        d[f"p2p_{mf}"] = R["p2p"]; d[f"m2l_{mf}"] = R["m2l"]
        d[f"local_{mf}"] = R["local"]; d[f"acc_{mf}"] = R["acc_sorted"]
    buf = np.zeros(9 * n, np.float32); buf[:3 * n] = st[0].ravel()
    ref.eval(0, buf, n, par)
    d["acc_direct"] = buf[6 * n:].reshape(n, 3).copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "L", R["levels"], {k: v.shape for k, v in d.items() if hasattr(v, "shape") and v.ndim})

def fixture_shallow(name, n, order, dens):
    """A tree of two levels (the reference's own level rule with a tiny density parameter, fmm_cart3_kdtree.cuh:1508-1512)
    whose level-1 nodes hold more particles than a bottom CTA of our kd build: pins the virtual-level build and the oracle to
    the reference.  Tie-free coordinates (the reference's segmented sorts are unstable), zero velocities (they compress)."""
    from refs import unique_axes
    st = nb.init_ga(n)
    pos = unique_axes(st[0].copy())
    par = nb.default_param(n)
    ref = Ref(order=order, threads=4, dens_inhom=dens)
    d = {"pos": pos.copy(), "vel": np.zeros_like(pos), "param": par, "order": np.int32(order), "dens_inhom": np.float32(dens)}
    for mf in (0, 1):
        R = ref.fmm3_phases(pos, par, mf)
        if mf == 0:
            for k in ("perm", "lbound", "rbound", "center", "mult", "index", "splitdim", "mpole", "pos_sorted"):
                d[k] = R[k]
            d["levels"] = np.int32(R["levels"])
        d[f"p2p_{mf}"] = R["p2p"]; d[f"m2l_{mf}"] = R["m2l"]
        d[f"local_{mf}"] = R["local"]; d[f"acc_{mf}"] = R["acc_sorted"]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "L", R["levels"], "level-1 node:", int(R["mult"][1]), {k: v.shape for k, v in d.items() if hasattr(v, "shape") and v.ndim})


def trajectory(name, n, scheme, which, steps):
    st = nb.init_ga(n)
    par = nb.default_param(n)
    ref = Ref(order=3, threads=4, unsort=0)
    buf = np.zeros(9 * n, np.float32); buf[:6 * n] = st.reshape(-1)
    ref.eval(which, buf, n, par)
    ref.integrate(scheme, which, buf, n, par, 5e-4, steps)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), state0=st, param=par, final=buf.reshape(3, n, 3),
                        scheme=np.int32(scheme), which=np.int32(which), steps=np.int32(steps))
    print(name, "done")

if __name__ == "__main__":
    assert Ref.available(), "build oracle/_ref first (make -f oracle/Makefile)"
    fixture("fmm_ga_n3000_p3", 3000, 3, "ga")
    fixture("fmm_cube_n4096_p4", 4096, 4, "cube")
    fixture("fmm_ga_n2500_p1", 2500, 1, "ga")
    fixture_shallow("fmm_ga_n17000_p2_shallow", 17000, 2, 1e-4)
    trajectory("traj_direct_leapfrog_n512", 512, 1, 2, 20)   # coulombOscillatorDirect_cpu, config 1 in miniature
    trajectory("traj_fmm_pefrl_n2048", 2048, 3, 3, 4)        # coulombOscillatorFMMKD3_cpu under PEFRL
