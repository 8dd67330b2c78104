"""Pins the 2D oracle (oracle/nbco_oracle2d.c) against the golden fixtures generated from the unmodified
reference (tools/make_golden2d.py: fmm_cart_cpu / direct2_cpu / integrators with SCAL = double, DIM = 2)
and, where oracle/_ref was built, against the reference run live.  Tolerance: 1e-12 relative (north star,
fp64); observed <= 4e-14."""
import os

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from refs2d import Oracle2, Ref2, by_position, rel_err2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["fmm2d_kv_n6000_p5", "fmm2d_ga_n5000_p3", "fmm2d_kv_n4000_p8_r2"]
TOL = 1e-12


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle2_against_reference_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n = g["pos"].shape[0]
    orc = Oracle2(order=int(g["order"]), radius=int(g["radius"]))
    assert orc.levels(n) == int(g["levels"])
    r = orc.fmm(g["pos"], g["vel"], g["param"])
    perm = r["perm"]
    assert np.array_equal(r["pos"], g["pos"][perm]) and np.array_equal(r["vel"], g["vel"][perm])
    m, mx = rel_err2(r["acc"], g["acc_fmm"][perm])
    assert mx < TOL, (m, mx)
    m, mx = rel_err2(orc.direct(g["pos"], g["param"]), g["acc_direct"])
    assert mx < TOL, (m, mx)
    buf = np.concatenate([g["pos"], g["vel"], np.zeros((n, 2))]).copy()
    orc.eval(3, buf, n, g["param"])
    m, mx = rel_err2(buf[2 * n:], g["acc_osc_fmm"][perm])
    assert mx < TOL, (m, mx)
    # the FMM approximates the direct sum (same order of magnitude as the reference's own error)
    m_ref, _ = rel_err2(g["acc_fmm"], g["acc_direct"])
    m_orc, _ = rel_err2(r["acc"], g["acc_direct"][perm])
    assert abs(m_orc - m_ref) <= 1e-9 * max(m_ref, 1e-30) + 1e-15


def test_oracle2_tree_invariants():
    g = np.load(os.path.join(GOLD, "fmm2d_kv_n6000_p5.npz"))
    n = g["pos"].shape[0]
    r = Oracle2(order=5).fmm(g["pos"], None, g["param"])
    L = r["levels"]
    tb = lambda l: (4 ** l - 1) // 3
    for l in range(2, L + 1):
        assert r["mult"][tb(l):tb(l + 1)].sum() == n
    idx = r["index"][tb(L):tb(L + 1)]
    assert np.all(np.diff(idx) >= 0) and idx[0] == 0
    assert sorted(r["perm"].tolist()) == list(range(n))
    # monopole = multiplicity, dipole about the centre of charge identically 0
    assert np.array_equal(r["mpole"][tb(2):, 0], r["mult"][tb(2):].astype(float))
    assert np.all(r["mpole"][:, 1:3] == 0)


@pytest.mark.parametrize("name", ["traj2d_fmm_pefrl_n3000", "traj2d_fmm_fr_n3000"])
def test_oracle2_trajectory_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    st = g["state0"]
    n = st.shape[1]
    orc = Oracle2(order=int(g["order"]))
    buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    orc.eval(3, buf, n, g["param"])
    orc.integrate(int(g["scheme"]), 3, buf, n, g["param"], 5e-4, int(g["steps"]))
    o = by_position(buf[:n])
    fin = buf.reshape(3, n, 2)[:, o]
    for k in range(3):
        assert np.abs(fin[k] - g["final"][k]).max() <= TOL * np.abs(g["final"][k]).max(), k


@pytest.mark.skipif(not Ref2.available(), reason="oracle/_ref/libnbco_ref2d.so not built")
@pytest.mark.parametrize("n,order,dist", [(3000, 2, "kv"), (9000, 6, "ga"), (2048, 10, "kv")])
def test_oracle2_against_live_reference(n, order, dist):
    st = nb.init_kv2(n) if dist == "kv" else nb.init_ga2(n)
    par = nb.default_param2(n)
    b1 = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    b2 = b1.copy()
    Ref2(order=order, threads=4).eval(3, b1, n, par)
    Oracle2(order=order).eval(3, b2, n, par)
    k1, k2 = by_position(b1[:n]), by_position(b2[:n])
    assert np.array_equal(b1[:n][k1], b2[:n][k2]) and np.array_equal(b1[n:2 * n][k1], b2[n:2 * n][k2])
    m, mx = rel_err2(b2[2 * n:][k2], b1[2 * n:][k1])
    assert mx < TOL, (m, mx)


def test_oracle2_energy_drift_orders():
    """with a softening of the order of the particle spacing the dynamics is smooth: leapfrog drifts
    as dt^2, PEFRL as dt^4 (with the default eps = 1e-9 hard collisions dominate the drift)"""
    n = 600
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    orc = Oracle2(order=6, eps2=1e-7)

    def drift(scheme, dt, steps):
        buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
        orc.eval(2, buf, n, par)
        e0 = orc.energy(buf, n, par).sum()
        orc.integrate(scheme, 2, buf, n, par, dt, steps)
        return abs(orc.energy(buf, n, par).sum() - e0) / abs(e0)

    l1, l2 = drift(nb.LEAPFROG, 5e-4, 40), drift(nb.LEAPFROG, 2.5e-4, 80)
    p1, p2 = drift(nb.PEFRL, 5e-4, 40), drift(nb.PEFRL, 2.5e-4, 80)
    assert 3.5 < l1 / l2 < 4.5
    assert 10 < p1 / p2 < 25 and p1 < 1e-10
