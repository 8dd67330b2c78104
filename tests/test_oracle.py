"""Pins the oracle (oracle/nbco_oracle.c): against the golden fixtures generated from the
unmodified reference (tools/make_golden.py), against the reference's own -test table recorded in
SURVEY.md section 6, and -- where oracle/_ref was built -- against the reference run live."""
import os

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from refs import Oracle, Ref, mean_rel_err, unique_axes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# the last one: a two-level tree whose level-1 nodes hold 8500 particles (shallow: built with virtual levels on the GPU)
FIXTURES = ["fmm_ga_n3000_p3", "fmm_cube_n4096_p4", "fmm_ga_n2500_p1", "fmm_ga_n17000_p2_shallow"]


def fixture_cfg(g):
    return {"dens_inhom": float(g["dens_inhom"])} if "dens_inhom" in g.files else {}


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("m2l_first", [0, 1])
def test_oracle_against_reference_fixture(name, m2l_first):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    orc = Oracle(order=int(g["order"]), unsort=0, m2l_first=m2l_first, **fixture_cfg(g))
    pos = g["pos"].copy()
    acc = orc.fmm3_kd(pos, None, g["param"])
    T = orc.tree()
    P, M = orc.lists()
    assert T["levels"] == int(g["levels"])
    # integer / geometry: bit-exact
    for k in ("perm", "lbound", "rbound", "center", "mult", "index", "splitdim"):
        assert np.array_equal(T[k], g[k]), k
    assert np.array_equal(pos, g["pos_sorted"])
    assert np.array_equal(P, g[f"p2p_{m2l_first}"]) and np.array_equal(M, g[f"m2l_{m2l_first}"])
    # floating point: tolerance 1e-5 relative (north star), observed ~1e-7
    assert np.abs(T["mpole"] - g["mpole"]).max() <= 1e-6 * np.abs(g["mpole"]).max()
    assert np.abs(T["local"] - g[f"local_{m2l_first}"]).max() <= 1e-5 * np.abs(g[f"local_{m2l_first}"]).max()
    m, mx = mean_rel_err(acc, g[f"acc_{m2l_first}"])
    assert m < 1e-6 and mx < 1e-5


def test_oracle_direct_against_reference_fixture():
    g = np.load(os.path.join(GOLD, "fmm_ga_n3000_p3.npz"))
    a = Oracle().direct3(g["pos"].copy(), g["param"])
    m, mx = mean_rel_err(a, g["acc_direct"])
    assert m < 1e-7 and mx < 1e-5


def test_unsort_mode_returns_input_order():
    g = np.load(os.path.join(GOLD, "fmm_ga_n3000_p3.npz"))
    pos = g["pos"].copy()
    a = Oracle(order=3, unsort=1).fmm3_kd(pos, None, g["param"])
    assert np.array_equal(pos, g["pos"])
    m, mx = mean_rel_err(a[g["perm"]], g["acc_0"])
    assert mx < 1e-5


def test_reference_test_table():
    """`nbco3 -cpu -test -n 8192` prints the mean relative error of the FMM against direct3 on the
    uniform cube for p = 1..10 (main3.cu:790-811); SURVEY.md section 6 recorded the reference's output."""
    table = [0.2315, 0.1070, 0.0595, 0.0299, 0.0153, 0.00934]
    n = 8192
    st = nb.init_test_cube(n)
    par = nb.default_param(n)
    d = Oracle().direct3(st[0].copy(), par)
    for p, want in enumerate(table, start=1):
        a = Oracle(order=p, unsort=1).fmm3_kd(st[0].copy(), None, par)
        got = mean_rel_err(a, d)[0]
        assert abs(got - want) <= 0.01 * want + 2e-5, (p, got, want)


@pytest.mark.parametrize("name", ["traj_direct_leapfrog_n512", "traj_fmm_pefrl_n2048"])
def test_trajectories_against_reference_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n = g["state0"].shape[1]
    buf = np.zeros(9 * n, np.float32)
    buf[:6 * n] = g["state0"].ravel()
    orc = Oracle(order=3, unsort=0, tree_steps=1)
    orc.eval(int(g["which"]), buf, n, g["param"])
    orc.integrate(int(g["scheme"]), int(g["which"]), buf, n, g["param"], 5e-4, int(g["steps"]))
    got, want = buf.reshape(3, n, 3), g["final"]
    # same tree order on both sides (rebuild every call, identical partition)
    assert np.abs(got[0] - want[0]).max() <= 1e-6 * np.abs(want[0]).max()
    assert np.abs(got[1] - want[1]).max() <= 1e-5 * np.abs(want[1]).max()


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built in this checkout")
@pytest.mark.parametrize("n,p,dist", [(8192, 3, "ga"), (20011, 2, "ga"), (30000, 5, "cube"), (16384, 6, "ga")])
def test_oracle_against_live_reference(n, p, dist):
    st = nb.init_ga(n) if dist == "ga" else nb.init_test_cube(n)
    st[0] = unique_axes(st[0])   # the reference's tie order is undefined (unstable sorts, thread count)
    par = nb.default_param(n)
    for m2l_first in (0, 1):
        R = Ref(order=p, threads=4).fmm3_phases(st[0], par, m2l_first)
        orc = Oracle(order=p, unsort=0, m2l_first=m2l_first)
        pos = st[0].copy()
        acc = orc.fmm3_kd(pos, None, par)
        T = orc.tree()
        P, M = orc.lists()
        for k in ("perm", "lbound", "rbound", "center", "mult", "index", "splitdim"):
            assert np.array_equal(T[k], R[k]), k
        assert np.array_equal(P, R["p2p"]) and np.array_equal(M, R["m2l"])
        m, mx = mean_rel_err(acc, R["acc_sorted"])
        assert m < 1e-6 and mx < 1e-5


def test_energy_of_two_particles():
    # H = 1/2 v^2 + 1/2 k x^2 + (xi/N) / |d| for two particles, checked by hand
    n = 2
    buf = np.zeros(9 * n, np.float32)
    buf[0:6] = [0.5, 0, 0, -0.5, 0, 0]
    buf[6:12] = [0, 1, 0, 0, -1, 0]
    par = np.array([0.25, 0, 0, 2.0, 1.0, 1.0], np.float32)
    e = Oracle().energy(buf, n, par)
    assert np.allclose(e, [1.0, 0.5, 0.25], rtol=1e-6)


def _chain_push(axis, chain):
    """axes in order of most recent use, without repeats (fmm3_common.cuh: chain_push)"""
    return [axis] + [a for a in chain if a != axis][:2]


@pytest.mark.parametrize("n,max_level,quantise", [(3000, 2, False), (5000, 4, True), (4097, 3, True)])
def test_order_inside_a_leaf_pair_is_the_total_order_of_their_parent(n, max_level, quantise):
    """What the CUDA kd build relies on (kdtree.cu header): only the LAST level's order is observable, and it is the total order
    (coordinate on the parent's split axis, coordinates on the previously used distinct axes, most recent first, input index) of
    each level-(L-1) node.  That is also why a shallow tree can be finished with virtual levels that inherit the parent's axis
    and chain.  Checked on the oracle's output, with many equal keys in the quantised cases."""
    rng = np.random.default_rng(n)
    pos = rng.normal(size=(n, 3)).astype(np.float32)
    if quantise:
        pos = (np.round(pos * 8) / 8).astype(np.float32)          # ~50 distinct values per axis
    orc = Oracle(order=2, unsort=0, m2l_first=1, max_level=max_level)
    sp = pos.copy()
    orc.fmm3_kd(sp, None, nb.default_param(n))
    T = orc.tree()
    L = int(T["levels"])
    assert L == max_level
    perm, split, index, mult = T["perm"], T["splitdim"], T["index"], T["mult"]
    assert np.array_equal(sp, pos[perm])
    bits = pos.view(np.uint32)                                    # the sort key: fp32 bits made monotone (-0.0 sorts before +0.0)
    ob = np.where(bits >> 31 == 0, bits ^ np.uint32(0x80000000), ~bits)
    chains = {0: _chain_push(int(split[0]), [])}
    for node in range(1, (1 << L) - 1):                           # nodes above the leaves
        chains[node] = _chain_push(int(split[node]), chains[(node - 1) >> 1])
    for node in range((1 << (L - 1)) - 1, (1 << L) - 1):          # level L-1
        lo, cnt = int(index[node]), int(mult[node])
        ids = perm[lo:lo + cnt]
        keys = [tuple(int(ob[i, a]) for a in chains[node]) + (int(i),) for i in ids]
        assert keys == sorted(keys), node
