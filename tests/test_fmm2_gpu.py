"""GPU parity of the 2D fp64 path (nbco_force_fmm2 & co.) through the C ABI.

Gate 1 (integer / geometry, bit-exact): levels, permutation (stable cell sort), leaf index, multiplicities
and cell centres equal the oracle's.  Gate 2 (floating point, tolerance written here): accelerations within
1e-12 relative (rel_diff1, max over particles; north star fp64), multipole / local coefficients compared in
harmonic form (see include/nbco.h, nbco_fmm2_get_tree)."""
import os

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
import cx2d
from refs2d import Oracle2, Ref2, by_position, rel_err2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_eval(ctx, evaluator, pos, vel, par):
    """compute_force2 on a device state buffer; returns (pos, vel, acc) as numpy"""
    import torch
    n = pos.shape[0]
    buf = dev(np.concatenate([pos, vel, np.zeros((n, 2))]))
    dpar = None if par is None else dev(par)
    ctx.compute_force2(evaluator, buf.data_ptr(), n, None if par is None else dpar.data_ptr())
    torch.cuda.synchronize()
    b = buf.cpu().numpy().reshape(3, n, 2)
    return b[0], b[1], b[2]


def check_against_oracle(pos0, vel0, par, order, **cfg):
    ocfg = dict(order=order, radius=int(cfg.get("radius", 1)), eps2=cfg.get("eps2_d", 1e-18),
                dens=cfg.get("dens_inhom", 1.0), coll=cfg.get("coll", 1))
    ctx = nb.Context(order=order, **cfg)
    pos, vel, acc = gpu_eval(ctx, nb.EVAL_FMM2, pos0, vel0, par)
    o = Oracle2(**ocfg).fmm(pos0, vel0, par)
    T = ctx.fmm2_tree()
    L = o["levels"]
    tb = lambda l: (4 ** l - 1) // 3
    assert T["levels"] == L
    # gate 1: bit-exact
    assert np.array_equal(T["perm"], o["perm"])
    assert np.array_equal(pos, o["pos"]) and np.array_equal(vel, o["vel"])
    assert np.array_equal(T["mult"][tb(2):], o["mult"][tb(2):])
    assert np.array_equal(T["leaf_index"][:-1], o["index"][tb(L):tb(L + 1)]) and T["leaf_index"][-1] == pos0.shape[0]
    oc = o["center"][:, 0] + 1j * o["center"][:, 1]
    assert np.array_equal(T["center"][tb(L):], oc[tb(L):])              # sequential mean in cell order
    assert np.abs(T["center"][tb(2):] - oc[tb(2):]).max() <= 1e-15 * np.abs(oc).max()
    # gate 2
    Z, Lc = cx2d.reduce_sym(o["mpole"], order), cx2d.local_cx(o["local"], order)
    sel = slice(tb(2), None)
    assert np.all(np.abs(T["mpole"][sel] - Z[sel]).max(0) <= 1e-12 * np.abs(Z[sel]).max(0) + 1e-300)
    # the oracle (like the reference) evaluates order-2p polynomials in monomial form: coefficients agree to ~1e-11
    assert np.all(np.abs(T["local"][sel] - Lc[sel]).max(0) <= 1e-10 * np.abs(Lc[sel]).max(0) + 1e-300)
    m, mx = rel_err2(acc, o["acc"])
    assert mx < TOL, (m, mx)
    return ctx, acc, o


@pytest.mark.parametrize("n,order,dist", [
    (1, 3, "kv"), (7, 2, "kv"), (300, 1, "ga"), (4097, 3, "kv"), (6000, 5, "kv"), (20000, 4, "ga"), (50001, 6, "kv"),
    (30000, 8, "ga"), (12345, 10, "kv"), (200000, 5, "ga"), (1 << 18, 3, "kv"),
])
def test_fmm2_matches_oracle(n, order, dist):
    if n < 16:  # the reference's samplers normalise by the r.m.s. (0/0 for one particle)
        st = np.random.default_rng(n).normal(size=(2, n, 2)) * 1e-3
    else:
        st = nb.init_kv2(n) if dist == "kv" else nb.init_ga2(n)
    check_against_oracle(st[0], st[1], nb.default_param2(n), order)


@pytest.mark.parametrize("cfg", [dict(radius=2.0), dict(radius=3.7), dict(dens_inhom=4.0), dict(coll=0), dict(eps2_d=1e-7),
                                 dict(dens_inhom=0.25)])
def test_fmm2_options_match_oracle(cfg):
    n = 30000
    st = nb.init_ga2(n)
    check_against_oracle(st[0], st[1], nb.default_param2(n), 4, **cfg)


def test_fmm2_degenerate_inputs():
    """all particles in one point (delta clamps to eps, fmm_cart.cuh:472-474); duplicates; unscaled call"""
    n = 5000
    pos = np.zeros((n, 2)); pos[:] = (0.25, -0.5)
    check_against_oracle(pos, np.zeros((n, 2)), None, 3)
    st = nb.init_kv2(n)
    pos = st[0].copy(); pos[n // 2:] = pos[:n - n // 2]
    check_against_oracle(pos, st[1], None, 5)


@pytest.mark.parametrize("name", ["fmm2d_kv_n6000_p5", "fmm2d_ga_n5000_p3", "fmm2d_kv_n4000_p8_r2"])
def test_fmm2_against_reference_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n = g["pos"].shape[0]
    ctx = nb.Context(order=int(g["order"]), radius=float(g["radius"]))
    pos, vel, acc = gpu_eval(ctx, nb.EVAL_FMM2, g["pos"], g["vel"], g["param"])
    perm = ctx.fmm2_tree()["perm"]
    assert ctx.fmm2_info().levels == int(g["levels"])
    assert np.array_equal(pos, g["pos"][perm]) and np.array_equal(vel, g["vel"][perm])
    m, mx = rel_err2(acc, g["acc_fmm"][perm])
    assert mx < TOL, (m, mx)
    pos, vel, acc = gpu_eval(ctx, nb.EVAL_COULOMB_FMM2, g["pos"], g["vel"], g["param"])
    m, mx = rel_err2(acc, g["acc_osc_fmm"][perm])
    assert mx < TOL, (m, mx)
    pos, vel, acc = gpu_eval(ctx, nb.EVAL_DIRECT2, g["pos"], g["vel"], g["param"])
    m, mx = rel_err2(acc, g["acc_direct"])
    assert mx < TOL, (m, mx)
    # FMM error against the direct sum no worse than the reference's
    m_ref, _ = rel_err2(g["acc_fmm"], g["acc_direct"])
    pos, vel, accf = gpu_eval(ctx, nb.EVAL_FMM2, g["pos"], g["vel"], g["param"])
    m_us, _ = rel_err2(accf, acc[perm])
    assert m_us <= m_ref * (1 + 1e-9) + 1e-15


@pytest.mark.skipif(not Ref2.available(), reason="oracle/_ref/libnbco_ref2d.so not built")
def test_fmm2_against_live_reference():
    n, order = 100000, 5
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    b = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    Ref2(order=order).eval(3, b, n, par)
    pos, vel, acc = gpu_eval(nb.Context(order=order), nb.EVAL_COULOMB_FMM2, st[0], st[1], par)
    k1, k2 = by_position(b[:n]), by_position(pos)
    assert np.array_equal(b[:n][k1], pos[k2]) and np.array_equal(b[n:2 * n][k1], vel[k2])
    m, mx = rel_err2(acc[k2], b[2 * n:][k1])
    assert mx < TOL, (m, mx)


@pytest.mark.parametrize("n", [1, 255, 4096, 33333])
def test_direct2_matches_oracle(n):
    st = nb.init_ga2(n) if n > 1 else np.ones((2, 1, 2))
    par = nb.default_param2(n)
    ctx = nb.Context()
    _, _, acc = gpu_eval(ctx, nb.EVAL_DIRECT2, st[0], st[1], par)
    m, mx = rel_err2(acc, Oracle2().direct(st[0], par))
    assert mx < TOL, (m, mx)
    _, _, acc2 = gpu_eval(ctx, nb.EVAL_COULOMB_DIRECT2, st[0], st[1], par)
    assert np.abs(acc2 - (acc - st[0] * par[2:4])).max() <= 1e-15 * np.abs(acc2).max()


def test_direct2_sharded_targets_tile_the_result():
    import torch
    n = 20000
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    _, _, full = gpu_eval(nb.Context(), nb.EVAL_DIRECT2, st[0], st[1], par)
    out = np.full((n, 2), np.nan)
    for r in range(3):
        ctx = nb.Context(rank=r, world=3)
        p, a, dp = dev(st[0]), dev(np.full((n, 2), np.nan)), dev(par)
        ctx.force_direct2(p.data_ptr(), a.data_ptr(), n, dp.data_ptr())
        b, e = nb.shard_range(n, r, 3)
        got = a.cpu().numpy()
        assert np.isnan(got[:b]).all() and np.isnan(got[e:]).all()
        out[b:e] = got[b:e]
    assert np.array_equal(out, full)


@pytest.mark.parametrize("scheme", [nb.EULER, nb.LEAPFROG, nb.FORESTRUTH, nb.PEFRL])
def test_schemes2_match_oracle(scheme):
    n, order, steps = 5000, 4, 3
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    s = st.copy()
    acc = nb.Context(order=order).run_host2(scheme, nb.EVAL_COULOMB_FMM2, s, par, 5e-4, steps, want_acc=True)
    buf = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    orc = Oracle2(order=order)
    orc.eval(3, buf, n, par)
    orc.integrate(scheme, 3, buf, n, par, 5e-4, steps)
    # same stable sort on both sides -> same particle order
    assert np.abs(s[0] - buf[:n]).max() <= TOL * np.abs(buf[:n]).max()
    assert np.abs(s[1] - buf[n:2 * n]).max() <= TOL * np.abs(buf[n:2 * n]).max()
    m, mx = rel_err2(acc, buf[2 * n:])
    assert mx < 1e-11, (m, mx)


@pytest.mark.parametrize("name", ["traj2d_fmm_pefrl_n3000", "traj2d_fmm_fr_n3000"])
def test_trajectory2_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    s = g["state0"].copy()
    nb.Context(order=int(g["order"])).run_host2(int(g["scheme"]), nb.EVAL_COULOMB_FMM2, s, g["param"], 5e-4, int(g["steps"]))
    o = by_position(s[0])
    for k in range(2):
        assert np.abs(s[k][o] - g["final"][k]).max() <= TOL * np.abs(g["final"][k]).max(), k


def test_energy2_and_drift():
    import torch
    n = 4000
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    ctx = nb.Context(order=8, eps2_d=1e-7)
    buf = dev(np.concatenate([st[0], st[1], np.zeros((n, 2))]))
    dpar = dev(par)
    e_orc = Oracle2(eps2=1e-7).energy(np.concatenate([st[0], st[1], np.zeros((n, 2))]), n, par)
    e0 = np.array(ctx.energy2(buf.data_ptr(), n, dpar.data_ptr()))
    assert np.all(np.abs(e0 - e_orc) <= 1e-12 * np.abs(e_orc))
    ctx.compute_force2(nb.EVAL_COULOMB_DIRECT2, buf.data_ptr(), n, dpar.data_ptr())
    ctx.integrate2(nb.PEFRL, nb.EVAL_COULOMB_DIRECT2, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, 20)
    e1 = np.array(ctx.energy2(buf.data_ptr(), n, dpar.data_ptr()))
    assert abs(e1.sum() - e0.sum()) / abs(e0.sum()) < 1e-10
    # FMM-driven run: drift bounded by the expansion error, not by the integrator
    buf = dev(np.concatenate([st[0], st[1], np.zeros((n, 2))]))
    ctx.compute_force2(nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, dpar.data_ptr())
    ctx.integrate2(nb.PEFRL, nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, 20)
    e2 = np.array(ctx.energy2(buf.data_ptr(), n, dpar.data_ptr()))
    drift = abs(e2.sum() - e0.sum()) / abs(e0.sum())
    # ... and no worse than the reference algorithm's own drift on the same run (oracle = reference to 1e-13)
    ob = np.concatenate([st[0], st[1], np.zeros((n, 2))]).copy()
    orc = Oracle2(order=8, eps2=1e-7)
    orc.eval(3, ob, n, par)
    orc.integrate(nb.PEFRL, 3, ob, n, par, 5e-4, 20)
    drift_ref = abs(orc.energy(ob, n, par).sum() - e_orc.sum()) / abs(e_orc.sum())
    assert drift < 1e-4 and drift <= drift_ref * 1.001 + 1e-12, (drift, drift_ref)


def test_step2_elastic2_relerr2():
    import torch
    n = 100001
    rng = np.random.default_rng(0)
    b0, a0 = rng.normal(size=(n, 2)), rng.normal(size=(n, 2))
    b, a = dev(b0), dev(a0)
    ctx = nb.Context()
    ctx.step2(b.data_ptr(), a.data_ptr(), 0.37, n)
    want = b0 + a0 * 0.37
    assert np.all(np.abs(b.cpu().numpy() - want) <= np.spacing(np.abs(want)) + np.spacing(np.abs(a0 * 0.37)))   # one fma vs mul + add
    k = dev(np.array([1.5, 0.25]))
    acc = dev(a0)
    ctx.add_elastic2(b.data_ptr(), acc.data_ptr(), n, k.data_ptr())
    want = a0 - b.cpu().numpy() * [1.5, 0.25]
    assert np.all(np.abs(acc.cpu().numpy() - want) <= np.spacing(np.abs(want)) + np.spacing(np.abs(b.cpu().numpy() * 1.5)))
    m, mx = ctx.mean_rel_err2(acc.data_ptr(), a.data_ptr(), n)
    mm, mmx = rel_err2(acc.cpu().numpy(), a0)
    assert abs(m - mm) <= 1e-12 * mm and abs(mx - mmx) <= 1e-12 * mmx


def test_fmm2_full_size_properties():
    """BASELINE config 4 size (N = 4M, p = 5): size-independent properties -- the output is a permutation of
    the input, cells are sorted, multiplicities sum to N on every level, and the FMM agrees with the direct
    sum on a shard of targets to the expansion error of the order."""
    import torch
    n, order = 1 << 22, 5
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    ctx = nb.Context(order=order)
    pos, vel, acc = gpu_eval(ctx, nb.EVAL_FMM2, st[0], st[1], par)
    T = ctx.fmm2_tree()
    L = T["levels"]
    assert L == nb.fmm2_levels(n, order) == 9
    perm = T["perm"]
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.int32))
    assert np.array_equal(pos, st[0][perm]) and np.array_equal(vel, st[1][perm])
    tb = lambda l: (4 ** l - 1) // 3
    for l in range(2, L + 1):
        assert T["mult"][tb(l):tb(l + 1)].sum() == n
    assert np.all(np.diff(T["leaf_index"]) >= 0)
    # direct sum on 1/64 of the targets (sorted order)
    c2 = nb.Context(rank=5, world=64)
    p, a, dp = dev(pos), dev(np.zeros((n, 2))), dev(par)
    c2.force_direct2(p.data_ptr(), a.data_ptr(), n, dp.data_ptr())
    b, e = nb.shard_range(n, 5, 64)
    m, mx = rel_err2(acc[b:e], a.cpu().numpy()[b:e])
    assert m < 2e-5, (m, mx)


@pytest.mark.parametrize("world", [2, 4])
def test_fmm2_rank_shards_tile_the_full_result(world):
    """2D multi-GPU path on one device: with cfg.rank / cfg.world the tree is replicated and the near-field + L2P kernel
    writes only the rank's range of the cell-sorted particles; the ranges assemble to the world = 1 result bit for bit
    (every acceleration is one independent sum)"""
    import torch
    n, p = 50001, 5
    st = nb.init_kv2(n)
    par = torch.from_numpy(nb.default_param2(n)).cuda()

    def run(rank, w):
        c = nb.Context(order=p, rank=rank, world=w)
        b = torch.from_numpy(np.concatenate([st[0], st[1], np.full((n, 2), np.nan)])).cuda()
        c.compute_force2(nb.EVAL_COULOMB_FMM2, b.data_ptr(), n, par.data_ptr())
        return b.cpu().numpy().reshape(3, n, 2)

    full = run(0, 1)
    out = np.full((n, 2), np.nan)
    for r in range(world):
        got = run(r, world)
        assert np.array_equal(got[0], full[0]) and np.array_equal(got[1], full[1])      # same sort on every rank
        b, e = nb.shard_range(n, r, world)
        assert np.isnan(got[2][:b]).all() and np.isnan(got[2][e:]).all()
        out[b:e] = got[2][b:e]
    assert np.array_equal(out, full[2])
