"""GPU parity of the time-stepping pieces: step, add_elastic, the four schemes, energy, rel-err."""
import os

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from refs import Oracle, mean_rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_step_and_elastic_are_single_fma():
    import torch
    n = 100001
    rng = np.random.default_rng(0)
    b = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).cuda()
    a = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).cuda()
    ctx = nb.Context()
    want = torch.from_numpy(np.float32(b.cpu().double().numpy() + a.cpu().double().numpy() * np.float64(np.float32(0.37)))).cuda()
    ctx.step(b.data_ptr(), a.data_ptr(), 0.37, n)                 # b = fma(ds, a, b): one rounding
    assert torch.equal(b, want)
    k = torch.tensor([1.199025, 1.0, 1.0], dtype=torch.float32).cuda()
    x = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).cuda()
    acc = a.clone()
    ctx.add_elastic(x.data_ptr(), acc.data_ptr(), n, k.data_ptr())
    want = torch.from_numpy(np.float32(a.cpu().double().numpy() - x.cpu().double().numpy() * k.cpu().double().numpy())).cuda()
    assert torch.equal(acc, want)


@pytest.mark.parametrize("scheme", [nb.EULER, nb.LEAPFROG, nb.FORESTRUTH, nb.PEFRL])
def test_schemes_match_oracle_direct(scheme):
    n = 2048
    st = nb.init_ga(n)
    par = nb.default_param(n)
    s = st.copy()
    nb.Context().run_host(scheme, nb.EVAL_COULOMB_DIRECT3, s, par, 5e-4, 5)
    buf = np.zeros(9 * n, np.float32)
    buf[:6 * n] = st.ravel()
    orc = Oracle()
    orc.eval(2, buf, n, par)
    orc.integrate(scheme, 2, buf, n, par, 5e-4, 5)
    o = buf.reshape(3, n, 3)
    assert np.abs(s[0] - o[0]).max() <= 2e-6 * np.abs(o[0]).max()
    assert np.abs(s[1] - o[1]).max() <= 2e-5 * np.abs(o[1]).max()


def test_config1_trajectory_against_reference_fixture():
    """BASELINE config 1 in miniature: direct sum + leapfrog, against the reference's own run"""
    g = np.load(os.path.join(GOLD, "traj_direct_leapfrog_n512.npz"))
    s = g["state0"].copy()
    nb.Context().run_host(int(g["scheme"]), nb.EVAL_COULOMB_DIRECT3, s, g["param"], 5e-4, int(g["steps"]))
    assert np.abs(s[0] - g["final"][0]).max() <= 2e-6 * np.abs(g["final"][0]).max()
    assert np.abs(s[1] - g["final"][1]).max() <= 2e-5 * np.abs(g["final"][1]).max()


def test_pefrl_fmm_trajectory_against_reference_fixture():
    g = np.load(os.path.join(GOLD, "traj_fmm_pefrl_n2048.npz"))
    s = g["state0"].copy()
    # the reference CPU path rebuilds the tree at every evaluation and tests leaves first
    nb.Context(order=3, unsort=0, tree_steps=1, m2l_first=0).run_host(int(g["scheme"]), nb.EVAL_COULOMB_FMM3_KD, s, g["param"], 5e-4, int(g["steps"]))
    assert np.abs(s[0] - g["final"][0]).max() <= 2e-6 * np.abs(g["final"][0]).max()
    assert np.abs(s[1] - g["final"][1]).max() <= 5e-5 * np.abs(g["final"][1]).max()


def test_energy_and_relerr_match_oracle():
    import torch
    n = 3000
    st = nb.init_ga(n)
    par = nb.default_param(n)
    buf = np.zeros(9 * n, np.float32)
    buf[:6 * n] = st.ravel()
    d = torch.from_numpy(buf).cuda()
    dpar = torch.from_numpy(par).cuda()
    ctx = nb.Context()
    e = np.array(ctx.energy(d.data_ptr(), n, dpar.data_ptr()))
    o = Oracle().energy(buf, n, par)
    assert np.allclose(e, o, rtol=2e-6)
    a = torch.from_numpy(st[0].copy()).cuda()
    b = torch.from_numpy((st[0] * 1.001).astype(np.float32)).cuda()
    m, mx = ctx.mean_rel_err(a.data_ptr(), b.data_ptr(), n)
    wm, wmx = mean_rel_err(st[0], (st[0] * 1.001).astype(np.float32))
    assert abs(m - wm) < 1e-6 * wm + 1e-12 and abs(mx - wmx) < 1e-5 * wmx


def test_energy_drift_fmm_no_worse_than_reference_algorithm():
    """leapfrog, dt = 5e-4, 100 steps, N = 8192 (the survey's drift experiment):
    |H - H0| / H0 of the GPU path stays within the reference's own drift class"""
    import torch
    n = 8192
    st = nb.init_ga(n)
    par = nb.default_param(n)
    ctx = nb.Context(order=3, unsort=0, tree_steps=8)
    buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
    buf[:6 * n] = torch.from_numpy(st.ravel()).cuda()
    dpar = torch.from_numpy(par).cuda()
    ctx.compute_force(nb.EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, dpar.data_ptr())
    h0 = sum(ctx.energy(buf.data_ptr(), n, dpar.data_ptr()))
    ctx.integrate(nb.LEAPFROG, nb.EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, 100)
    h1 = sum(ctx.energy(buf.data_ptr(), n, dpar.data_ptr()))
    assert abs(h0 - 1.9998) < 2e-3            # H(step 1) = 1.99980 in the survey's run
    assert abs(h1 - h0) / h0 < 3e-4           # reference p = 3: 4.4e-5 after 100, 2.7e-4 after 200 steps


@pytest.mark.parametrize("evaluator", [nb.EVAL_COULOMB_DIRECT3, nb.EVAL_COULOMB_FMM3_KD])
def test_step_host_matches_device_integration(evaluator):
    """nbco_step_host (H2D, one step, D2H; the read-back of the positions overlaps the force evaluation on the steps
    that do not rebuild the tree) called step by step = nbco_integrate on a resident state"""
    import torch
    n, steps = 30000, 10
    st = nb.init_ga(n)
    par = nb.default_param(n)
    dpar = torch.from_numpy(par).cuda()
    c1 = nb.Context(order=3, unsort=0, tree_steps=4)
    b1 = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
    b1[:6 * n] = torch.from_numpy(st.ravel()).cuda()
    c1.compute_force(evaluator, b1.data_ptr(), n, dpar.data_ptr())
    h = torch.empty(9 * n, dtype=torch.float32).pin_memory()
    h.copy_(b1.cpu())
    c1.integrate(nb.LEAPFROG, evaluator, b1.data_ptr(), n, dpar.data_ptr(), 5e-4, steps)
    want = b1.cpu().numpy().reshape(3, n, 3)
    c2 = nb.Context(order=3, unsort=0, tree_steps=4)
    if evaluator == nb.EVAL_COULOMB_FMM3_KD:
        # same tree phase as c1: the state on the host is already in tree order, evaluation 0 has been done there
        c2.eval_host(nb.EVAL_FMM3_KD, st[0].copy(), st[1].copy(), par)
    for _ in range(steps):
        c2.step_host(nb.LEAPFROG, evaluator, h.numpy(), n, par, 5e-4, 1)
    got = h.numpy().reshape(3, n, 3)
    if evaluator == nb.EVAL_COULOMB_DIRECT3:
        assert np.array_equal(got, want)
    else:
        o = np.lexsort((want[0][:, 2], want[0][:, 1], want[0][:, 0]))
        g = np.lexsort((got[0][:, 2], got[0][:, 1], got[0][:, 0]))
        assert np.abs(got[0][g] - want[0][o]).max() <= 1e-6 * np.abs(want[0]).max()
        assert np.abs(got[1][g] - want[1][o]).max() <= 1e-5 * np.abs(want[1]).max()


@pytest.mark.parametrize("scheme", [nb.EULER, nb.LEAPFROG, nb.FORESTRUTH, nb.PEFRL])
def test_fused_energy_reduction_matches_the_separate_pass(scheme):
    """nbco_integrate_energy: the last kick / drift of the integration also reduces 1/2 v^2 and 1/2 k x^2 of the state it
    writes; same state bits as nbco_integrate, same sums as the separate nbco_energy pass (double accumulation)"""
    import torch
    n, steps = 30000, 3
    st = nb.init_ga(n)
    par = torch.from_numpy(nb.default_param(n)).cuda()
    ev = nb.EVAL_COULOMB_DIRECT3          # deterministic evaluator: the two runs must agree bit for bit
    bufs = []
    for fused in (0, 1):
        c = nb.Context()
        b = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
        b[:6 * n] = torch.from_numpy(st.ravel()).cuda()
        c.compute_force(ev, b.data_ptr(), n, par.data_ptr())
        if fused:
            ke, el = c.integrate_energy(scheme, ev, b.data_ptr(), n, par.data_ptr(), 5e-4, steps)
        else:
            c.integrate(scheme, ev, b.data_ptr(), n, par.data_ptr(), 5e-4, steps)
        bufs.append(b)
    assert torch.equal(bufs[0][:6 * n], bufs[1][:6 * n])
    e = nb.Context().energy(bufs[1].data_ptr(), n, par.data_ptr())
    assert abs(ke - e[0]) <= 1e-12 * abs(e[0]) and abs(el - e[1]) <= 1e-12 * abs(e[1])
