"""Host-side pieces of the 2D `nbco` surface (no GPU): default beam of main.cu:294-313, KV / Gaussian samplers
(main.cu:120-170), fp64 state files, level rule (fmm_cart.cuh:416-418)."""
import ctypes as C
import os

import numpy as np

import coulomb_oscillators_b200 as nb
from refs2d import Oracle2, Ref2


def test_default_beam_is_rms_matched():
    """the quartic of main.cu:300-309 makes the beam matched in BOTH planes: w0^2 - w^2 = 2 xi / (A (Ax + Ay)) (KV envelope
    equations), with emittances w A^2 / 4 = (0.03e-3, 0.01e-3)"""
    b = nb.beam_params2()
    w0, A, om = np.array(nb.OMEGA0_2D), b["A"], b["omega"]
    dom = (w0 + om) * (w0 - om)
    assert abs(om[1] / w0[1] - 0.8) < 1e-15                                   # tune depression in y (:295)
    assert np.allclose(dom * A * (A[0] + A[1]) / 2, b["xi"], rtol=1e-12)     # same perveance from x and from y
    assert np.allclose(om * A * A / 4, nb.EMIT_2D, rtol=1e-12)


def test_samplers_moments_and_determinism():
    n = 20000
    b = nb.beam_params2()
    kv, ga = nb.init_kv2(n), nb.init_ga2(n)
    for st, x, u in ((kv, b["A"] / 2, b["omega"] * b["A"] / 2), (ga, b["A"] / 2, b["omega"] * b["A"] / 2)):
        assert np.abs(st[0].mean(0)).max() < 1e-18 and np.abs(st[1].mean(0)).max() < 1e-16   # centerDist
        assert np.allclose(np.sqrt((st[0] ** 2).mean(0)), x, rtol=1e-12)                    # adjustRMS
        assert np.allclose(np.sqrt((st[1] ** 2).mean(0)), u, rtol=1e-12)
    assert np.array_equal(kv, nb.init_kv2(n))                                                 # fixed seed (main.cu:779-780)
    # a KV beam fills an ellipse of semi-axes ~A uniformly: no particle far outside it
    r2 = (kv[0] / b["A"]) ** 2
    assert r2.sum(1).max() < 1.05


def test_state_file_roundtrip_fp64(tmp_path):
    n = 777
    st = nb.init_ga2(n)
    path = str(tmp_path / "out0_0.000500.bin").encode()
    assert nb.lib.nbco_state_write2(path, st.ctypes.data_as(C.c_void_p), n) == 0
    assert os.path.getsize(path) == 32 * n                       # n = bytes / 2 / sizeof(double2)
    assert np.array_equal(np.fromfile(path, np.float64).reshape(2, n, 2), st)
    p, m = C.c_void_p(), C.c_int64()
    assert nb.lib.nbco_state_read2(path, C.byref(p), C.byref(m)) == 0 and m.value == n
    back = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(4 * n,)).copy()
    nb.lib.nbco_free(p)
    assert np.array_equal(back.reshape(2, n, 2), st)


def test_level_rule_matches_reference():
    for n, p in [(1000, 1), (30001, 5), (1 << 20, 3), (1 << 22, 5), (5, 10)]:
        L = nb.fmm2_levels(n, p)
        assert L == Oracle2(order=p).levels(n) and L >= 2
        if Ref2.available():
            assert L == Ref2(order=p, threads=1).levels(n)
    assert nb.fmm2_levels(1 << 22, 5) == 9 and nb.fmm2_levels(1 << 22, 5, 4.0) == 10
