"""The drop-in, compiled and run: oracle/_ref/libnbco_dropin.so is the UNMODIFIED reference translation unit (main3.cu)
plus include/nbco_shim.cuh (the ~40-line binding of INTEGRATION.md section 1) linked against libnbco.so.  The
reference's own host code -- compute_force + leapfrog (integrator.cuh:22-28,68-96) and test_accuracy
(main3.cu:139-182) -- then drives this repo's evaluators through the reference's function-pointer plugin type."""
import ctypes as C
import os

import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from refs import ROOT

pytestmark = pytest.mark.gpu
DROPIN_SO = os.path.join(ROOT, "oracle", "_ref", "libnbco_dropin.so")
_f = np.ctypeslib.ndpointer(np.float32, flags="C")


@pytest.fixture(scope="module")
def dropin():
    if not os.path.exists(DROPIN_SO):
        pytest.skip("oracle/_ref/libnbco_dropin.so not shipped")
    L = C.CDLL(DROPIN_SO)
    L.dropin_config.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
    L.dropin_leapfrog.argtypes = [_f, C.c_int, _f, C.c_double, C.c_int]
    L.dropin_test_accuracy.restype = C.c_float
    L.dropin_test_accuracy.argtypes = [_f, C.c_int, _f]
    return L


def test_reference_leapfrog_over_the_shim_matches_nbco_integrate(dropin):
    """main3.cu:835-846 with coulombOscillatorFMMKD3_b200 / step_b200: 10 of the reference's leapfrog steps (two tree
    rebuilds) against nbco_integrate on the same inputs"""
    n, steps = 30000, 10
    st = nb.init_ga(n)
    par = nb.default_param(n)
    buf = np.zeros(9 * n, np.float32)
    buf[:6 * n] = st.ravel()
    dropin.dropin_config(3, 1.0, 1e-18, 1.0, 1, 0, 8)        # the simulation mode of main3.cu:834 (b_unsort = false)
    assert dropin.dropin_leapfrog(buf, n, par, 5e-4, steps) == 0
    got = buf.reshape(3, n, 3)
    ctx = nb.Context(order=3, unsort=0, tree_steps=8, m2l_first=1)
    s = st.copy()
    ctx.run_host(nb.LEAPFROG, nb.EVAL_COULOMB_FMM3_KD, s, par, 5e-4, steps)
    assert np.abs(got[0] - s[0]).max() <= 1e-6 * np.abs(s[0]).max()      # same particle order, same schedule; fp32 atomic order differs
    assert np.abs(got[1] - s[1]).max() <= 1e-5 * np.abs(s[1]).max()


def test_reference_test_accuracy_over_the_shim(dropin):
    """the reference's own accuracy check (test_accuracy + relerrReduce2, main3.cu:139-182) fed with the shim's FMM and
    direct sum reports the error class of `nbco3 -test` (0.06 at p = 3 on the uniform cube)"""
    n = 8192
    st = nb.init_test_cube(n)
    par = nb.default_param(n)
    buf = np.zeros(9 * n, np.float32)
    buf[:6 * n] = st.ravel()
    dropin.dropin_config(3, 1.0, 1e-18, 1.0, 1, 1, 8)
    err = dropin.dropin_test_accuracy(buf, n, par)
    assert 0.02 < err < 0.12, err
