"""The unrolled operator templates the CUDA kernels instantiate (csrc/fmm_ops.cuh) are compiled
for the host and compared with the oracle's runtime-loop operators, order 1..6."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from refs import ORACLE_SO, ROOT

f = np.ctypeslib.ndpointer(np.float32, flags="C")


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ops") / "libops_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                    os.path.join(ROOT, "tests", "ops_host.cpp")], check=True)
    H, O = C.CDLL(so), C.CDLL(ORACLE_SO)
    for L, pre in ((H, "ops_"), (O, "orc_op_")):
        getattr(L, pre + "p2m").argtypes = [f, C.c_int, f]
        getattr(L, pre + "m2m").argtypes = [f, f, C.c_int, f]
        getattr(L, pre + "m2l").argtypes = [f, f, C.c_int, f, C.c_float]
        getattr(L, pre + "l2l").argtypes = [f, f, C.c_int, f]
        getattr(L, pre + "l2p").argtypes = [f, f, C.c_int, f]
    return H, O


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_operators_match_oracle(libs, p):
    H, O = libs
    rng = np.random.default_rng(p)
    offM, offL = p * (p + 1) * (p + 2) // 6, (p + 1) ** 2
    for trial in range(5):
        d = (rng.normal(size=3) * 0.3).astype(np.float32)
        a, b = np.zeros(offM, np.float32), np.zeros(offM, np.float32)
        O.orc_op_p2m(a, p, d); H.ops_p2m(b, p, d)
        assert rel(b, a) < 1e-6
        Mi = rng.normal(size=offM).astype(np.float32)
        Mi[1:4] = 0  # dipole of a centre-of-charge expansion
        a[:] = 0; b[:] = 0
        O.orc_op_m2m(a, Mi, p, d); H.ops_m2m(b, Mi, p, d)
        assert rel(b, a) < 2e-6
        dd = rng.normal(size=3).astype(np.float32)
        r = np.float32(np.sqrt((dd * dd).sum() + 1e-3))
        u = (dd / r).astype(np.float32)
        la, lb = np.zeros(offL, np.float32), np.zeros(offL, np.float32)
        O.orc_op_m2l(la, Mi, p, u, r); H.ops_m2l(lb, Mi, p, u, r)
        assert rel(lb, la) < 5e-6 and la[0] == 0 and lb[0] == 0
        Lp = rng.normal(size=offL).astype(np.float32)
        la[:] = 0; lb[:] = 0
        O.orc_op_l2l(la, Lp, p, d); H.ops_l2l(lb, Lp, p, d)
        assert rel(lb, la) < 2e-6
        fa, fb = np.zeros(3, np.float32), np.zeros(3, np.float32)
        O.orc_op_l2p(fa, Lp, p, d); H.ops_l2p(fb, Lp, p, d)
        assert rel(fb, fa) < 2e-6


def test_m2l_is_gradient_of_inverse_distance(libs):
    """independent check of the maths: L2P(M2L(monopole)) reproduces the Coulomb field d/|d|^3"""
    H, _ = libs
    p = 6
    offM, offL = p * (p + 1) * (p + 2) // 6, (p + 1) ** 2
    M = np.zeros(offM, np.float32); M[0] = 1
    c = np.array([1.0, -0.7, 0.4], np.float32)            # target centre - source centre
    r = np.float32(np.linalg.norm(c))
    L = np.zeros(offL, np.float32)
    H.ops_m2l(L, M, p, (c / r).astype(np.float32), r)
    x = np.array([0.05, -0.03, 0.04], np.float32)          # offset from the target centre
    fld = np.zeros(3, np.float32)
    H.ops_l2p(fld, L, p, x)
    exact = (c + x) / np.linalg.norm(c + x) ** 3
    assert np.allclose(fld, exact, rtol=2e-5)
