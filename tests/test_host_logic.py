"""Host-side pieces of the path: shard ranges, initial conditions, state files, kd depth rule."""
import os

import numpy as np

import coulomb_oscillators_b200 as nb
from refs import Oracle, Ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_shard_range_is_the_kd_split():
    # ceil(n*i/w), fmm_cart3_kdtree.cuh:117-118; shards tile [0, n) without gaps
    for n in (1, 7, 8192, 100003, 1 << 24):
        for w in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(w):
                b, e = nb.shard_range(n, r, w)
                assert b == prev and e == -(-n * (r + 1) // w)
                prev = e
            assert prev == n


def test_init_ga_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "fmm_ga_n3000_p3.npz"))
    st = nb.init_ga(3000)
    assert np.array_equal(st[0], g["pos"]) and np.array_equal(st[1], g["vel"])
    # first values printed by the reference build for n=8192 (recorded in DESIGN.md)
    s = nb.init_ga(8192)
    assert np.allclose(s[0, 0], [-9.2383561e-04, -8.4511550e-05, -5.6287115e-03], rtol=1e-6)
    assert np.allclose(s[0].std(0), [0.003, 0.001, 0.01], rtol=1e-3)


def test_init_test_cube_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "fmm_cube_n4096_p4.npz"))
    st = nb.init_test_cube(4096)
    assert np.array_equal(st[0], g["pos"])
    assert np.abs(st[0]).max() <= 1.05


def test_init_matches_live_reference():
    if not Ref.available():
        import pytest
        pytest.skip("oracle/_ref not built")
    L = Ref.lib()
    n = 1000
    x = np.array([0.003, 0.001, 0.01], np.float32)
    u = (x * np.array([1.095, 1, 1], np.float32)).astype(np.float32)
    buf = np.zeros(6 * n, np.float32)
    L.ref_init_ga(buf, n, x, u)
    assert np.array_equal(buf.reshape(2, n, 3), nb.init_ga(n))
    L.ref_init_test_cube(buf, n, x, u)
    assert np.array_equal(buf.reshape(2, n, 3), nb.init_test_cube(n))


def test_state_file_round_trip(tmp_path):
    import ctypes as C
    st = nb.init_ga(777)
    p = str(tmp_path / "out0_0.000500.bin").encode()
    assert nb.lib.nbco_state_write(p, st.ctypes.data_as(C.c_void_p), 777) == 0
    raw = np.fromfile(p.decode(), np.float32)
    assert raw.size == 6 * 777 and np.array_equal(raw, st.ravel())  # all positions, then all velocities
    ptr, n = C.c_void_p(), C.c_int64()
    assert nb.lib.nbco_state_read(p, C.byref(ptr), C.byref(n)) == 0
    back = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(6 * 777,)).copy()
    nb.lib.nbco_free(ptr)
    assert n.value == 777 and np.array_equal(back, st.ravel())


def test_kd_depth_rule():
    # fmm_cart3_kdtree.cuh:1507-1516: L = clamp(round(log2(dens*n/p^2)), 2, 30), then 2^L <= n
    L = Oracle.lib().orc_kd_levels
    assert L(1 << 20, 3, 1.0, 0) == 17 and L(1 << 24, 3, 1.0, 0) == 21
    assert L(1 << 20, 5, 1.0, 0) == 15
    assert L(8192, 3, 1.0, 0) == 10
    assert L(9, 1, 1.0, 0) == 3 and L(8, 3, 1.0, 0) == 2
    assert L(1000, 3, 1.0, 20) == 9   # -maxlevel is clamped to 2^L <= n
