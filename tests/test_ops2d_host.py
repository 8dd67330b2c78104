"""The harmonic (complex) form of the 2D operators that csrc/fmm2.cu implements, stated in numpy
(tests/cx2d.py), pinned to the oracle's Cartesian-tensor form (fmm_cart_base.cuh restated in
oracle/nbco_oracle2d.c) on a real tree.  CPU only."""
import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
import cx2d
from refs2d import Oracle2


@pytest.mark.parametrize("n,p", [(2500, 4), (1500, 7)])
def test_complex_operators_match_oracle(n, p):
    st = nb.init_ga2(n)
    eps2 = 1e-18
    r = Oracle2(order=p, eps2=eps2).fmm(st[0], None, None)
    L = r["levels"]
    tb = lambda l: (4 ** l - 1) // 3
    ntot = tb(L + 1)
    Z, Lc = cx2d.reduce_sym(r["mpole"], p), cx2d.local_cx(r["local"], p)
    c = r["center"][:, 0] + 1j * r["center"][:, 1]
    z = r["pos"][:, 0] + 1j * r["pos"][:, 1]
    mult, idx = r["mult"], r["index"]
    scale = np.abs(Z[tb(2):]).max(0) + 1e-300
    # P2M
    for i in range(tb(L), ntot):
        if mult[i]:
            assert np.all(np.abs(cx2d.p2m(z[idx[i]:idx[i] + mult[i]], c[i], p) - Z[i]) <= 1e-13 * scale)
    # M2M
    for l in range(L - 1, 1, -1):
        sl, slp = 1 << l, 2 << l
        for ij0 in range(sl * sl):
            i, j = divmod(ij0, sl)
            ij, ijp = tb(l) + ij0, tb(l + 1) + 2 * (i * slp + j)
            if not mult[ij]:
                continue
            out = sum(cx2d.m2m(Z[ch], c[ij] - c[ch], p) for ch in (ijp, ijp + 1, ijp + slp, ijp + slp + 1))
            out[1] = 0
            assert np.all(np.abs(out - Z[ij]) <= 1e-12 * scale), (l, ij0)
    # M2L + L2L
    mine = np.zeros((ntot, p + 1), complex)
    rad = 1
    for l in range(2, L + 1):
        sl = 1 << l
        for ij in range(sl * sl):
            i, j = divmod(ij, sl)
            t = tb(l) + ij
            if mult[t] > 0:
                im, jm = i // 2 * 2, j // 2 * 2
                for k in range(max(im - 2 * rad, 0), min(im + 2 * rad + 1, sl - 1) + 1):
                    for g in range(max(jm - 2 * rad, 0), min(jm + 2 * rad + 1, sl - 1) + 1):
                        if not (k > i + rad or k < i - rad or g > j + rad or g < j - rad):
                            continue
                        s = tb(l) + k * sl + g
                        if mult[s]:
                            mine[t] += cx2d.m2l(Z[s], c[t] - c[s], p, eps2)
            if l > 2:
                par = tb(l - 1) + (i // 2) * (sl // 2) + (j // 2)
                mine[t] += cx2d.l2l(mine[par], c[t] - c[par], p)
        sel = slice(tb(l), tb(l + 1))
        err = (np.abs(mine[sel] - Lc[sel]) / np.abs(Lc[sel]).max(0)).max()
        assert err <= 1e-12, (l, err)
    # L2P + near field = acc
    side = 1 << L
    far = np.zeros(n, complex)
    for cell in range(side * side):
        node = tb(L) + cell
        for q in range(idx[node], idx[node] + mult[node]):
            far[q] = cx2d.l2p(Lc[node], z[q] - c[node], p)
    o2 = Oracle2(order=p, eps2=eps2, coll=0).fmm(st[0], None, None)
    ref = o2["acc"][:, 0] + 1j * o2["acc"][:, 1]
    assert np.array_equal(o2["perm"], r["perm"])
    print("far-field rel err", np.abs(far - ref).max() / np.abs(ref).max())
    assert np.abs(far - ref).max() <= 1e-11 * np.abs(ref).max()
