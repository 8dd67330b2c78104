"""numpy statement of the harmonic (complex) form of the reference's 2D operators that csrc/fmm2.cu
implements.  Test helper: tests/test_ops2d_host.py pins it to the oracle's Cartesian-tensor form
(oracle/nbco_oracle2d.c, following fmm_cart_base.cuh), so that the formulas the CUDA kernels use are
checked on the CPU as well.

With z = x + i y, a symmetric order-q multipole M_q (q+1 entries, fmm_cart_base.cuh:111) enters every
contraction with the traceless gradient only through  Z_q = sum_k binom(q,k) i^k M_q[k]
(= sum_particles (-z)^q / q!), and a traceless order-n local (2 entries, :116) is the complex number
L_n = L_n[0] + i L_n[1].  Then
  P2M   Z_q  = sum_j (-(z_j - c))^q / q!
  M2M   Z'_n = sum_m Z_{n-m} (c_child - c_parent)... see m2m()
  M2L   L_n += 1/n! sum_q conj(Z_q) g_{n+q},  g_m = (-1)^m (m-1)! E_m / r^m,  r^2 = dx^2+dy^2+eps2, see m2l()
  L2L   L'_q = sum_{m>=q} binom(m,q) L_m conj(d)^(m-q)
  L2P   f    = -sum_n n L_n conj(d)^(n-1)
"""
from math import comb, factorial

import numpy as np


def sym_off(p):
    return p * (p + 1) // 2


def trl_off(p):
    return 0 if p == 0 else 2 * p - 1


def reduce_sym(mp, p):
    """symmetric tuples (nodes, offM) -> complex Z (nodes, p+1)"""
    Z = np.zeros((mp.shape[0], p + 1), complex)
    for q in range(p + 1):
        M = mp[:, sym_off(q):sym_off(q + 1)]
        for k in range(q + 1):
            Z[:, q] += comb(q, k) * (1j ** k) * M[:, k]
    return Z


def local_cx(lc, p):
    """traceless tuples (nodes, offL) -> complex (nodes, p+1); order 0 has a real entry only"""
    Lc = np.zeros((lc.shape[0], p + 1), complex)
    Lc[:, 0] = lc[:, 0]
    for n in range(1, p + 1):
        Lc[:, n] = lc[:, trl_off(n)] + 1j * lc[:, trl_off(n) + 1]
    return Lc


def p2m(z, c, p):
    Z = np.zeros(p + 1, complex)
    Z[0] = len(z)
    for q in range(2, p + 1):
        Z[q] = ((-(z - c)) ** q).sum() / factorial(q)
    return Z


def m2m(Zc, d, p):
    """shift child moments by d = c_parent - c_child (complex): Z'_n = sum_m Z_{n-m} d^m / m!"""
    out = np.zeros(p + 1, complex)
    for n in range(p + 1):
        for m in range(n + 1):
            out[n] += Zc[n - m] * d ** m / factorial(m)
    return out


def m2l(Zs, dz, p, eps2):
    """dz = c_target - c_source.  The reference normalises by r = sqrt(|dz|^2 + eps2) (fmm_cart.cuh:241-249), so
    its direction d = dz / r is NOT a unit vector and its gradient tuple is the Chebyshev pair
    g_m = (-1)^m (m-1)! r^-m (T_m(d0), d1 U_(m-1)(d0)); both entries obey E_(m+1) = 2 d0 E_m - E_(m-1)."""
    r = np.sqrt(dz.real ** 2 + dz.imag ** 2 + eps2)
    d0, d1 = dz.real / r, dz.imag / r
    E = [1.0 + 0j, d0 + 1j * d1]
    for m in range(2, 2 * p + 1):
        E.append(2 * d0 * E[m - 1] - E[m - 2])
    out = np.zeros(p + 1, complex)
    for n in range(p + 1):
        for q in range(p + 1):
            m = n + q
            if m == 0:
                continue
            g = (-1) ** m * factorial(m - 1) * E[m] / r ** m
            out[n] += np.conj(Zs[q]) * g / factorial(n)
    out[0] = out[0].real
    return out


def l2l(Lp, d, p):
    """d = c_child - c_parent"""
    out = np.zeros(p + 1, complex)
    for q in range(p + 1):
        for m in range(q, p + 1):
            out[q] += comb(m, q) * Lp[m] * np.conj(d) ** (m - q)
    out[0] = out[0].real
    return out


def l2p(Lc, d, p):
    f = 0
    for n in range(1, p + 1):
        f -= n * Lc[n] * np.conj(d) ** (n - 1)
    return f
