"""ctypes access to the 2D fp64 CHECKERS: oracle/libnbco_oracle.so (orc2_*, our C restatement of
fmm_cart_cpu) and oracle/_ref/libnbco_ref2d.so (the unmodified reference headers compiled with
SCAL = double, DIM = 2 behind oracle/ref_harness2d.cu).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

from refs import ORACLE_SO, ROOT

REF2D_SO = os.path.join(ROOT, "oracle", "_ref", "libnbco_ref2d.so")
vp = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def sym_off(p):
    return p * (p + 1) // 2


def trl_off(p):
    return 0 if p == 0 else 2 * p - 1


class Oracle2:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(ORACLE_SO)
            L.orc2_levels.argtypes = [C.c_int, C.c_int, C.c_double]
            L.orc2_fmm.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int] + [vp] * 6
            L.orc2_direct.argtypes = [vp, vp, C.c_int, vp, C.c_double]
            L.orc2_eval.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
            L.orc2_integrate.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_double, C.c_longlong, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_int]
            L.orc2_energy.argtypes = [vp, C.c_int, vp, C.c_double, vp]
            cls._lib = L
        return cls._lib

    def __init__(self, order=3, radius=1, eps2=1e-18, dens=1.0, coll=1):
        self.L = self.lib()
        self.order, self.radius, self.eps2, self.dens, self.coll = order, radius, eps2, dens, coll

    def levels(self, n):
        return self.L.orc2_levels(n, self.order, self.dens)

    def fmm(self, pos, vel=None, param=None):
        """returns dict: pos/vel sorted (copies), acc (cell order), perm, center, mpole, local, mult, index, levels"""
        n, p = pos.shape[0], self.order
        Lv = self.levels(n)
        ntot = (4 ** (Lv + 1) - 1) // 3
        r = dict(pos=np.array(pos, np.float64), vel=None if vel is None else np.array(vel, np.float64),
                 acc=np.zeros((n, 2)), perm=np.zeros(n, np.int32), center=np.zeros((ntot, 2)),
                 mpole=np.zeros((ntot, sym_off(p + 1))), local=np.zeros((ntot, trl_off(p + 1))),
                 mult=np.zeros(ntot, np.int32), index=np.zeros(ntot, np.int32))
        got = self.L.orc2_fmm(_p(r["pos"]), _p(r["vel"]), _p(r["acc"]), n, _p(param), p, self.radius, self.eps2, self.dens,
                              self.coll, _p(r["perm"]), _p(r["center"]), _p(r["mpole"]), _p(r["local"]), _p(r["mult"]), _p(r["index"]))
        assert got == Lv
        r["levels"] = Lv
        return r

    def direct(self, pos, param=None):
        n = pos.shape[0]
        a = np.zeros((n, 2))
        self.L.orc2_direct(_p(np.ascontiguousarray(pos, np.float64)), _p(a), n, _p(param), self.eps2)
        return a

    def eval(self, evaluator, buf, n, param=None):
        assert self.L.orc2_eval(evaluator, _p(buf), n, _p(param), self.order, self.radius, self.eps2, self.dens, self.coll) == 0

    def integrate(self, scheme, evaluator, buf, n, param, dt, nsteps):
        assert self.L.orc2_integrate(scheme, evaluator, _p(buf), n, _p(param), dt, nsteps, self.order, self.radius,
                                     self.eps2, self.dens, self.coll) == 0

    def energy(self, buf, n, param=None):
        out = np.zeros(3)
        self.L.orc2_energy(_p(buf), n, _p(param), self.eps2, _p(out))
        return out


class Ref2:
    """the unmodified 2D reference (CPU path); evaluator 0 direct2_cpu, 1 fmm_cart_cpu, 2/3 + elastic term"""
    _lib = None

    @staticmethod
    def available():
        return os.path.exists(REF2D_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF2D_SO)
            L.ref2_config.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
            L.ref2_eval.argtypes = [C.c_int, vp, C.c_int, vp]
            L.ref2_integrate.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_double, C.c_int]
            L.ref2_levels.argtypes = [C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, order=3, radius=1, eps2=1e-18, dens=1.0, threads=None, coll=1):
        self.L = self.lib()
        self.cfg = (order, float(radius), eps2, dens, threads or min(os.cpu_count() or 8, 64), coll)

    def eval(self, which, buf, n, param=None):
        self.L.ref2_config(*self.cfg)
        assert self.L.ref2_eval(which, _p(buf), n, _p(param)) == 0

    def integrate(self, scheme, which, buf, n, param, dt, nsteps):
        self.L.ref2_config(*self.cfg)
        assert self.L.ref2_integrate(scheme, which, _p(buf), n, _p(param), dt, nsteps) == 0

    def levels(self, n):
        self.L.ref2_config(*self.cfg)
        return self.L.ref2_levels(n)


def by_position(pos):
    """canonical order of a particle set (the reference's CPU sort is unstable inside a cell)"""
    return np.lexsort((pos[:, 1], pos[:, 0]))


def rel_err2(a, ref):
    """rel_diff1 (reductions.cuh:37-42) per particle: (mean, max)"""
    e = np.sqrt(((a - ref) ** 2).sum(1) / ((ref ** 2).sum(1) + 1e-18))
    return float(e.mean()), float(e.max())
