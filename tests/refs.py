"""ctypes access to the two CHECKERS: oracle/libnbco_oracle.so (our C restatement) and
oracle/_ref/libnbco_ref.so (the unmodified reference compiled behind oracle/ref_harness.cu).
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libnbco_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnbco_ref.so")

_f = np.ctypeslib.ndpointer(np.float32, flags="C")
_i = np.ctypeslib.ndpointer(np.int32, flags="C")
_l = np.ctypeslib.ndpointer(np.int64, flags="C")
_d = np.ctypeslib.ndpointer(np.float64, flags="C")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """our plain-C restatement (oracle/nbco_oracle.c)"""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(ORACLE_SO)
            L.orc_create.restype = C.c_void_p
            L.orc_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            L.orc_destroy.argtypes = [C.c_void_p]
            L.orc_fmm3_kd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.orc_direct3.argtypes = [_f, _f, C.c_int, C.c_void_p, C.c_float]
            L.orc_add_elastic.argtypes = [_f, _f, C.c_int, C.c_void_p]
            L.orc_step.argtypes = [_f, _f, C.c_float, C.c_int]
            L.orc_eval.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_void_p]
            L.orc_integrate.argtypes = [C.c_void_p, C.c_int, C.c_int, _f, C.c_int, C.c_void_p, C.c_double, C.c_longlong]
            L.orc_energy.argtypes = [_f, C.c_int, C.c_void_p, C.c_float, _d]
            L.orc_mean_rel_err.restype = C.c_double
            L.orc_mean_rel_err.argtypes = [_f, _f, C.c_int, C.POINTER(C.c_double)]
            L.orc_info.argtypes = [C.c_void_p, _l]
            L.orc_get_tree.argtypes = [C.c_void_p] + [C.c_void_p] * 9
            L.orc_get_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.orc_kd_levels.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, order=3, radius=1.0, eps2=1e-18, dens_inhom=1.0, max_level=0, tree_steps=1,
                 coll=1, unsort=1, m2l_first=0):
        self.L = self.lib()
        self.h = self.L.orc_create(order, radius, eps2, dens_inhom, max_level, tree_steps, coll, unsort, m2l_first)
        assert self.h
        self.eps2 = eps2

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def fmm3_kd(self, pos, vel=None, param=None):
        n = pos.shape[0]
        acc = np.empty((n, 3), np.float32)
        assert self.L.orc_fmm3_kd(self.h, _ptr(pos), _ptr(vel), _ptr(acc), n, _ptr(param)) == 0
        return acc

    def direct3(self, pos, param=None):
        n = pos.shape[0]
        acc = np.empty((n, 3), np.float32)
        self.L.orc_direct3(pos, acc, n, _ptr(param), self.eps2)
        return acc

    def eval(self, evaluator, buf, n, param=None):
        assert self.L.orc_eval(self.h, evaluator, buf, n, _ptr(param)) == 0

    def integrate(self, scheme, evaluator, buf, n, param, dt, nsteps):
        assert self.L.orc_integrate(self.h, scheme, evaluator, buf, n, _ptr(param), dt, nsteps) == 0

    def energy(self, buf, n, param=None):
        out = np.zeros(3)
        self.L.orc_energy(buf, n, _ptr(param), self.eps2, out)
        return out

    def info(self):
        o = np.zeros(9, np.int64)
        self.L.orc_info(self.h, o)
        return dict(levels=int(o[0]), order=int(o[1]), n=int(o[2]), nodes=int(o[3]), p2p_pairs=int(o[4]),
                    m2l_pairs=int(o[5]), off_m=int(o[6]), off_l=int(o[7]), rebuilt=int(o[8]))

    def tree(self):
        i = self.info()
        nn, n = i["nodes"], i["n"]
        t = dict(center=np.empty((nn, 3), np.float32), lbound=np.empty((nn, 3), np.float32),
                 rbound=np.empty((nn, 3), np.float32), mpole=np.empty((nn, i["off_m"]), np.float32),
                 local=np.empty((nn, i["off_l"]), np.float32), mult=np.empty(nn, np.int32),
                 index=np.empty(nn, np.int32), splitdim=np.empty(nn, np.int32), perm=np.empty(n, np.int32))
        self.L.orc_get_tree(self.h, *[_ptr(t[k]) for k in ("center", "lbound", "rbound", "mpole", "local", "mult", "index", "splitdim", "perm")])
        t["levels"] = i["levels"]
        return t

    def lists(self):
        i = self.info()
        p2p = np.empty((i["p2p_pairs"], 2), np.int32)
        m2l = np.empty((i["m2l_pairs"], 2), np.int32)
        self.L.orc_get_lists(self.h, _ptr(p2p), _ptr(m2l))
        return p2p, m2l


def mean_rel_err(x, ref):
    """rel_diff1 of reductions.cuh:37-42 in numpy (float64 accumulate): (mean, max)"""
    x = np.asarray(x, np.float32).reshape(-1, 3)
    ref = np.asarray(ref, np.float32).reshape(-1, 3)
    d2 = ((x - ref) ** 2).sum(1, dtype=np.float32)
    r2 = (ref ** 2).sum(1, dtype=np.float32) + np.float32(1e-18)
    e = np.sqrt(np.maximum(d2 / r2, 0).astype(np.float64))
    return float(e.mean()), float(e.max())


class Ref:
    """the unmodified reference (CPU path) behind oracle/ref_harness.cu; None-safe: Ref.available()"""
    _lib = None

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(REF_SO)
            L.ref_config.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
            L.ref_init_ga.argtypes = [_f, C.c_int, _f, _f]
            L.ref_init_test_cube.argtypes = [_f, C.c_int, _f, _f]
            L.ref_eval.argtypes = [C.c_int, _f, C.c_int, C.c_void_p]
            L.ref_integrate.argtypes = [C.c_int, C.c_int, _f, C.c_int, C.c_void_p, C.c_double, C.c_int]
            L.ref_mean_rel_err.restype = C.c_double
            L.ref_mean_rel_err.argtypes = [_f, _f, C.c_int]
            L.ref_kd_levels.argtypes = [C.c_int]
            L.ref_fmm3_phases.argtypes = [_f, _f, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 9 + [C.c_void_p, C.c_void_p, C.c_longlong, _l]
            L.ref_direct3_gpu_seconds.restype = C.c_double
            L.ref_direct3_gpu_seconds.argtypes = [_f, _f, C.c_int, _f, C.c_int]
            if hasattr(L, "ref_fmm3_gpu_step_seconds"):   # harness of round 2 (reference GPU path as a second baseline)
                L.ref_fmm3_gpu_step_seconds.restype = C.c_double
                L.ref_fmm3_gpu_step_seconds.argtypes = [_f, C.c_int, _f, C.c_double, C.c_int, C.c_int]
                L.ref_eval_gpu.argtypes = [C.c_int, _f, C.c_int, _f]
            cls._lib = L
        return cls._lib

    def __init__(self, order=3, radius=1.0, eps2=1e-18, dens_inhom=1.0, threads=None, coll=1, unsort=1, tree_steps=8):
        self.L = self.lib()
        self.cfg = (order, radius, eps2, dens_inhom, threads or min(os.cpu_count() or 8, 64), coll, unsort, tree_steps)
        self.order = order
        self.apply()

    def apply(self):
        self.L.ref_config(*self.cfg)

    @classmethod
    def init_ga(cls, n, sigma_x=(0.003, 0.001, 0.01), omega0=(1.095, 1.0, 1.0)):
        """the reference's own initGA (main3.cu:114-137,662-664) -> (2, n, 3) float32 [pos, vel]"""
        x3 = np.asarray(sigma_x, np.float32)
        u3 = (np.asarray(omega0, np.float32) * x3).astype(np.float32)   # main3.cu:241-245
        buf = np.zeros(6 * n, np.float32)
        cls.lib().ref_init_ga(buf, n, x3, u3)
        return buf.reshape(2, n, 3)

    def eval(self, which, buf, n, param=None):
        self.apply()
        assert self.L.ref_eval(which, buf, n, _ptr(param)) == 0

    def integrate(self, scheme, which, buf, n, param, dt, nsteps):
        self.apply()
        assert self.L.ref_integrate(scheme, which, buf, n, _ptr(param), dt, nsteps) == 0

    def fmm3_phases(self, pos, param=None, m2l_first=0):
        """returns dict with sorted pos, acc (tree order), tree arrays, sorted-unique lists"""
        self.apply()
        n = pos.shape[0]
        L = self.L.ref_kd_levels(n)
        nn = (1 << (L + 1)) - 1
        p = self.order
        off_m, off_l = p * (p + 1) * (p + 2) // 6, (p + 1) ** 2
        t = dict(center=np.zeros((nn, 3), np.float32), lbound=np.zeros((nn, 3), np.float32),
                 rbound=np.zeros((nn, 3), np.float32), mpole=np.zeros((nn, off_m), np.float32),
                 local=np.zeros((nn, off_l), np.float32), mult=np.zeros(nn + 1, np.int32),
                 index=np.zeros(nn + 1, np.int32), splitdim=np.zeros(nn, np.int32), perm=np.zeros(n, np.int32))
        cap = max(64 * nn, 1 << 16)
        p2p = np.zeros((cap, 2), np.int32)
        m2l = np.zeros((cap, 2), np.int32)
        counts = np.zeros(2, np.int64)
        spos = np.ascontiguousarray(pos, np.float32).copy()
        acc = np.zeros((n, 3), np.float32)
        r = self.L.ref_fmm3_phases(spos, acc, n, _ptr(param), m2l_first, _ptr(t["perm"]), _ptr(t["center"]), _ptr(t["lbound"]),
                                   _ptr(t["rbound"]), _ptr(t["mpole"]), _ptr(t["local"]), _ptr(t["mult"]), _ptr(t["index"]),
                                   _ptr(t["splitdim"]), _ptr(p2p), _ptr(m2l), cap, counts)
        assert r == L, r
        t["mult"] = t["mult"][:nn]
        t["index"] = t["index"][:nn]
        t["levels"] = L
        t["pos_sorted"] = spos
        t["acc_sorted"] = acc

        def canon(a, k):
            a = a[:k].copy()
            o = np.lexsort((a[:, 1], a[:, 0]))
            return a[o]
        t["p2p"] = canon(p2p, counts[0])
        t["m2l"] = canon(m2l, counts[1])
        return t


REF64_SO = os.path.join(ROOT, "oracle", "_ref", "libnbco_ref_f64.so")


class Ref64:
    """the unmodified reference compiled with SCAL = double (oracle/ref_harness_f64.cu): fp64 ground truth"""
    _lib = None

    @staticmethod
    def available():
        return os.path.exists(REF64_SO)

    def __init__(self, order=3, radius=1.0, eps2=1e-18, dens_inhom=1.0, threads=None, coll=1):
        if Ref64._lib is None:
            L = C.CDLL(REF64_SO)
            L.ref64_config.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
            L.ref64_eval.argtypes = [C.c_int, _d, _d, C.c_int, C.c_void_p]
            Ref64._lib = L
        self.cfg = (order, radius, eps2, dens_inhom, threads or min(os.cpu_count() or 8, 64), coll)

    def _eval(self, which, pos, param):
        n = pos.shape[0]
        p64 = np.ascontiguousarray(pos, np.float64).ravel()
        acc = np.zeros(3 * n, np.float64)
        par = None if param is None else np.ascontiguousarray(param, np.float64)
        self._lib.ref64_config(*self.cfg)
        assert self._lib.ref64_eval(which, p64, acc, n, _ptr(par)) == 0
        return acc.reshape(n, 3)

    def fmm3(self, pos, param=None):
        return self._eval(1, pos, param)

    def direct3(self, pos, param=None):
        return self._eval(0, pos, param)


def unique_axes(pos):
    """make every coordinate of every axis distinct (bump duplicates by one ulp until strictly
    increasing): on such inputs the reference's unstable sorts have a unique answer.
    Vectorised: in the monotone integer image k of the floats the rule v[i] = max(v[i], v[i-1] + 1 ulp) is
    k'[i] = max(k[i], k'[i-1] + 1), i.e. k' - i = running maximum of k - i (usable at N = 2^24)."""
    pos = np.array(pos, np.float32, copy=True)
    for k in range(3):
        o = np.argsort(pos[:, k], kind="stable")
        v = pos[o, k].copy()
        v[v == 0] = 0.0                                      # -0.0 and +0.0 are one value
        b = v.view(np.int32).astype(np.int64)
        key = np.where(b >= 0, b, -(b & 0x7FFFFFFF))          # monotone in the float value, one step per ulp
        i = np.arange(len(v), dtype=np.int64)
        key = np.maximum.accumulate(key - i) + i
        nb_ = np.where(key >= 0, key, (-key) | 0x80000000).astype(np.uint32)
        v = nb_.view(np.float32)
        assert np.all(np.diff(v) > 0)
        pos[o, k] = v
    return pos
