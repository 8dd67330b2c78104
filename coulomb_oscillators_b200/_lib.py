"""ctypes binding of include/nbco.h (the C ABI).  No numerics here."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NBCO_LIB: an alternative build of the SAME C ABI (tools/ab_phases.py times an older build beside the current one)
lib_path = os.environ.get("NBCO_LIB") or os.path.join(_HERE, "libnbco.so")

EVAL_DIRECT3, EVAL_FMM3_KD, EVAL_COULOMB_DIRECT3, EVAL_COULOMB_FMM3_KD = 0, 1, 2, 3
EVAL_DIRECT2, EVAL_FMM2, EVAL_COULOMB_DIRECT2, EVAL_COULOMB_FMM2 = 4, 5, 6, 7
EULER, LEAPFROG, FORESTRUTH, PEFRL = 0, 1, 2, 3


class NbcoError(RuntimeError):
    pass


class Config(C.Structure):
    """nbco_config (include/nbco.h); replaces the globals of reference constants.cuh:36-52."""
    _fields_ = [("device", C.c_int32), ("order", C.c_int32), ("radius", C.c_float), ("eps2", C.c_float),
                ("dens_inhom", C.c_float), ("max_level", C.c_int32), ("tree_steps", C.c_int32),
                ("coll", C.c_int32), ("unsort", C.c_int32), ("m2l_first", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32), ("eps2_d", C.c_double),
                ("reproducible", C.c_int32), ("reserved0", C.c_int32)]


class FmmInfo(C.Structure):
    _fields_ = [("levels", C.c_int32), ("order", C.c_int32), ("n", C.c_int64), ("nodes", C.c_int64),
                ("p2p_pairs", C.c_int64), ("m2l_pairs", C.c_int64), ("off_m", C.c_int32), ("off_l", C.c_int32),
                ("rebuilt", C.c_int32), ("mlt_max", C.c_int32), ("kernel_launches", C.c_int64),
                ("counter", C.c_int32), ("reserved", C.c_int32)]


class Fmm2Info(C.Structure):
    _fields_ = [("levels", C.c_int32), ("order", C.c_int32), ("n", C.c_int64), ("nodes", C.c_int64),
                ("coeffs", C.c_int32), ("reserved", C.c_int32), ("kernel_launches", C.c_int64), ("evals", C.c_int64)]


# every symbol include/nbco.h declares (tests/test_abi.py checks this list against the header)
SYMBOLS = [
    "nbco_default_config", "nbco_abi_version", "nbco_last_error", "nbco_create", "nbco_destroy",
    "nbco_set_config", "nbco_get_config", "nbco_stream", "nbco_force_direct3", "nbco_force_fmm3_kd",
    "nbco_coulomb_direct3", "nbco_coulomb_fmm3_kd", "nbco_add_elastic", "nbco_step", "nbco_compute_force",
    "nbco_integrate", "nbco_integrate_energy", "nbco_mean_rel_err", "nbco_energy", "nbco_eval_host", "nbco_run_host", "nbco_step_host",
    "nbco_fmm_get_info", "nbco_fmm_get_tree", "nbco_fmm_get_lists", "nbco_fmm_get_phase_ms", "nbco_fmm_phase_totals",
    "nbco_force_direct2", "nbco_force_fmm2", "nbco_coulomb_direct2", "nbco_coulomb_fmm2", "nbco_add_elastic2", "nbco_step2",
    "nbco_compute_force2", "nbco_integrate2", "nbco_mean_rel_err2", "nbco_energy2", "nbco_run_host2", "nbco_step_host2",
    "nbco_fmm2_levels", "nbco_fmm2_get_info", "nbco_fmm2_get_tree", "nbco_fmm2_get_phase_ms",
    "nbco_init_ga2", "nbco_init_kv2", "nbco_beam_params2", "nbco_state_read2", "nbco_state_write2",
    "nbco_peer_export", "nbco_peer_attach", "nbco_peer_attach_local", "nbco_peer_commit", "nbco_peer_barrier", "nbco_peer_gather", "nbco_peer_detach",
    "nbco_track_ids", "nbco_shard_range", "nbco_init_ga", "nbco_init_test_cube", "nbco_state_read", "nbco_state_write", "nbco_free",
]


def _load():
    if not os.path.exists(lib_path):
        raise NbcoError(f"{lib_path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no fallback implementation)")
    L = C.CDLL(lib_path)
    L.nbco_last_error.restype = C.c_char_p
    L.nbco_stream.restype = C.c_void_p
    L.nbco_stream.argtypes = [C.c_void_p]
    L.nbco_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.nbco_destroy.argtypes = [C.c_void_p]
    L.nbco_destroy.restype = None
    L.nbco_set_config.argtypes = [C.c_void_p, C.POINTER(Config)]
    L.nbco_get_config.argtypes = [C.c_void_p, C.POINTER(Config)]
    vp, i64, f32, f64 = C.c_void_p, C.c_int64, C.c_float, C.c_double
    L.nbco_force_direct3.argtypes = [vp, vp, vp, i64, vp]
    L.nbco_force_fmm3_kd.argtypes = [vp, vp, vp, i64, vp]
    L.nbco_coulomb_direct3.argtypes = [vp, vp, vp, i64, vp]
    L.nbco_coulomb_fmm3_kd.argtypes = [vp, vp, vp, i64, vp]
    L.nbco_add_elastic.argtypes = [vp, vp, vp, i64, vp]
    L.nbco_step.argtypes = [vp, vp, vp, f32, i64]
    L.nbco_compute_force.argtypes = [vp, C.c_int, vp, i64, vp]
    L.nbco_integrate.argtypes = [vp, C.c_int, C.c_int, vp, i64, vp, f64, i64]
    L.nbco_integrate_energy.argtypes = [vp, C.c_int, C.c_int, vp, i64, vp, f64, i64, C.POINTER(f64)]
    L.nbco_mean_rel_err.argtypes = [vp, vp, vp, i64, C.POINTER(f64), C.POINTER(f64)]
    L.nbco_energy.argtypes = [vp, vp, i64, vp, C.POINTER(f64)]
    L.nbco_eval_host.argtypes = [vp, C.c_int, vp, vp, vp, i64, vp]
    L.nbco_run_host.argtypes = [vp, C.c_int, C.c_int, vp, vp, i64, vp, f64, i64]
    L.nbco_step_host.argtypes = [vp, C.c_int, C.c_int, vp, i64, vp, f64, i64]
    L.nbco_track_ids.argtypes = [vp, vp]
    L.nbco_fmm_get_info.argtypes = [vp, C.POINTER(FmmInfo)]
    L.nbco_fmm_get_tree.argtypes = [vp] + [vp] * 9
    L.nbco_fmm_get_lists.argtypes = [vp, vp, i64, vp, i64]
    L.nbco_fmm_get_phase_ms.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(f32), C.c_int]
    L.nbco_fmm_phase_totals.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(f64), C.c_int, C.POINTER(i64), C.c_int]
    L.nbco_shard_range.argtypes = [i64, C.c_int32, C.c_int32, C.POINTER(i64), C.POINTER(i64)]
    L.nbco_shard_range.restype = None
    L.nbco_init_ga.argtypes = [vp, i64, vp, vp]
    L.nbco_init_test_cube.argtypes = [vp, i64, vp, vp]
    L.nbco_state_read.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(i64)]
    L.nbco_state_write.argtypes = [C.c_char_p, vp, i64]
    L.nbco_free.argtypes = [vp]
    L.nbco_peer_export.argtypes = [vp, i64, vp]
    L.nbco_peer_attach.argtypes = [vp, C.c_int32, vp]
    L.nbco_peer_attach_local.argtypes = [vp, C.c_int32, vp]
    L.nbco_peer_commit.argtypes = [vp]
    L.nbco_peer_barrier.argtypes = [vp]
    L.nbco_peer_gather.argtypes = [vp, vp, i64]
    L.nbco_peer_detach.argtypes = [vp]
    # 2D fp64 path
    for f in ("nbco_force_direct2", "nbco_force_fmm2", "nbco_coulomb_direct2", "nbco_coulomb_fmm2", "nbco_add_elastic2"):
        getattr(L, f).argtypes = [vp, vp, vp, i64, vp]
    L.nbco_step2.argtypes = [vp, vp, vp, f64, i64]
    L.nbco_compute_force2.argtypes = [vp, C.c_int, vp, i64, vp]
    L.nbco_integrate2.argtypes = [vp, C.c_int, C.c_int, vp, i64, vp, f64, i64]
    L.nbco_mean_rel_err2.argtypes = [vp, vp, vp, i64, C.POINTER(f64), C.POINTER(f64)]
    L.nbco_energy2.argtypes = [vp, vp, i64, vp, C.POINTER(f64)]
    L.nbco_run_host2.argtypes = [vp, C.c_int, C.c_int, vp, vp, i64, vp, f64, i64]
    L.nbco_step_host2.argtypes = [vp, C.c_int, C.c_int, vp, i64, vp, f64, i64]
    L.nbco_fmm2_levels.argtypes = [i64, C.c_int32, f64]
    L.nbco_fmm2_get_info.argtypes = [vp, C.POINTER(Fmm2Info)]
    L.nbco_fmm2_get_tree.argtypes = [vp] + [vp] * 6
    L.nbco_fmm2_get_phase_ms.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(f32), C.c_int]
    L.nbco_init_ga2.argtypes = [vp, i64, vp, vp]
    L.nbco_init_kv2.argtypes = [vp, i64, vp, vp]
    L.nbco_beam_params2.argtypes = [vp, vp, f64, vp]
    L.nbco_state_read2.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(i64)]
    L.nbco_state_write2.argtypes = [C.c_char_p, vp, i64]
    L.nbco_free.restype = None
    return L


lib = _load()


def _check(status):
    if status != 0:
        raise NbcoError(f"nbco status {status}: {lib.nbco_last_error().decode()}")


def _hp(a):
    """host pointer of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def default_config(**kw):
    cfg = Config()
    lib.nbco_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def default_param(n, xi=2e-6, omega0=(1.095, 1.0, 1.0)):
    """parameter block of main3.cu:685-692: {xi/N, 0, 0, w0x^2, w0y^2, w0z^2}"""
    w = np.asarray(omega0, np.float32)
    return np.array([np.float32(xi) / np.float32(n), 0, 0, w[0] * w[0], w[1] * w[1], w[2] * w[2]], np.float32)


def init_ga(n, sigma_x=(0.003, 0.001, 0.01), omega0=(1.095, 1.0, 1.0)):
    """reference initGA state [pos|vel] as a (2, n, 3) float32 array (main3.cu:241-245,662-664)"""
    sx = np.asarray(sigma_x, np.float32)
    su = (np.asarray(omega0, np.float32) * sx).astype(np.float32)
    out = np.empty((2, n, 3), np.float32)
    _check(lib.nbco_init_ga(_hp(out), n, _hp(sx), _hp(su)))
    return out


def init_test_cube(n, sigma_x=(0.003, 0.001, 0.01), omega0=(1.095, 1.0, 1.0)):
    sx = np.asarray(sigma_x, np.float32)
    su = (np.asarray(omega0, np.float32) * sx).astype(np.float32)
    out = np.empty((2, n, 3), np.float32)
    _check(lib.nbco_init_test_cube(_hp(out), n, _hp(sx), _hp(su)))
    return out


TWOPI = 6.283185307179586476925286766559
OMEGA0_2D = (6.22 * TWOPI, 6.21 * TWOPI)   # main.cu:272
EMIT_2D = (0.03e-3, 0.01e-3)               # main.cu:294


def beam_params2(omega0=OMEGA0_2D, emit=EMIT_2D, tune_dep_y=0.8):
    """default beam of main.cu:294-313: dict A (2), omega (2), xi"""
    o = np.zeros(5)
    _check(lib.nbco_beam_params2(_hp(np.asarray(omega0, np.float64)), _hp(np.asarray(emit, np.float64)), tune_dep_y, _hp(o)))
    return dict(A=o[0:2].copy(), omega=o[2:4].copy(), xi=float(o[4]))


def default_param2(n, xi=None, omega0=OMEGA0_2D):
    """parameter block of main.cu:803-808: {xi/N, 0, w0x^2, w0y^2} (float64)"""
    if xi is None:
        xi = beam_params2(omega0)["xi"]
    return np.array([xi / n, 0.0, omega0[0] * omega0[0], omega0[1] * omega0[1]], np.float64)


def init_kv2(n, A=None, omega=None):
    """reference initKV state [pos|vel] as a (2, n, 2) float64 array (main.cu:120-145,779-784)"""
    b = beam_params2()
    A = np.asarray(b["A"] if A is None else A, np.float64)
    omega = np.asarray(b["omega"] if omega is None else omega, np.float64)
    out = np.empty((2, n, 2), np.float64)
    _check(lib.nbco_init_kv2(_hp(out), n, _hp(A), _hp(omega)))
    return out


def init_ga2(n, x=None, u=None):
    """reference 2D initGA state (main.cu:147-170): std.dev. x = A/2, u = omega*A/2 by default (:312-313)"""
    b = beam_params2()
    x = np.asarray(b["A"] / 2 if x is None else x, np.float64)
    u = np.asarray(b["omega"] * b["A"] / 2 if u is None else u, np.float64)
    out = np.empty((2, n, 2), np.float64)
    _check(lib.nbco_init_ga2(_hp(out), n, _hp(x), _hp(u)))
    return out


def fmm2_levels(n, order, dens_inhom=1.0):
    return lib.nbco_fmm2_levels(n, order, dens_inhom)


def shard_range(n, rank, world):
    b, e = C.c_int64(), C.c_int64()
    lib.nbco_shard_range(n, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


class Context:
    """Owns one nbco_ctx.  Device-pointer methods take integers (e.g. torch_tensor.data_ptr())."""

    def __init__(self, **cfg):
        self.cfg = default_config(**cfg)
        self._h = C.c_void_p()
        _check(lib.nbco_create(C.byref(self.cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            lib.nbco_destroy(self._h)
            self._h = None

    __del__ = close

    def set(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        _check(lib.nbco_set_config(self._h, C.byref(self.cfg)))

    @property
    def stream(self):
        return lib.nbco_stream(self._h)

    # ---- device-pointer calls (the reference plugin signature) ----
    def force_direct3(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_force_direct3(self._h, d_pos, d_acc, n, d_param))

    def force_fmm3_kd(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_force_fmm3_kd(self._h, d_pos, d_acc, n, d_param))

    def coulomb_direct3(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_coulomb_direct3(self._h, d_pos, d_acc, n, d_param))

    def coulomb_fmm3_kd(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_coulomb_fmm3_kd(self._h, d_pos, d_acc, n, d_param))

    def add_elastic(self, d_pos, d_acc, n, d_k3=None):
        _check(lib.nbco_add_elastic(self._h, d_pos, d_acc, n, d_k3))

    def step(self, d_b, d_a, ds, n):
        _check(lib.nbco_step(self._h, d_b, d_a, ds, n))

    def compute_force(self, evaluator, d_buf, n, d_param=None):
        _check(lib.nbco_compute_force(self._h, evaluator, d_buf, n, d_param))

    def integrate(self, scheme, evaluator, d_buf, n, d_param, dt, nsteps):
        _check(lib.nbco_integrate(self._h, scheme, evaluator, d_buf, n, d_param, dt, nsteps))

    def integrate_energy(self, scheme, evaluator, d_buf, n, d_param, dt, nsteps):
        """nbco_integrate with the last update fused with the energy reduction: returns (kinetic, elastic) of the final state"""
        out = (C.c_double * 2)()
        _check(lib.nbco_integrate_energy(self._h, scheme, evaluator, d_buf, n, d_param, dt, nsteps, out))
        return out[0], out[1]

    def mean_rel_err(self, d_a, d_ref, n):
        m, x = C.c_double(), C.c_double()
        _check(lib.nbco_mean_rel_err(self._h, d_a, d_ref, n, C.byref(m), C.byref(x)))
        return m.value, x.value

    def energy(self, d_buf, n, d_param=None):
        out = (C.c_double * 3)()
        _check(lib.nbco_energy(self._h, d_buf, n, d_param, out))
        return tuple(out)

    # ---- host-buffer calls ----
    def eval_host(self, evaluator, pos, vel=None, param=None):
        """pos (n,3) float32 [updated in place when the FMM leaves tree order]; returns acc (n,3)"""
        n = pos.shape[0]
        acc = np.empty((n, 3), np.float32)
        _check(lib.nbco_eval_host(self._h, evaluator, _hp(pos), _hp(vel), _hp(acc), n, _hp(param)))
        return acc

    def run_host(self, scheme, evaluator, pos_vel, param, dt, nsteps, want_acc=False):
        """pos_vel (2,n,3) float32, updated in place"""
        n = pos_vel.shape[1]
        acc = np.empty((n, 3), np.float32) if want_acc else None
        _check(lib.nbco_run_host(self._h, scheme, evaluator, _hp(pos_vel), _hp(acc), n, _hp(param), dt, nsteps))
        return acc

    def step_host(self, scheme, evaluator, buf, n, param, dt, nsteps=1):
        """buf: flat float32 [pos|vel|acc] (9n), updated in place"""
        _check(lib.nbco_step_host(self._h, scheme, evaluator, _hp(buf), n, _hp(param), dt, nsteps))

    # ---- FMM introspection ----
    def fmm_info(self):
        info = FmmInfo()
        _check(lib.nbco_fmm_get_info(self._h, C.byref(info)))
        return info

    def fmm_tree(self):
        i = self.fmm_info()
        nn, n = i.nodes, i.n
        t = dict(center=np.empty((nn, 3), np.float32), lbound=np.empty((nn, 3), np.float32),
                 rbound=np.empty((nn, 3), np.float32), mpole=np.empty((nn, i.off_m), np.float32),
                 local=np.empty((nn, i.off_l), np.float32), mult=np.empty(nn, np.int32),
                 index=np.empty(nn, np.int32), splitdim=np.empty(nn, np.int32), perm=np.empty(n, np.int32))
        _check(lib.nbco_fmm_get_tree(self._h, _hp(t["center"]), _hp(t["lbound"]), _hp(t["rbound"]), _hp(t["mpole"]),
                                     _hp(t["local"]), _hp(t["mult"]), _hp(t["index"]), _hp(t["splitdim"]), _hp(t["perm"])))
        t["levels"] = i.levels
        return t

    def fmm_lists(self):
        i = self.fmm_info()
        p2p = np.empty((i.p2p_pairs, 2), np.int32)
        m2l = np.empty((i.m2l_pairs, 2), np.int32)
        _check(lib.nbco_fmm_get_lists(self._h, _hp(p2p), i.p2p_pairs, _hp(m2l), i.m2l_pairs))
        return p2p, m2l

    def track_ids(self, d_ids):
        """device pointer to n int32 ids that follow the particles through the tree rebuilds (0 / None: stop)"""
        _check(lib.nbco_track_ids(self._h, C.c_void_p(d_ids or 0)))

    def fmm_phase_totals(self, reset=False):
        """({phase: total ms}, evaluations, rebuilds) since the last reset"""
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        ev = (C.c_int64 * 2)()
        k = lib.nbco_fmm_phase_totals(self._h, names, ms, 32, ev, 1 if reset else 0)
        return {names[j].decode(): ms[j] for j in range(k)}, ev[0], ev[1]

    # ---- multi-GPU over peer memory ----
    def peer_export(self, n):
        h = np.zeros(192, np.uint8)
        _check(lib.nbco_peer_export(self._h, n, _hp(h)))
        return h

    def peer_attach(self, rank, handles):
        _check(lib.nbco_peer_attach(self._h, rank, _hp(np.ascontiguousarray(handles, np.uint8))))

    def peer_attach_local(self, rank, other):
        _check(lib.nbco_peer_attach_local(self._h, rank, other._h))

    def peer_commit(self):
        _check(lib.nbco_peer_commit(self._h))

    def peer_barrier(self):
        _check(lib.nbco_peer_barrier(self._h))

    def peer_gather(self, d_buf, n):
        _check(lib.nbco_peer_gather(self._h, d_buf, n))

    def peer_detach(self):
        _check(lib.nbco_peer_detach(self._h))

    # ---- 2D fp64 path ----
    def force_direct2(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_force_direct2(self._h, d_pos, d_acc, n, d_param))

    def force_fmm2(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_force_fmm2(self._h, d_pos, d_acc, n, d_param))

    def coulomb_direct2(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_coulomb_direct2(self._h, d_pos, d_acc, n, d_param))

    def coulomb_fmm2(self, d_pos, d_acc, n, d_param=None):
        _check(lib.nbco_coulomb_fmm2(self._h, d_pos, d_acc, n, d_param))

    def add_elastic2(self, d_pos, d_acc, n, d_k2=None):
        _check(lib.nbco_add_elastic2(self._h, d_pos, d_acc, n, d_k2))

    def step2(self, d_b, d_a, ds, n):
        _check(lib.nbco_step2(self._h, d_b, d_a, ds, n))

    def compute_force2(self, evaluator, d_buf, n, d_param=None):
        _check(lib.nbco_compute_force2(self._h, evaluator, d_buf, n, d_param))

    def integrate2(self, scheme, evaluator, d_buf, n, d_param, dt, nsteps):
        _check(lib.nbco_integrate2(self._h, scheme, evaluator, d_buf, n, d_param, dt, nsteps))

    def mean_rel_err2(self, d_a, d_ref, n):
        m, x = C.c_double(), C.c_double()
        _check(lib.nbco_mean_rel_err2(self._h, d_a, d_ref, n, C.byref(m), C.byref(x)))
        return m.value, x.value

    def energy2(self, d_buf, n, d_param=None):
        out = (C.c_double * 3)()
        _check(lib.nbco_energy2(self._h, d_buf, n, d_param, out))
        return tuple(out)

    def run_host2(self, scheme, evaluator, pos_vel, param, dt, nsteps, want_acc=False):
        """pos_vel (2,n,2) float64, updated in place"""
        n = pos_vel.shape[1]
        acc = np.empty((n, 2), np.float64) if want_acc else None
        _check(lib.nbco_run_host2(self._h, scheme, evaluator, _hp(pos_vel), _hp(acc), n, _hp(param), dt, nsteps))
        return acc

    def step_host2(self, scheme, evaluator, buf, n, param, dt, nsteps=1):
        """buf: flat float64 [pos|vel|acc] (6n), updated in place"""
        _check(lib.nbco_step_host2(self._h, scheme, evaluator, _hp(buf), n, _hp(param), dt, nsteps))

    def fmm2_info(self):
        info = Fmm2Info()
        _check(lib.nbco_fmm2_get_info(self._h, C.byref(info)))
        return info

    def fmm2_tree(self):
        """center (nodes,) complex, mpole/local (nodes, order+1) complex, mult, leaf_index (4^L+1), perm"""
        i = self.fmm2_info()
        nn, c, m = i.nodes, i.coeffs, 1 << (2 * i.levels)
        t = dict(center=np.empty(nn, np.complex128), mpole=np.empty((nn, c), np.complex128),
                 local=np.empty((nn, c), np.complex128), mult=np.empty(nn, np.int32),
                 leaf_index=np.empty(m + 1, np.int32), perm=np.empty(i.n, np.int32))
        _check(lib.nbco_fmm2_get_tree(self._h, _hp(t["center"]), _hp(t["mpole"]), _hp(t["local"]), _hp(t["mult"]),
                                      _hp(t["leaf_index"]), _hp(t["perm"])))
        t["levels"] = i.levels
        return t

    def fmm2_phase_ms(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        k = lib.nbco_fmm2_get_phase_ms(self._h, names, ms, 16)
        return {names[j].decode(): ms[j] for j in range(k)}

    def fmm_phase_ms(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        k = lib.nbco_fmm_get_phase_ms(self._h, names, ms, 32)
        return {names[j].decode(): ms[j] for j in range(k)}
