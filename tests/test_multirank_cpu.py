"""world_size-2 gloo test of the multi-GPU host logic (coulomb_oscillators_b200/parallel.py):
target sharding + all-gather reassemble exactly the single-rank result.  The per-rank compute is
the oracle here (CPU); on the GPU box the same plumbing wraps nbco_force_direct3."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import coulomb_oscillators_b200 as nb
    from coulomb_oscillators_b200.parallel import sharded_direct3
    from refs import Oracle
    st = nb.init_ga(n)
    par = nb.default_param(n)
    full = Oracle().direct3(st[0].copy(), par)

    def compute_shard(r, w):
        b, e = nb.shard_range(n, r, w)
        out = np.zeros((n, 3), np.float32)
        out[b:e] = full[b:e]          # what a rank-r context writes: only its own targets
        out[:b] = np.nan
        out[e:] = np.nan
        return torch.from_numpy(out)

    got = sharded_direct3(compute_shard, st[0], n).numpy()
    ok = bool(np.array_equal(got, full))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_sharded_direct_sum_gloo_world2():
    world, n = 2, 1001          # ragged: shards of 501 and 500 targets
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


class _FakePeerCtx:
    """records what peer_setup does to a context (the real one needs a GPU: tests/test_peer_gpu.py, tools/peer_check.py)"""

    class _Cfg:
        unsort = 0

    def __init__(self, rank, world):
        self.cfg = self._Cfg()
        self.cfg.rank, self.cfg.world = rank, world
        self.calls = []

    def peer_export(self, n):
        self.calls.append(("export", n))
        return np.full(192, 10 + self.cfg.rank, np.uint8)       # a recognisable "handle block"

    def peer_attach(self, q, handles):
        self.calls.append(("attach", q, int(handles[0]), len(handles)))

    def peer_commit(self):
        self.calls.append(("commit",))

    def peer_barrier(self):
        self.calls.append(("barrier",))


def _peer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coulomb_oscillators_b200.parallel import peer_setup
    ctx = _FakePeerCtx(rank, world)
    peer_setup(ctx, 12345)
    dist.destroy_process_group()
    q.put((rank, ctx.calls))


def test_peer_setup_exchanges_every_handle_block_gloo_world2():
    """every rank exports once, attaches the 192-byte block of every OTHER rank exactly once, commits, then meets the
    others at the peer barrier"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        other = 1 - r
        assert res[r] == [("export", 12345), ("attach", other, 10 + other, 192), ("commit",), ("barrier",)]


class _OracleCtx2:
    """stands in for a GPU context in fmm2_integrate_sharded: the pointer-based calls are served by the 2D oracle on
    host memory; like nbco_coulomb_fmm2 with cfg.rank / cfg.world it leaves only the rank's own range of acc valid"""

    class _Cfg:
        pass

    def __init__(self, rank, world, order):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from refs2d import Oracle2
        self.cfg = self._Cfg()
        self.cfg.rank, self.cfg.world = rank, world
        self.orc = Oracle2(order=order)

    @staticmethod
    def _arr(ptr, k):
        import ctypes as C
        return np.ctypeslib.as_array((C.c_double * k).from_address(ptr))

    def step2(self, d_b, d_a, ds, n):
        b, a = self._arr(d_b, 2 * n), self._arr(d_a, 2 * n)
        b += a * ds

    def coulomb_fmm2(self, d_pos, d_acc, n, d_param):
        import coulomb_oscillators_b200 as nb
        buf = self._arr(d_pos, 6 * n)                      # [pos | vel | acc] starts at pos
        tmp = buf.copy()
        self.orc.eval(3, tmp, n, self._arr(d_param, 4))     # coulombOscillatorFMM: sorts pos / vel by cell, writes acc
        b, e = nb.shard_range(n, self.cfg.rank, self.cfg.world)
        buf[:4 * n] = tmp[:4 * n]
        buf[4 * n:] = np.nan
        buf[4 * n + 2 * b:4 * n + 2 * e] = tmp[4 * n + 2 * b:4 * n + 2 * e]


def _fmm2_worker(rank, world, port, n, order, scheme, steps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import coulomb_oscillators_b200 as nb
    from coulomb_oscillators_b200.parallel import fmm2_integrate_sharded
    from refs2d import Oracle2
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    want = np.concatenate([st[0], st[1], np.zeros((n, 2))]).ravel().copy()
    o = Oracle2(order=order)
    o.eval(3, want, n, par)
    o.integrate(scheme, 3, want, n, par, 5e-4, steps)
    buf = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))]).ravel().copy())
    ctx = _OracleCtx2(rank, world, order)
    par_t = torch.from_numpy(par.copy())
    ctx.coulomb_fmm2(buf.data_ptr(), buf.data_ptr() + 8 * 4 * n, n, par_t.data_ptr())   # compute_force before the loop, own range only ...
    sizes = [nb.shard_range(n, r, world)[1] - nb.shard_range(n, r, world)[0] for r in range(world)]
    b, e = nb.shard_range(n, rank, world)
    pad = 2 * max(sizes)
    tmp = torch.zeros(pad, dtype=torch.float64)
    tmp[:2 * (e - b)] = buf[4 * n + 2 * b:4 * n + 2 * e]
    out = [torch.empty_like(tmp) for _ in range(world)]
    dist.all_gather(out, tmp)                                                              # ... then gathered like inside the loop
    buf[4 * n:] = torch.cat([x[:2 * s] for x, s in zip(out, sizes)])
    fmm2_integrate_sharded(ctx, scheme, buf, n, par_t.data_ptr(), 5e-4, steps)
    got = buf.numpy()
    ok = bool(np.isfinite(got).all() and np.abs(got - want).max() <= 1e-12 * np.abs(want).max())
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_fmm2_sharded_integration_gloo_world2():
    """2D multi-GPU host logic (parallel.fmm2_integrate_sharded): replicated tree, per-rank ranges of the accelerations
    all-gathered after every evaluation, PEFRL schedule with long-double coefficients -- equal to the oracle's own
    PEFRL integration to 1e-12 (ragged ranges: n odd)"""
    world, n, order, scheme, steps = 2, 3001, 4, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fmm2_worker, args=(r, world, port, n, order, scheme, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
