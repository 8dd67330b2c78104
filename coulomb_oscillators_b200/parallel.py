"""Multi-GPU plumbing (one process per GPU, torch.distributed): which rank owns which targets and
how the per-rank results are put back together.  No numerics here: the compute callables are the
C-ABI evaluators (GPU) -- tests substitute the oracle to exercise this logic on CPU with gloo.

Direct sum: every rank holds all sources, computes the accelerations of its own
contiguous target shard [ceil(n r/w), ceil(n (r+1)/w)) -- the kd-tree's own split rule -- and the
shards are all-gathered.  There is no reduction: targets are independent."""
import numpy as np

from ._lib import shard_range


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def all_gather_shards(local_shard, n, group=None):
    """local_shard: torch tensor (count_r, 3) of this rank's targets; returns the full (n, 3) tensor"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n, world)
    if len(set(sizes)) == 1:
        # equal shards (n divisible by the world size): one collective straight into the result, no padding / cat passes
        out = torch.empty((n, 3), dtype=local_shard.dtype, device=local_shard.device)
        dist.all_gather_into_tensor(out, local_shard.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, 3), dtype=local_shard.dtype, device=local_shard.device)
    buf[: local_shard.shape[0]] = local_shard
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def sharded_direct3(compute_shard, pos, n, group=None):
    """compute_shard(rank, world) -> (n, 3) tensor whose rows [begin, end) of this rank are valid
    (what nbco_force_direct3 writes with cfg.rank/cfg.world set); returns the assembled (n, 3)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(n, rank, world)
    acc = compute_shard(rank, world)
    return all_gather_shards(acc[b:e].contiguous(), n, group)


def next_eval_rebuilds(ctx):
    """True when the next FMM evaluation of this context rebuilds the kd-tree (counter % tree_steps == 0,
    fmm_cart3_kdtree.cuh:1619); a context that has not evaluated yet always builds."""
    from ._lib import NbcoError
    try:
        return ctx.fmm_info().counter % ctx.cfg.tree_steps == 0
    except NbcoError:
        return True


def fmm_leapfrog_sharded(ctx, buf, n, d_param, dt, nsteps, group=None, gather_final=True):
    """Leapfrog steps of coulombOscillatorFMMKD3 over `world` GPUs (one process per GPU).

    Partition: rank r of 2^g owns the subtree of kd node (g, r), i.e. the
    contiguous tree-order range [ceil(n r/w), ceil(n (r+1)/w)) of particles.  Every rank keeps the
    full position array: the tree (build, P2M/M2M, traversal) is replicated, while P2P, M2L,
    L2L/L2P and the kick/drift run on the rank's own range only (ctx.cfg.rank/world).  Exchange per
    step: one all-gather of the drifted positions; on tree-rebuild steps also of the velocities,
    because the rebuild permutes the whole state.

    buf: torch float32 tensor [pos | vel | acc] (9n) on this rank's GPU, identical on all ranks at
    entry and with acc already computed (main3.cu:835-839); ctx: Context(rank=r, world=w, unsort=0)."""
    import torch
    import torch.distributed as dist
    from ._lib import LEAPFROG, EVAL_COULOMB_FMM3_KD  # noqa: F401
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    assert ctx.cfg.rank == rank and ctx.cfg.world == world and ctx.cfg.unsort == 0
    b, e = shard_range(n, rank, world)
    cnt = e - b
    pos, vel, acc = buf[:3 * n], buf[3 * n:6 * n], buf[6 * n:]
    sizes = shard_sizes(n, world)
    equal = len(set(sizes)) == 1
    dtf = float(np.float32(dt))
    half = float(np.float32(np.longdouble(dtf) * np.longdouble(0.5)))

    def gather(full, lo, hi):
        if world == 1:
            return
        local = full[3 * lo:3 * hi].clone()
        if equal:
            dist.all_gather_into_tensor(full, local, group=group)
        else:
            pad = 3 * max(sizes)
            tmp = torch.zeros(pad, dtype=full.dtype, device=full.device)
            tmp[:local.numel()] = local
            out = [torch.empty_like(tmp) for _ in range(world)]
            dist.all_gather(out, tmp, group=group)
            full.copy_(torch.cat([o[:3 * s] for o, s in zip(out, sizes)]))
        torch.cuda.current_stream().synchronize()

    p0 = buf.data_ptr()
    for _ in range(nsteps):
        ctx.step(p0 + 4 * (3 * n + 3 * b), p0 + 4 * (6 * n + 3 * b), half, cnt)   # v += a dt/2   (own range)
        ctx.step(p0 + 4 * (3 * b), p0 + 4 * (3 * n + 3 * b), dtf, cnt)            # x += v dt
        gather(pos, b, e)
        if next_eval_rebuilds(ctx):
            gather(vel, b, e)                                                     # the rebuild permutes everything
        ctx.coulomb_fmm3_kd(p0, p0 + 4 * 6 * n, n, d_param)                       # a = f(x)     (own range written)
        ctx.step(p0 + 4 * (3 * n + 3 * b), p0 + 4 * (6 * n + 3 * b), half, cnt)   # v += a dt/2
    if gather_final:   # leave the full state on every rank (callers that only read their own range skip this)
        gather(vel, b, e)
        gather(acc, b, e)


def peer_setup(ctx, n, group=None):
    """Map every rank's published buffers into every other rank (csrc/peer.cu): the 192-byte CUDA IPC handle
    blocks travel through torch.distributed (an all-gather of uint8 tensors over NCCL); afterwards the C library
    talks to its peers directly over NVLink."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert ctx.cfg.rank == rank and ctx.cfg.world == world and ctx.cfg.unsort == 0
    mine = torch.from_numpy(np.ascontiguousarray(ctx.peer_export(n), np.uint8))
    if "nccl" in str(dist.get_backend(group)):
        mine = mine.cuda()
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    for q, h in enumerate(every):
        if q != rank:
            ctx.peer_attach(q, h.cpu().numpy())
    ctx.peer_commit()
    dist.barrier(group=group)
    ctx.peer_barrier()


def fmm_leapfrog_peer(ctx, buf, n, d_param, dt, nsteps, gather_final=True):
    """Leapfrog steps of coulombOscillatorFMMKD3 over the ranks of a peer_setup() context: ONE C call per
    rank (nbco_integrate), all exchanges inside the kernels / the flag barrier of csrc/peer.cu."""
    from ._lib import LEAPFROG, EVAL_COULOMB_FMM3_KD
    ctx.integrate(LEAPFROG, EVAL_COULOMB_FMM3_KD, buf.data_ptr(), n, d_param, dt, nsteps)
    if gather_final:
        ctx.peer_gather(buf.data_ptr(), n)


# ---- 2D fp64 FMM over several GPUs ----
def _ld(x):
    return np.longdouble(x)


def scheme_schedule(scheme, dt):
    """[(op, coefficient)] of one step of integrator.cuh:32-167: op 'K' (v += c a), 'D' (x += c v), 'F' (a = f(x)).
    Coefficients are formed in long double and cast to the scalar type at the call, like the reference."""
    from ._lib import EULER, LEAPFROG, FORESTRUTH, PEFRL
    dt = _ld(dt)
    if scheme == EULER:
        return [("K", dt), ("D", dt), ("F", None)]
    if scheme == LEAPFROG:
        return [("K", dt * _ld(0.5)), ("D", dt), ("F", None), ("K", dt * _ld(0.5))]
    if scheme == FORESTRUTH:
        th = _ld("1.3512071919596576340476878089715")
        return [("D", dt * th / 2), ("F", None), ("K", dt * th), ("D", dt * (1 - th) / 2), ("F", None), ("K", dt * (1 - 2 * th)),
                ("D", dt * (1 - th) / 2), ("F", None), ("K", dt * th), ("D", dt * th / 2)]
    if scheme == PEFRL:
        xi, la, ch = _ld("0.1786178958448091E+00"), _ld("-0.2123418310626054E+00"), _ld("-0.6626458266981849E-01")
        return [("D", dt * xi), ("F", None), ("K", dt * (1 - 2 * la) / 2), ("D", dt * ch), ("F", None), ("K", dt * la),
                ("D", dt * (1 - 2 * (ch + xi))), ("F", None), ("K", dt * la), ("D", dt * ch), ("F", None),
                ("K", dt * (1 - 2 * la) / 2), ("D", dt * xi)]
    raise ValueError(scheme)


def fmm2_integrate_sharded(ctx, scheme, buf, n, d_param, dt, nsteps, group=None):
    """nsteps steps of a symplectic scheme over coulombOscillatorFMM (2D fp64, main.cu) on `world` GPUs, one process
    per GPU.  Every rank holds the full state and integrates it (kick / drift are 48 B per particle); the force evaluation
    -- nbco_coulomb_fmm2 with cfg.rank / cfg.world -- sorts and summarises all particles on every rank (the tree is
    replicated) and runs the dominant near-field + L2P kernel on the rank's own range of the cell-sorted particles; the
    ranges of the accelerations are exchanged with ONE NCCL all-gather per evaluation (the real exchange step of this
    path).  buf: torch float64 [pos | vel | acc] (6 n), identical on all ranks at entry."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    assert ctx.cfg.rank == rank and ctx.cfg.world == world
    p0 = buf.data_ptr()
    pos, vel, acc = p0, p0 + 8 * 2 * n, p0 + 8 * 4 * n
    sizes = shard_sizes(n, world)
    b, e = shard_range(n, rank, world)
    acc_t = buf[4 * n:]
    equal = len(set(sizes)) == 1

    def force():
        ctx.coulomb_fmm2(pos, acc, n, d_param)          # writes acc[b:e] (cell order; pos / vel permuted identically on all ranks)
        if world == 1:
            return
        if equal:
            dist.all_gather_into_tensor(acc_t, acc_t[2 * b:2 * e].clone(), group=group)
        else:
            pad = 2 * max(sizes)
            tmp = torch.zeros(pad, dtype=acc_t.dtype, device=acc_t.device)
            tmp[:2 * (e - b)] = acc_t[2 * b:2 * e]
            out = [torch.empty_like(tmp) for _ in range(world)]
            dist.all_gather(out, tmp, group=group)
            acc_t.copy_(torch.cat([o[:2 * s] for o, s in zip(out, sizes)]))
        if acc_t.is_cuda:
            torch.cuda.current_stream().synchronize()

    sched = scheme_schedule(scheme, dt)
    for _ in range(nsteps):
        for op, c in sched:
            if op == "K":
                ctx.step2(vel, acc, float(c), n)
            elif op == "D":
                ctx.step2(pos, vel, float(c), n)
            else:
                force()
