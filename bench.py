#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): 3D kd-tree FMM particle-steps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n N] [--order P]

A "step" = one leapfrog step of coulombOscillatorFMMKD3 (kick, drift, FMM + elastic force, kick)
over all N particles, tree rebuilt every tree_steps = 8 force evaluations like the reference GPU
path.  One JSON line is printed by rank 0 (see the task contract for the keys).  Inputs are the
reference's own Gaussian initial conditions (initGA, fixed seed) generated on the host.

  value      particle-steps/s with the state resident in HBM (nbco_integrate on device pointers)
  e2e        the same metric through the host-buffer C-ABI call nbco_step_host: every step copies
             [pos|vel|acc] host->device from pinned memory, runs one step, copies it back
  roofline   dominant phase of the step against measured HBM bandwidth (algorithmic bytes from
             SURVEY.md section 8(d), evaluated with the actual list sizes)
  cpu_baseline  the UNMODIFIED reference CPU path (oracle/_ref, kind "reference") or, when that
             library is absent, our C restatement (kind "port"), timed on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[0]))
                    for nme, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# single-kernel phases (their CUDA-event time is one kernel launch).  Orders 1..6 run the by-target flow (round 2): the
# "p2p" phase is the bucketing of the lists by target (6 small kernels), "m2l" is empty (the M2L sums are gathered inside
# the "l2l" levels) and "l2p" = L2P + near field in one kernel.
SINGLE_KERNEL_PHASES = {"kd_bottom": "kd_bottom_kernel", "l2p": "l2p_near_kernel"}
SINGLE_KERNEL_PHASES_PAIRS = {"kd_bottom": "kd_bottom_kernel", "p2p": "p2p_kernel", "m2l": "m2l_kernel", "l2p": "l2p_kernel"}
# DRAM bytes per launch from `ncu --set full` captures of this round (profiles/r02_notes.md); None = not captured
NCU_TRAFFIC = {"kd_bottom": {"n": 1 << 24, "bytes": 272.42368e6 + 367.023872e6}}


def kd_top_levels(n, L, cap=8192):
    """levels partitioned globally before the shared-memory kernel takes over (kdtree.cu: kd_reserve)"""
    lt = 0
    while ((n - 1) >> lt) + 1 > cap:
        lt += 1
    return min(lt, L - 1)


def by_target_flow(order):
    """cfg.reproducible (lists bucketed by target, no atomics) is opt-in and slower; the bench runs the default pair flow"""
    return False


def fmm_bytes_per_eval(n, order, L, p2p_pairs, m2l_pairs, survey_model=False):
    """algorithmic bytes of one FMM evaluation per phase.  The survey's
    rebuild figure (16 L + 52) N is apportioned by level count between the global levels (kd_top), the
    shared-memory levels (kd_bottom) and the final permutation of pos/vel (permute, 40 N)."""
    Nl, Nn = 1 << L, (1 << (L + 1)) - 1
    SM = 4 * order * (order + 1) * (order + 2) // 6
    SL = 4 * (order + 1) ** 2
    lt = kd_top_levels(n, L)
    b = {
        "kd_top": (16 * lt + 12) * n,
        "kd_bottom": 16 * (L - lt) * n,
        "permute": 40 * n,
        "p2m_m2m": 12 * n + Nl * (12 + SM) + (2 * Nn - Nl) * (16 + SM),
        "traverse": 40 * Nn + 8 * (m2l_pairs + p2p_pairs),
        "m2l": 8 * m2l_pairs + Nn * (12 + SM) + Nn * SL,
        "p2p": 8 * p2p_pairs + 8 * Nl + 24 * n,
        "l2l": Nn * (12 + 2 * SL),
        "l2p": Nl * (12 + SL) + 36 * n,
    }
    if by_target_flow(order) and not survey_model:
        # same algorithmic bytes, attributed to the phases that now do the work: the levels gather the M2L sums (m2l + l2l),
        # the leaf kernel gathers the near field (l2p + p2p), "p2p" = bucketing the lists (read twice, 4 B per directed entry, row offsets)
        pairs = m2l_pairs + p2p_pairs
        b["l2l"] = b["l2l"] + b["m2l"]
        b["l2p"] = b["l2p"] + b["p2p"]
        b["p2p"] = 16 * pairs + 8 * pairs + 8 * (Nn + Nl)
        b["m2l"] = 0
    return b


REBUILD_PHASES = ("kd_top", "kd_bottom", "permute")


def run_ours(args):
    import torch
    import torch.distributed as dist
    import coulomb_oscillators_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n, order = args.n, args.order
    peaks, peak_kind = measured_peaks()

    # Multi-GPU (N > 1): ONE system of n particles, strong scaling.  Rank r of 2^g owns the subtree of kd node
    # (g, r): it builds, summarises, traverses and steps its own tree-order range; remote nodes and leaves are
    # read from their owners over NVLink peer memory inside the kernels, the ranks meet at a flag barrier in
    # peer memory (csrc/peer.cu).  NCCL only carries the IPC handles and the timing reduce.
    from coulomb_oscillators_b200.parallel import fmm_leapfrog_peer, fmm_leapfrog_sharded, peer_setup
    state = nb.init_ga(n)
    par = nb.default_param(n)
    ctx = nb.Context(device=local, order=order, unsort=0, tree_steps=8, m2l_first=args.m2l_first, rank=rank, world=world)
    buf = torch.empty(9 * n, dtype=torch.float32, device="cuda")
    buf[:6 * n] = torch.from_numpy(state.reshape(-1)).cuda()
    dpar = torch.from_numpy(par).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream)
    ev = nb.EVAL_COULOMB_FMM3_KD
    dt = 5e-4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peer_ok = world > 1
    if world > 1:
        # peer memory needs CUDA IPC + P2P between the ranks' devices; if any rank cannot map the others, ALL ranks
        # fall back to the replicated-tree mode (all-gather of the positions per step over NCCL)
        try:
            peer_setup(ctx, n)
            flag = 1
        except Exception as e:  # noqa: BLE001
            print(f"rank {rank}: peer setup failed ({e}); using the replicated-tree mode", file=sys.stderr, flush=True)
            flag = 0
        t = torch.tensor([flag], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        peer_ok = bool(t.item())
        if not peer_ok:
            ctx = nb.Context(device=local, order=order, unsort=0, tree_steps=8, m2l_first=args.m2l_first, rank=rank, world=world)
            stream = torch.cuda.ExternalStream(ctx.stream)

    def run_steps(k):
        if world == 1:
            ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, dpar.data_ptr(), dt, k)
        elif peer_ok:
            fmm_leapfrog_peer(ctx, buf, n, dpar.data_ptr(), dt, k, gather_final=False)
        else:
            fmm_leapfrog_sharded(ctx, buf, n, dpar.data_ptr(), dt, k)

    ctx.compute_force(ev, buf.data_ptr(), n, dpar.data_ptr())       # main3.cu:835-839
    run_steps(args.warmup)
    ctx.fmm_phase_totals(reset=True)
    l0 = ctx.fmm_info().kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    e1.synchronize()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    info = ctx.fmm_info()
    launches = info.kernel_launches - l0
    totals, evals, rebuilds = ctx.fmm_phase_totals(reset=True)
    assert np.isfinite(buf[:6 * n].sum().item()), "state diverged"
    steps_done = args.warmup + args.steps

    # ---- parity on the TIMED state (not timed): FMM forces against the direct sum on sampled targets, and for
    # N > 1 ranks the whole state against a single-rank run of the same steps ----
    if world > 1 and peer_ok:
        ctx.peer_gather(buf.data_ptr(), n)          # every rank: full [pos | vel | acc] of the sharded run
    parity = parity_block(nb, torch, ctx if world == 1 else None, buf, n, order, dpar, local, args.m2l_first,
                          state if (world > 1 and rank == 0) else None, steps_done, dt) if (rank == 0 or world == 1) else None
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region, every step ----
    # One GPU: nbco_run_host, the reference's resume-from-snapshot flow (main3.cu:629-658,835-858): the host holds the
    # state [pos | vel] (the reference's state file), every step uploads it from pinned memory, recomputes a = f(x),
    # takes one leapfrog step and reads [pos | vel] back: 24 N bytes each way and TWO force evaluations per step.
    # (Round 1 moved [pos | vel | acc], 36 N each way, with one evaluation: PCIe-bound at 22.8 ms per step.)
    hbuf = torch.empty(9 * n, dtype=torch.float32).pin_memory()
    hbuf.copy_(buf.cpu())
    hnp = hbuf.numpy()
    hpv = hnp[:6 * n].reshape(2, n, 3)
    ctx_e = nb.Context(device=local, order=order, unsort=0, tree_steps=8, m2l_first=args.m2l_first) if world == 1 else ctx
    e2e_steps = max(3, min(args.steps, 8))
    lo, hi = nb.shard_range(n, rank, world)

    def e2e_step():
        if world == 1:
            ctx_e.run_host(nb.LEAPFROG, ev, hpv, par, dt, 1)
        else:
            # every rank owns a tree-order range: it uploads its range of pos / vel / acc, the ranks step together
            # (remote data moves between the GPUs inside the evaluator), every rank reads its range back
            for k in (0, 1, 2):
                buf[3 * n * k + 3 * lo:3 * n * k + 3 * hi].copy_(hbuf[3 * n * k + 3 * lo:3 * n * k + 3 * hi], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if peer_ok:
                fmm_leapfrog_peer(ctx_e, buf, n, dpar.data_ptr(), dt, 1, gather_final=False)
            else:
                fmm_leapfrog_sharded(ctx_e, buf, n, dpar.data_ptr(), dt, 1, gather_final=False)
            for k in (0, 1, 2):
                hbuf[3 * n * k + 3 * lo:3 * n * k + 3 * hi].copy_(buf[3 * n * k + 3 * lo:3 * n * k + 3 * hi], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_step()   # warm-up (allocations, first rebuild)
    barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    if world > 1:
        t = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e_value = n * e2e_steps / te

    # secondary metric (BASELINE config 3): O(N^2) direct sum, targets sharded over the ranks
    direct = bench_direct(nb, torch, dist, local, peaks, rank, world) if args.direct else None

    # secondary metric (BASELINE config 4): 2D fp64 FMM under PEFRL, one GPU
    fmm2d = None
    if args.fmm2d:
        fmm2d = bench_fmm2d(nb, torch, local) if world == 1 else bench_fmm2d_sharded(nb, torch, dist, local, rank, world)

    if world > 1 and peer_ok:
        ctx.peer_detach()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n * args.steps / (ms * 1e-3)   # one system of n particles, whatever the number of GPUs
    # ---- per-phase times (CUDA events on the context stream, summed over the timed region) ----
    bytes_eval = fmm_bytes_per_eval(n, order, info.levels, info.p2p_pairs * world, info.m2l_pairs * world)
    if world > 1:   # per-rank share of the algorithmic bytes (rank 0's phase times are reported)
        bytes_eval = {k: v / world for k, v in bytes_eval.items()}
    phases = {}
    total_ms = max(sum(totals.values()), 1e-9)
    for k, tot in totals.items():
        calls = rebuilds if k in REBUILD_PHASES else evals
        if calls == 0 or bytes_eval.get(k, 0) == 0:
            continue
        avg_ms = tot / calls
        phases[k] = {"avg_ms": round(avg_ms, 4), "launches": int(calls), "share": round(tot / total_ms, 4),
                     "GBps": round(bytes_eval[k] / (avg_ms * 1e-3) / 1e9, 1) if avg_ms > 0 else None}
    # ---- roofline of the dominant KERNEL: the single-kernel phase with the largest total time ----
    single = SINGLE_KERNEL_PHASES if by_target_flow(order) else SINGLE_KERNEL_PHASES_PAIRS
    dom = max(single, key=lambda k: totals.get(k, 0.0))
    dcalls = rebuilds if dom in REBUILD_PHASES else evals
    dom_ms = totals[dom] / max(dcalls, 1)
    achieved = bytes_eval[dom] / (dom_ms * 1e-3) / 1e9
    tr = NCU_TRAFFIC.get(dom)
    # whole step against HBM: SURVEY.md section 8(d)'s model unchanged (the list bucketing of the by-target flow is overhead, not work)
    sv = fmm_bytes_per_eval(n, order, info.levels, info.p2p_pairs * world, info.m2l_pairs * world, survey_model=True)
    step_bytes = sum(v for k, v in sv.items() if k not in REBUILD_PHASES) + sum(sv[k] for k in REBUILD_PHASES) / 8 + 60 * n
    roofline = {"bound": "hbm", "kernel": single[dom], "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                "traffic": tr["bytes"] if tr and tr["n"] == n else None, "peak_kind": peak_kind,
                "avg_launch_ms": round(dom_ms, 4), "launches": int(dcalls),
                "algorithmic_bytes_per_launch": int(bytes_eval[dom]),  # per rank
                "share_of_step": round(totals[dom] / total_ms, 4),
                "step_bytes_per_particle": round(step_bytes / n, 1),
                "step_hbm_frac": round(step_bytes * (args.steps / (ms * 1e-3)) / 1e9 / (world * peaks["hbm_gbs"]), 4)}

    cpu = cpu_baseline(n, order, args.m2l_first, bounded=True) if world == 1 else None
    ref_gpu = reference_gpu_baseline(local) if (world == 1 and args.ref_gpu) else None
    config1 = config1_direct_leg(nb, torch, local) if world == 1 else None
    out = {
        "metric": "3D FMM particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"3D kd-tree FMM leapfrog, N={n}, p={order}, r=1, tree_steps=8, reference initGA ICs"
                               + (f", sharded over {world} GPUs (rank r owns the subtree of kd node (log2 N, r); remote nodes / "
                                  f"leaves read over NVLink peer memory inside the kernels, flag barriers in peer memory)" if world > 1 and peer_ok else "")
                               + (f", sharded over {world} GPUs (replicated tree, all-gather of positions per step; peer setup failed)" if world > 1 and not peer_ok else ""),
                   "n": n, "order": order, "levels": int(info.levels), "p2p_pairs": int(info.p2p_pairs),
                   "m2l_pairs": int(info.m2l_pairs), "m2l_first": args.m2l_first,
                   "l2_hygiene": "inputs larger than L2 (state 36 B x N + tree slab)" if 36 * n > 126e6 else "working set may fit L2",
                   "e2e_copies": "every step: H2D [pos|vel] from pinned memory, a = f(x) recomputed on the device, one step, D2H [pos|vel]"},
        "roofline": roofline, "phases": phases, "cpu_baseline": cpu, "parity": parity,
        "rebuilds_per_step": round(rebuilds / max(evals, 1), 4), "reference_gpu_baseline": ref_gpu, "config1_direct_n8192": config1,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s",
                "h2d_bytes_per_step": (24 if world == 1 else 36) * n, "d2h_bytes_per_step": (24 if world == 1 else 36) * n, "steps": e2e_steps,
                "evaluations_per_step": 2 if world == 1 else 1,
                "api": "nbco_run_host([pos|vel] in pinned host memory): upload, a = f(x), one leapfrog step, read back" if world == 1
                       else "per rank: own range of [pos|vel|acc] host -> device, one peer-mode step, device -> host"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if direct is not None:
        out["direct_sum"] = direct
    if fmm2d is not None:
        out["fmm2d"] = fmm2d
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_block(nb, torch, ctx, buf, n, order, dpar, local, m2l_first, ic_state, steps_done, dt, shards=8, per_shard=512):
    """Accuracy evidence on the state the timed steps produced (SURVEY.md section 8(d) "accuracy reporting"):
      * Coulomb-only FMM forces at the final positions against nbco_force_direct3 on shards x per_shard sampled
        targets (sharded direct-sum calls: cfg.rank/world select the target rows), metric = mean rel_diff1
        (reductions.cuh:37-42; test_accuracy, main3.cu:139-182).  The reference's own class at p = 3, r = 1 on these
        ICs is 0.06-0.11; bit-exact tree/list parity at this N is the job of
        tests/test_fmm_gpu.py::test_fmm_headline_size_matches_live_reference.
      * N > 1 ranks (ic_state given): positions / velocities of the sharded run against a single-rank run of the
        same number of steps from the same initial conditions (max |diff| / max |value|)."""
    out = {}
    pos_ptr = buf.data_ptr()
    ev_ctx = ctx if ctx is not None else nb.Context(device=local, order=order, unsort=0, tree_steps=8, m2l_first=m2l_first)
    a_fmm = torch.empty(3 * n, dtype=torch.float32, device="cuda")
    if ctx is None:
        work = buf.clone()                      # a fresh context sorts its input: keep the gathered state intact
        pos_ptr = work.data_ptr()
    ev_ctx.force_fmm3_kd(pos_ptr, a_fmm.data_ptr(), n, dpar.data_ptr())
    a_dir = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
    w = max(n // per_shard, 1)
    dctx = nb.Context(device=local)
    rows = []
    for k in range(shards):
        r = min(w - 1, (k * w) // shards + (w // (2 * shards)))
        dctx.set(rank=r, world=w)
        dctx.force_direct3(pos_ptr, a_dir.data_ptr(), n, dpar.data_ptr())
        b, e = nb.shard_range(n, r, w)
        rows.append((b, e))
    torch.cuda.synchronize()
    idx = torch.cat([torch.arange(b, e, device="cuda") for b, e in rows])
    f = a_fmm.view(n, 3)[idx].double()
    d = a_dir.view(n, 3)[idx].double()
    rel = ((f - d).pow(2).sum(1) / (d.pow(2).sum(1) + 1e-18)).sqrt()
    out["fmm_vs_direct"] = {"targets": int(idx.numel()), "mean_rel_diff1": float(rel.mean().item()), "max_rel_diff1": float(rel.max().item()),
                            "reference_class_p3_r1": [0.02, 0.2], "what": "Coulomb-only forces at the final positions of the timed run"}
    if order == 3:
        out["fmm_vs_direct"]["ok"] = bool(0.02 < out["fmm_vs_direct"]["mean_rel_diff1"] < 0.2)
    if ic_state is not None:
        c1 = nb.Context(device=local, order=order, unsort=0, tree_steps=8, m2l_first=m2l_first)
        b1 = torch.empty(9 * n, dtype=torch.float32, device="cuda")
        b1[:6 * n] = torch.from_numpy(ic_state.reshape(-1)).cuda()
        c1.compute_force(nb.EVAL_COULOMB_FMM3_KD, b1.data_ptr(), n, dpar.data_ptr())
        c1.integrate(nb.LEAPFROG, nb.EVAL_COULOMB_FMM3_KD, b1.data_ptr(), n, dpar.data_ptr(), dt, steps_done)
        g = buf.view(3, n, 3)
        s1 = b1.view(3, n, 3)
        dev = [float(((g[k] - s1[k]).abs().max() / s1[k].abs().max()).item()) for k in range(3)]
        out["vs_single_rank"] = {"steps": int(steps_done), "pos": dev[0], "vel": dev[1], "acc": dev[2],
                                 "ok": bool(dev[0] < 1e-5 and dev[1] < 1e-4),
                                 "what": "max |sharded - single rank| / max |single rank| after the same steps (same particle order: the kd build is bit-exact whoever runs it)"}
    return out


def bench_direct(nb, torch, dist, local, peaks, rank, world, n=1 << 20):
    """secondary metric of BASELINE.json: O(N^2) direct sum, FP32-FMA roofline (18 flop / interaction).
    Multi-GPU: every rank holds all sources and computes its own target shard (nbco_shard_range);
    the shards are all-gathered (coulomb_oscillators_b200/parallel.py); time = max over ranks."""
    from coulomb_oscillators_b200.parallel import all_gather_shards
    state = nb.init_ga(n)
    pos = torch.from_numpy(state[0].copy()).cuda()
    acc = torch.empty_like(pos)
    par = torch.from_numpy(nb.default_param(n)).cuda()
    ctx = nb.Context(device=local, rank=rank, world=world)
    st = torch.cuda.ExternalStream(ctx.stream)
    lo, hi = nb.shard_range(n, rank, world)

    def once():
        ctx.force_direct3(pos.data_ptr(), acc.data_ptr(), n, par.data_ptr())
        if world > 1:
            full = all_gather_shards(acc[lo:hi].contiguous(), n)
            torch.cuda.current_stream().synchronize()
            return full
        return acc

    once()
    best = 1e30
    for _ in range(2):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        once()
        e1.record(st)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = min(best, ms)
    inter = n * n / (best * 1e-3)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    peak = world * 2 * sms * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    return {"n": n, "n_gpus": world, "ms": best, "Ginteractions_per_s": inter / 1e9,
            "roofline": {"bound": "fp32_fma", "achieved": inter * 18 / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": inter * 18 / 1e12 / peak, "flop_per_interaction": 18}}


FP64_DFMA_PEAK = 17.1e12   # DFMA/s, measured on this pool's B200 with tools/dfma_peak.cu (profiles/r01_notes.md)


def bench_fmm2d(nb, torch, local, n=1 << 22, order=5, steps=4):
    """secondary metric (BASELINE config 4): 2D fp64 FMM (uniform grid, p = 5) under PEFRL, N = 4M, KV beam of
    main.cu.  One PEFRL step = 4 force evaluations + 5 streaming passes.  Roofline of the dominant kernel
    (near field + L2P): FP64 pipe, 9.75 DP instructions per pair interaction (4 pairs share one reciprocal)."""
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    ctx = nb.Context(device=local, order=order)
    buf = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))])).cuda()
    dpar = torch.from_numpy(par).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream)
    ev = nb.EVAL_COULOMB_FMM2
    ctx.compute_force2(ev, buf.data_ptr(), n, dpar.data_ptr())
    ctx.integrate2(nb.PEFRL, ev, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, 1)
    l0 = ctx.fmm2_info().kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    ctx.integrate2(nb.PEFRL, ev, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, steps)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    info = ctx.fmm2_info()
    ph = ctx.fmm2_phase_ms()          # phases of the last evaluation
    T = ctx.fmm2_tree()
    side = 1 << info.levels
    mm = T["mult"][(4 ** info.levels - 1) // 3:].reshape(side, side).astype(np.int64)
    pad = np.pad(mm, 1)
    inter = int((mm * sum(pad[1 + a:1 + a + side, 1 + b:1 + b + side] for a in (-1, 0, 1) for b in (-1, 0, 1))).sum())
    near_ms = ph["near_l2p"]
    dp = inter * 9.75 / (near_ms * 1e-3)
    # e2e: host state through nbco_step_host2 (H2D [pos|vel|acc], one PEFRL step, D2H), pinned memory
    hb = torch.empty(6 * n, dtype=torch.float64).pin_memory()
    hb.copy_(buf.cpu().reshape(-1))
    ctx.step_host2(nb.PEFRL, ev, hb.numpy(), n, par, 5e-4, 1)
    t0 = time.perf_counter()
    for _ in range(2):
        ctx.step_host2(nb.PEFRL, ev, hb.numpy(), n, par, 5e-4, 1)
    te = (time.perf_counter() - t0) / 2
    cpu = None
    try:
        from refs2d import Ref2
        if Ref2.available():
            ns, threads = 1 << 19, min(os.cpu_count() or 1, 64)
            s2 = nb.init_kv2(ns)
            p2 = nb.default_param2(ns)
            b2 = np.concatenate([s2[0], s2[1], np.zeros((ns, 2))]).copy()
            ref = Ref2(order=order, threads=threads)
            ref.eval(3, b2, ns, p2)
            t0 = time.perf_counter()
            ref.integrate(3, 3, b2, ns, p2, 5e-4, 1)
            tc = time.perf_counter() - t0
            cpu = {"value": ns / tc, "unit": "particle-steps/s", "cores": threads, "kind": "reference",
                   "sample": f"1 PEFRL step (4 evaluations of coulombOscillatorFMM_cpu) at N={ns}, p={order}, {tc:.1f} s"}
    except Exception as e:  # the checker is optional here
        cpu = {"error": str(e)}
    return {"metric": "2D fp64 FMM particle-steps/s (PEFRL)", "n": n, "order": order, "levels": int(info.levels),
            "value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps, "evals_per_step": 4, "dtype": "f64",
            "phases_ms_last_eval": {k: round(v, 4) for k, v in ph.items()},
            "roofline": {"bound": "fp64_fma", "kernel": "near_l2p2_kernel", "achieved": dp / 1e12, "peak": FP64_DFMA_PEAK / 1e12,
                         "unit": "T DP-instr/s", "frac": dp / FP64_DFMA_PEAK, "pair_interactions": inter,
                         "dp_instr_per_interaction": 9.75, "avg_launch_ms": near_ms},
            "e2e": {"value": n / te, "unit": "particle-steps/s", "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * n},
            "gpu_launches": int(info.kernel_launches - l0), "cpu_baseline": cpu}


def bench_fmm2d_sharded(nb, torch, dist, local, rank, world, n=1 << 22, order=5, steps=4):
    """BASELINE config 4 on `world` GPUs (coulomb_oscillators_b200/parallel.py: fmm2_integrate_sharded): replicated tree, the
    near-field + L2P kernel sharded over the cell-sorted particles, one NCCL all-gather of the accelerations per evaluation"""
    from coulomb_oscillators_b200.parallel import fmm2_integrate_sharded
    st = nb.init_kv2(n)
    par = nb.default_param2(n)
    ctx = nb.Context(device=local, order=order, rank=rank, world=world)
    buf = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))])).cuda().reshape(-1)
    dpar = torch.from_numpy(par).cuda()
    # a = f(x) before the loop (main.cu: compute_force), own range + all-gather = zero steps of the driver's force()
    ctx1 = nb.Context(device=local, order=order)
    ctx1.compute_force2(nb.EVAL_COULOMB_FMM2, buf.data_ptr(), n, dpar.data_ptr())
    fmm2_integrate_sharded(ctx, nb.PEFRL, buf, n, dpar.data_ptr(), 5e-4, 1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fmm2_integrate_sharded(ctx, nb.PEFRL, buf, n, dpar.data_ptr(), 5e-4, steps)
    torch.cuda.current_stream().wait_stream(torch.cuda.ExternalStream(ctx.stream))
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # parity: the sharded run against a single-rank run of the same steps on this rank
    b1 = torch.from_numpy(np.concatenate([st[0], st[1], np.zeros((n, 2))])).cuda().reshape(-1)
    ctx1.compute_force2(nb.EVAL_COULOMB_FMM2, b1.data_ptr(), n, dpar.data_ptr())
    ctx1.integrate2(nb.PEFRL, nb.EVAL_COULOMB_FMM2, b1.data_ptr(), n, dpar.data_ptr(), 5e-4, 1 + steps)
    dev = float(((buf[:4 * n] - b1[:4 * n]).abs().max() / b1[:4 * n].abs().max()).item())
    return {"metric": "2D fp64 FMM particle-steps/s (PEFRL)", "n": n, "order": order, "n_gpus": world, "value": n * steps / (ms * 1e-3),
            "ms_per_step": ms / steps, "evals_per_step": 4, "dtype": "f64", "scaling": "strong",
            "exchange": "one NCCL all-gather of the accelerations (16 B x N) per evaluation; tree replicated",
            "parity_vs_single_rank": {"max_rel_dev_pos_vel": dev, "ok": bool(dev < 1e-12)}}


def pick_ref_threads(order, cores, n_probe=1 << 20):
    """CPU_THREADS sweep of SURVEY.md section 8(d): parasort's splitter scan is O(N x threads), so more threads is not
    monotonically faster.  One steady-state leapfrog step per candidate at n_probe particles; returns (best, table)."""
    from refs import Ref
    cands = sorted({t for t in (8, 16, 32, cores) if 1 <= t <= max(cores, 8)})
    st = Ref.init_ga(n_probe)
    import coulomb_oscillators_b200 as nb
    par = nb.default_param(n_probe)
    table = {}
    for t in cands:
        ref = Ref(order=order, threads=t, unsort=0)
        buf = np.zeros(9 * n_probe, np.float32)
        buf[:6 * n_probe] = st.reshape(-1)
        ref.eval(3, buf, n_probe, par)
        ref.integrate(1, 3, buf, n_probe, par, 5e-4, 1)       # state now in tree order (steady state of a simulation)
        t0 = time.perf_counter()
        ref.integrate(1, 3, buf, n_probe, par, 5e-4, 2)
        table[t] = (time.perf_counter() - t0) / 2
    best = min(table, key=table.get)
    return best, {str(k): round(v, 4) for k, v in table.items()}


def cpu_baseline(n, order, m2l_first, bounded, steps=None, warmup=0, budget_s=150.0):
    """The reference's own CPU path (oracle/_ref = the unmodified reference compiled where it lies, kind "reference"):
    coulombOscillatorFMMKD3_cpu under leapfrog (main3.cu:65-69, integrator.cuh:68-96), initial conditions from the
    reference's own initGA, CPU_THREADS = best of a sweep.  bounded = True (the cpu_baseline leg of our arm): a sample
    of N = 2^22 particles, 3 steps.  bounded = False (--impl reference): the stated N; the number of timed steps is
    cut only if the run would not end within budget_s seconds (the line then reports the steps actually timed)."""
    from refs import Ref, Oracle
    import coulomb_oscillators_b200 as nb
    cores = os.cpu_count() or 1
    if Ref.available():
        ns = min(n, 1 << 22) if bounded else n
        threads, sweep = pick_ref_threads(order, cores)
        ref = Ref(order=order, threads=threads, unsort=0)
        buf = np.zeros(9 * ns, np.float32)
        buf[:6 * ns] = Ref.init_ga(ns).reshape(-1)
        par = nb.default_param(ns)
        t0 = time.perf_counter()
        ref.eval(3, buf, ns, par)                       # precompute accelerations (main3.cu:835-839); sorts the random input
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.integrate(1, 3, buf, ns, par, 5e-4, 1)      # first step: untimed warm-up (always), gives the step time
        t_step = time.perf_counter() - t0
        k = steps or 3
        w = max(warmup - 1, 0)
        if (k + w) * t_step > budget_s:                 # keep the whole run within a few minutes
            w = 0
            k = max(3, min(k, int(budget_s / t_step)))
        if w:
            ref.integrate(1, 3, buf, ns, par, 5e-4, w)
        t0 = time.perf_counter()
        ref.integrate(1, 3, buf, ns, par, 5e-4, k)
        t = time.perf_counter() - t0
        return {"value": ns * k / t, "unit": "particle-steps/s", "cores": threads, "kind": "reference", "n": ns, "steps": k,
                "host_cores": cores, "threads_sweep_s_per_step_at_2^20": sweep,
                "sample": f"{k} leapfrog steps of coulombOscillatorFMMKD3_cpu at N={ns} (p={order}, CPU_THREADS={threads} = best of the sweep, "
                          f"reference initGA ICs, state in tree order after {w + 1} warm-up step(s)), {t:.1f} s; first evaluation {t_first:.1f} s",
                "ms_per_step": 1e3 * t / k}
    ns = min(n, 1 << 18)
    orc = Oracle(order=order, unsort=0, m2l_first=m2l_first, tree_steps=1)
    buf = np.zeros(9 * ns, np.float32)
    buf[:6 * ns] = nb.init_ga(ns).reshape(-1)
    par = nb.default_param(ns)
    orc.eval(3, buf, ns, par)
    k = steps or 3
    t0 = time.perf_counter()
    orc.integrate(1, 3, buf, ns, par, 5e-4, k)
    t = time.perf_counter() - t0
    return {"value": ns * k / t, "unit": "particle-steps/s", "cores": 1, "kind": "port", "n": ns, "steps": k,
            "sample": f"{k} leapfrog steps of the C restatement at N={ns}, {t:.1f} s", "ms_per_step": 1e3 * t / k}


def reference_gpu_baseline(local, n=1 << 20, steps=16, timeout=240):
    """Second stated baseline ("the same generic source compiled with -arch=sm_100"): the
    reference's own GPU path, coulombOscillatorFMMKD3 under its leapfrog, timed in a separate process
    (tools/ref_gpu_baseline.py: a crash or a time-out of the reference cannot take the bench down)."""
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")   # the reference hard-wires device 0 (fmm_cart3_kdtree.cuh:1529): remap
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=vis.split(",")[local] if vis else str(local))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_baseline.py"), str(n), str(steps)],
                           capture_output=True, text=True, timeout=timeout, env=env)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": f"rc={r.returncode}: {(r.stderr or r.stdout).strip()[-200:]}"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def config1_direct_leg(nb, torch, local, n=8192, steps=100):
    """BASELINE config 1: N = 8192 direct sum, leapfrog, 100 steps.  Reference: leapfrog(coulombOscillatorDirect_cpu, ...,
    step_cpu) on the host cores (oracle/_ref); ours: nbco_integrate(LEAPFROG, COULOMB_DIRECT3) on the device."""
    out = {"n": n, "steps": steps}
    st = nb.init_ga(n)
    par = nb.default_param(n)
    ctx = nb.Context(device=local)
    buf = torch.zeros(9 * n, dtype=torch.float32, device="cuda")
    buf[:6 * n] = torch.from_numpy(st.reshape(-1)).cuda()
    dpar = torch.from_numpy(par).cuda()
    ev = nb.EVAL_COULOMB_DIRECT3
    ctx.compute_force(ev, buf.data_ptr(), n, dpar.data_ptr())
    ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, 5)
    stream = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    ctx.integrate(nb.LEAPFROG, ev, buf.data_ptr(), n, dpar.data_ptr(), 5e-4, steps)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    out["ours"] = {"ms": ms, "Ginteractions_per_s": n * n * steps / (ms * 1e-3) / 1e9, "particle_steps_per_s": n * steps / (ms * 1e-3)}
    try:
        from refs import Ref
        if Ref.available():
            threads = min(os.cpu_count() or 1, 64)
            ref = Ref(order=3, threads=threads)
            b = np.zeros(9 * n, np.float32)
            b[:6 * n] = st.reshape(-1)
            ref.eval(2, b, n, par)
            t0 = time.perf_counter()
            ref.integrate(1, 2, b, n, par, 5e-4, steps)
            t = time.perf_counter() - t0
            out["reference_cpu"] = {"s": t, "cores": threads, "Ginteractions_per_s": n * n * steps / t / 1e9,
                                    "particle_steps_per_s": n * steps / t, "kind": "reference",
                                    "what": "leapfrog(coulombOscillatorDirect_cpu, ..., step_cpu, 1) x 100 (main3.cu:53-57, integrator.cuh:68-96)"}
    except Exception as e:  # noqa: BLE001
        out["reference_cpu"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = cpu_baseline(args.n, args.order, args.m2l_first, bounded=False, steps=args.steps, warmup=args.warmup)
    out = {
        "impl": "reference", "metric": "3D FMM particle-steps/s", "value": cpu["value"], "unit": "particle-steps/s",
        "n_gpus": world, "steps": cpu["steps"], "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"3D kd-tree FMM leapfrog, N={cpu['n']}, p={args.order}, r=1, reference initGA ICs; the reference's own "
                               f"CPU path (coulombOscillatorFMMKD3_cpu, rebuilds the tree every evaluation) on the host cores",
                   "n": cpu["n"], "order": args.order, "same_n_as_requested": bool(cpu["n"] == args.n)},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--m2l-first", dest="m2l_first", type=int, default=1)
    ap.add_argument("--no-direct", dest="direct", action="store_false")
    ap.add_argument("--no-fmm2d", dest="fmm2d", action="store_false")
    ap.add_argument("--no-ref-gpu", dest="ref_gpu", action="store_false", help="skip the reference-GPU-path baseline (separate process)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
