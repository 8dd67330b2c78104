/* nbco.h -- C ABI of the B200-native force-evaluation / time-stepping path.
 *
 * Drop-in boundary for locuoco/coulomb_oscillators (reference paths are relative to its
 * Simulation/ directory).  Every entry point replaces one reference interface and is what a
 * maintainer would bind in place of it (see INTEGRATION.md for the reference-side stub):
 *
 *   reference plugin type   void f(VEC *p, VEC *a, int n, const SCAL *param)   integrator.cuh:22
 *   reference step type     void step(VEC *b, const VEC *a, SCAL ds, int n)    integrator.cuh:34, kernel.cuh:100
 *
 * Conventions kept from the reference:
 *   - VEC = float3 stored AoS (12 B, x y z), SCAL = float (constants.cuh:22-28);
 *   - the state buffer is one allocation [pos(n) | vel(n) | acc(n)] (integrator.cuh:24);
 *   - param is a DEVICE array {xi/N, 0, 0, kx, ky, kz} (main3.cu:685-692); the Coulomb evaluators
 *     read param[0], the elastic term reads param+3; param == NULL means "unscaled"
 *     (fmm_cart3_kdtree.cuh:1743, kernel.cuh:148-151);
 *   - with unsort == 0 the FMM evaluator permutes pos AND the velocities stored at pos+n in
 *     place into tree order on tree-rebuild calls (fmm_cart3_kdtree.cuh:1359-1360,1758-1759);
 *   - evaluators return after the work has completed on the device (:1762-1763).
 * Differences, all deliberate: configuration is an explicit struct instead of the mutable
 * globals of constants.cuh:36-52; scratch memory belongs to a context instead of function
 * statics; errors are returned (never exit()); sizes are 64-bit.
 *
 * No torch types, no C++ types: plain pointers and sizes only.  Pointers named d_* are device
 * pointers on the context's device, h_* are host pointers.  There is no CPU fallback: every
 * call needs a CUDA device and fails with NBCO_ERR_CUDA otherwise.
 */
#ifndef NBCO_H
#define NBCO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBCO_ABI_VERSION 3

typedef struct nbco_ctx nbco_ctx;

enum nbco_status
{
	NBCO_OK = 0,
	NBCO_ERR_INVALID = 1,   /* bad argument / unsupported configuration */
	NBCO_ERR_CUDA = 2,      /* CUDA runtime error, text in nbco_last_error() */
	NBCO_ERR_NOMEM = 3,
	NBCO_ERR_OVERFLOW = 4   /* an interaction list did not fit (never silently truncated) */
};

/* Replaces the globals of constants.cuh:36-52 (+ CLI options of main3.cu:263-305). */
typedef struct nbco_config
{
	int32_t device;        /* CUDA device ordinal (reference: hard-wired 0, fmm_cart3_kdtree.cuh:1529) */
	int32_t order;         /* fmm_order, "-p"; default 3 (constants.cuh:42); 1..NBCO_MAX_ORDER */
	float   radius;        /* tree_radius, "-r"; default 1 */
	float   eps2;          /* EPS2 = eps^2, "-eps"; default 1e-18; must be > 0 */
	float   dens_inhom;    /* dens_inhom, "-i"; default 1 */
	int32_t max_level;     /* tree_L, "-maxlevel"; 0 = automatic (fmm_cart3_kdtree.cuh:1508-1516) */
	int32_t tree_steps;    /* tree_steps: rebuild the kd-tree every this many evaluations; default 8 */
	int32_t coll;          /* 0 = "-ncoll": skip the P2P near field */
	int32_t unsort;        /* b_unsort: 1 = return pos/acc in input order, 0 = leave tree order */
	int32_t m2l_first;     /* 1 = MAC before leaf test (reference GPU kernel, :504-534),
	                          0 = leaf test first (reference CPU path, :586-598) */
	int32_t rank;          /* multi-GPU: this process' rank and the world size; the evaluators then */
	int32_t world;         /* compute only this rank's shard of targets (see nbco_shard_range)      */
	double  eps2_d;        /* EPS2 of the 2D fp64 path (SCAL = double, constants.cuh:39); default 1e-18;
	                          0 = use (double)eps2 */
	int32_t reproducible;  /* 3D kd FMM, orders 1..6.  0 (default): pair lists, float atomics like the reference (forces
	                          differ from run to run in the last bits).  1: lists bucketed by target and sorted, every
	                          local expansion and acceleration is one sum in registers: bit-reproducible forces, no
	                          atomics, about 1.6x the evaluation time */
	int32_t reserved0;
} nbco_config;

#define NBCO_MAX_ORDER 10   /* 3D kd-tree FMM: 1..6 unrolled templates, 7..10 runtime-order loops (main3.cu:790-811 sweeps 1..10) */
#define NBCO2_MAX_ORDER 10  /* 2D FMM (the reference's -test loop runs p = 1..10, main.cu:844) */

void nbco_default_config(nbco_config *cfg);
int  nbco_abi_version(void);
const char *nbco_last_error(void);

int  nbco_create(const nbco_config *cfg, nbco_ctx **out);
void nbco_destroy(nbco_ctx *ctx);
int  nbco_set_config(nbco_ctx *ctx, const nbco_config *cfg);    /* re-plans lazily, like :1502 */
int  nbco_get_config(const nbco_ctx *ctx, nbco_config *cfg);
/* The CUDA stream every call of this context is enqueued on (a cudaStream_t), for callers
 * that want to bracket calls with their own events. */
void *nbco_stream(nbco_ctx *ctx);

/* ---- evaluators: the reference plugin type, device pointers ---- */

/* direct3 (direct.cuh:233-245): a_i = param[0] * sum_j d (|d|^2+eps2)^(-3/2), d = x_i - x_j.
 * Multi-GPU: d_pos holds all n sources; only targets of this rank's shard are written. */
int nbco_force_direct3(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_param);

/* fmm_cart3_kdtree (fmm_cart3_kdtree.cuh:1478-1771). d_pos is followed by the velocities at
 * d_pos + n (float3) when unsort == 0, exactly like the reference. */
int nbco_force_fmm3_kd(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);

/* coulombOscillatorDirect / coulombOscillatorFMMKD3 (main3.cu:47-63): evaluator + elastic term */
int nbco_coulomb_direct3(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);
int nbco_coulomb_fmm3_kd(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);

/* add_elastic (kernel.cuh:145-152): a -= k o x, d_k3 = device float[3] or NULL (k = 1) */
int nbco_add_elastic(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_k3);

/* step (kernel.cuh:100-104): b += a * ds */
int nbco_step(nbco_ctx *ctx, void *d_b, const void *d_a, float ds, int64_t n);

/* ---- integrators (integrator.cuh:32-167) over the state buffer [pos|vel|acc] ---- */
enum nbco_scheme  { NBCO_EULER = 0, NBCO_LEAPFROG = 1, NBCO_FORESTRUTH = 2, NBCO_PEFRL = 3 };
enum nbco_evaluator { NBCO_EVAL_DIRECT3 = 0, NBCO_EVAL_FMM3_KD = 1,
                      NBCO_EVAL_COULOMB_DIRECT3 = 2, NBCO_EVAL_COULOMB_FMM3_KD = 3 };

/* compute_force (integrator.cuh:22-28) */
int nbco_compute_force(nbco_ctx *ctx, int evaluator, void *d_buf, int64_t n, const void *d_param);
/* nsteps steps of the chosen scheme; the caller does the initial compute_force like
 * main3.cu:835-839.  dt is rounded to float first (main3.cu:231). */
int nbco_integrate(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                   const void *d_param, double dt, int64_t nsteps);

/* nbco_integrate whose LAST update (the closing kick of leapfrog, the closing drift of the other schemes) is fused with the
 * energy reduction: h_kin_el[0] = sum 1/2 v^2, h_kin_el[1] = 1/2 sum k o x^2 of the final state, accumulated in double in the
 * same pass that writes it (no extra sweep over the state).  Peer mode: the sums cover the rank's own range.  (Euler ends
 * with a force evaluation: its energies belong to the state after the drift, i.e. the final positions and velocities.) */
int nbco_integrate_energy(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                          const void *d_param, double dt, int64_t nsteps, double *h_kin_el);

/* ---- diagnostics ---- */
/* mean over i of |a-ref| / sqrt(|ref|^2 + 1e-18) (rel_diff1, reductions.cuh:37-42; the index
 * bug of relerrReduce2 :89 is not reproduced).  Result on the host. */
int nbco_mean_rel_err(nbco_ctx *ctx, const void *d_a, const void *d_ref, int64_t n, double *h_mean, double *h_max);
/* Energy of the Coulomb-oscillator system (absent in the reference):
 * h_out[0] = sum 1/2 v^2, h_out[1] = 1/2 sum k o x^2, h_out[2] = param[0] * sum_{i<j} (d^2+eps2)^(-1/2)
 * (pair term by direct summation, O(n^2)). */
int nbco_energy(nbco_ctx *ctx, const void *d_buf, int64_t n, const void *d_param, double *h_out3);

/* ---- host-buffer convenience (the e2e path: H2D + evaluate + D2H in one call) ---- */
/* h_param = host float[6] or NULL.  evaluator as above.  h_pos (and h_vel when given and
 * unsort == 0) are updated like the device call would update them. */
int nbco_eval_host(nbco_ctx *ctx, int evaluator, float *h_pos, float *h_vel, float *h_acc,
                   int64_t n, const float *h_param);
/* Whole run from host state [pos|vel] (the binary state-file layout, main3.cu:629-652):
 * H2D, compute_force, nsteps steps, D2H of pos|vel (acc too when h_acc != NULL). */
int nbco_run_host(nbco_ctx *ctx, int scheme, int evaluator, float *h_pos_vel, float *h_acc, int64_t n,
                  const float *h_param, double dt, int64_t nsteps);

/* nsteps steps from a host state buffer [pos|vel|acc] (9n floats, acc valid on entry, e.g. from a
 * previous call): H2D of all three arrays, nsteps steps of the scheme, D2H of all three.  With
 * nsteps = 1 this is the end-to-end cost of one step when the state lives on the host. */
int nbco_step_host(nbco_ctx *ctx, int scheme, int evaluator, float *h_buf, int64_t n,
                   const float *h_param, double dt, int64_t nsteps);

/* ---- introspection of the last FMM evaluation (parity tests, profiling) ---- */
typedef struct nbco_fmm_info
{
	int32_t levels;        /* L: leaves are level L, root level 0 */
	int32_t order;
	int64_t n;
	int64_t nodes;         /* 2^(L+1)-1 */
	int64_t p2p_pairs, m2l_pairs;
	int32_t off_m, off_l;  /* floats per node: symmetric orders 0..p-1, traceless orders 0..p */
	int32_t rebuilt;       /* 1 if the last evaluation rebuilt the tree */
	int32_t mlt_max;
	int64_t kernel_launches; /* kernels launched by this context so far */
	int32_t counter;       /* FMM evaluations since the last re-plan; the next one rebuilds iff counter % tree_steps == 0 */
	int32_t reserved;
} nbco_fmm_info;

int nbco_fmm_get_info(nbco_ctx *ctx, nbco_fmm_info *info);
/* Copies to host whatever pointers are non-NULL.  Layouts follow fmmTree_kd
 * (fmm_cart3_kdtree.cuh:25-31, slab :1552-1560): center/lbound/rbound float3 per node;
 * mpole off_m floats per node (symmetric storage, fmm_cart_base3.cuh:180-213); local off_l
 * floats per node (traceless storage, :185-232); mult/index/splitdim int per node (index has
 * one extra entry per level as in evalBox, not exported: index[node] only).
 * perm[n]: sorted position -> position in the array passed to the last rebuild (d_unsort).
 * Lists are int2 pairs sorted ascending by (x, y) so that they can be compared as sets. */
int nbco_fmm_get_tree(nbco_ctx *ctx, float *h_center, float *h_lbound, float *h_rbound,
                      float *h_mpole, float *h_local, int32_t *h_mult, int32_t *h_index,
                      int32_t *h_splitdim, int32_t *h_perm);
int nbco_fmm_get_lists(nbco_ctx *ctx, int32_t *h_p2p_pairs, int64_t p2p_cap,
                       int32_t *h_m2l_pairs, int64_t m2l_cap);
/* Per-phase device times (ms) of the last FMM evaluation, measured with CUDA events on the
 * context stream: names[i] points to a static string. Returns the number of phases. */
int nbco_fmm_get_phase_ms(nbco_ctx *ctx, const char **names, float *ms, int cap);
/* The same per phase, summed over all FMM evaluations since the last reset (ms), with the
 * number of evaluations and how many of them rebuilt the tree in h_evals[0], h_evals[1].
 * reset != 0 clears the sums after reading.  Returns the number of phases. */
int nbco_fmm_phase_totals(nbco_ctx *ctx, const char **names, double *ms, int cap, int64_t *h_evals, int reset);

/* ==== 2D fp64 path (the reference's `nbco` binary: SCAL = double, VEC = double2, DIM = 2) ====
 * State buffer [pos(n) | vel(n) | acc(n)] of double2; param is a DEVICE array {xi/N, 0, kx, ky}
 * (main.cu:803-808): the Coulomb evaluators read param[0], the elastic term reads param+2.
 * nbco_force_fmm2 replaces fmm_cart (fmm_cart.cuh:395-545): like the reference it ALWAYS permutes pos
 * and the velocities stored at pos+n into cell order (:500-505; there is no unsort in 2D) and rebuilds
 * the grid on every call.  cfg.order 1..NBCO2_MAX_ORDER, cfg.radius is truncated to int (:398),
 * cfg.eps2_d, cfg.dens_inhom, cfg.coll and cfg.max_level (0 = automatic, :416-418) apply. */
int nbco_force_direct2(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_param);  /* direct2, direct.cuh:166-179 */
int nbco_force_fmm2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);
/* coulombOscillatorDirect / coulombOscillatorFMM (main.cu:69-89) */
int nbco_coulomb_direct2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);
int nbco_coulomb_fmm2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param);
int nbco_add_elastic2(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_k2);
int nbco_step2(nbco_ctx *ctx, void *d_b, const void *d_a, double ds, int64_t n);

enum nbco_evaluator2 { NBCO_EVAL_DIRECT2 = 4, NBCO_EVAL_FMM2 = 5, NBCO_EVAL_COULOMB_DIRECT2 = 6, NBCO_EVAL_COULOMB_FMM2 = 7 };
int nbco_compute_force2(nbco_ctx *ctx, int evaluator, void *d_buf, int64_t n, const void *d_param);
/* integrator.cuh:32-167 with SCAL = double; schemes as nbco_scheme */
int nbco_integrate2(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                    const void *d_param, double dt, int64_t nsteps);
int nbco_mean_rel_err2(nbco_ctx *ctx, const void *d_a, const void *d_ref, int64_t n, double *h_mean, double *h_max);
/* h_out[0] = sum 1/2 v^2, h_out[1] = 1/2 sum k o x^2, h_out[2] = param[0] * sum_{i<j} -1/2 log(d^2+eps2)
 * (the potential whose gradient is the 2D kernel d/(d^2+eps2), direct.cuh:23-31); O(n^2). */
int nbco_energy2(nbco_ctx *ctx, const void *d_buf, int64_t n, const void *d_param, double *h_out3);
/* host state [pos|vel] (4n doubles, the state-file layout): H2D, compute_force, nsteps steps, D2H */
int nbco_run_host2(nbco_ctx *ctx, int scheme, int evaluator, double *h_pos_vel, double *h_acc, int64_t n,
                   const double *h_param, double dt, int64_t nsteps);
/* nsteps steps from a host buffer [pos|vel|acc] (6n doubles, acc valid on entry): H2D, steps, D2H */
int nbco_step_host2(nbco_ctx *ctx, int scheme, int evaluator, double *h_buf, int64_t n,
                    const double *h_param, double dt, int64_t nsteps);

typedef struct nbco_fmm2_info
{
	int32_t levels;        /* L: the leaf grid is 2^L x 2^L, levels 2..L carry expansions */
	int32_t order;
	int64_t n;
	int64_t nodes;         /* (4^(L+1)-1)/3, node (i,j) of level l at (4^l-1)/3 + i*2^l + j (fmm_cart.cuh tree_beg) */
	int32_t coeffs;        /* complex coefficients per node and expansion: order + 1 */
	int32_t reserved;
	int64_t kernel_launches;
	int64_t evals;
} nbco_fmm2_info;
int nbco_fmm2_levels(int64_t n, int32_t order, double dens_inhom);   /* fmm_cart.cuh:416-418 */
int nbco_fmm2_get_info(nbco_ctx *ctx, nbco_fmm2_info *info);
/* Copies to host whatever is non-NULL.  center: double2 per node.  mpole / local: `coeffs` complex numbers
 * (re, im) per node -- the harmonic content of the reference's tuples: mpole[q] = sum_k binom(q,k) i^k M_q[k]
 * of the symmetric tuple M_q (fmm_cart_base.cuh:111), local[n] = L_n[0] + i L_n[1] of the traceless tuple
 * (:116).  mult: int per node.  leaf_index: 4^L + 1 ints, first particle of every leaf cell (indexLeaves).
 * perm[n]: sorted position -> position in the array passed to the last call. */
int nbco_fmm2_get_tree(nbco_ctx *ctx, double *h_center, double *h_mpole, double *h_local, int32_t *h_mult,
                       int32_t *h_leaf_index, int32_t *h_perm);
int nbco_fmm2_get_phase_ms(nbco_ctx *ctx, const char **names, float *ms, int cap);

/* 2D initial conditions and state files (host side, byte-compatible with main.cu) */
/* initGA (main.cu:147-170) / initKV (:120-145) with the fixed seed and discard of main.cu:779-784;
 * h_pos_vel = 4n doubles.  GA: std.dev. x2, u2.  KV: semi-axes A2, depressed phase advances omega2. */
int nbco_init_ga2(double *h_pos_vel, int64_t n, const double *x2, const double *u2);
int nbco_init_kv2(double *h_pos_vel, int64_t n, const double *A2, const double *omega2);
/* default beam of main.cu:272,294-313 from omega0 and the emittances: out = {A.x, A.y, omega.x, omega.y, xi} */
int nbco_beam_params2(const double *omega0_2, const double *emit2, double tune_dep_y, double *out5);
int nbco_state_read2(const char *path, double **h_pos_vel, int64_t *n);  /* caller frees with nbco_free */
int nbco_state_write2(const char *path, const double *h_pos_vel, int64_t n);

/* ---- particle identity (optional) ----
 * With unsort = 0 the evaluator permutes pos / vel into tree order at every rebuild and the reference loses the
 * particles' identity there (d_unsort is reset to iota, fmm_cart3_kdtree.cuh:1626).  nbco_track_ids registers a
 * device array of n int32 that is permuted together with pos / vel from then on (ids[j] = id of the particle now
 * stored at j); pass NULL to stop.  The caller initialises it (e.g. 0 .. n-1) and owns the memory.  Not available
 * while peers are attached. */
int nbco_track_ids(nbco_ctx *ctx, int32_t *d_ids);

/* ---- multi-GPU helpers ---- */
/* Target shard [begin, end) of rank r of w over n items: the kd-tree's own equal split
 * ceil(n*r/w) (fmm_cart3_kdtree.cuh:117-118). */
void nbco_shard_range(int64_t n, int32_t rank, int32_t world, int64_t *begin, int64_t *end);

/* ---- multi-GPU over NVLink peer memory (one process per GPU; csrc/peer.cu) ----
 * Rank r of w = 2^g (cfg.rank / cfg.world, unsort = 0) owns the subtree of kd node (g, r), i.e. the tree-order
 * range nbco_shard_range(n, r, w) of particles.  Every rank publishes its node centres, multipoles and its range of
 * positions through CUDA IPC; the evaluator reads remote nodes / leaves straight from their owners and the ranks
 * meet at a flag barrier in peer memory (no host round trip).  Protocol, on every rank:
 *   nbco_peer_export(ctx, n, handles)            192 bytes = 3 cudaIpcMemHandle_t, to be sent to every other rank
 *   nbco_peer_attach(ctx, q, handles_of_q)       for every q != rank
 *   nbco_peer_commit(ctx)
 * From then on nbco_force_fmm3_kd / nbco_coulomb_fmm3_kd / nbco_compute_force / nbco_integrate must be
 * called by all ranks together with the same n (all four schemes).  The FIRST evaluation expects the full, identical [pos | vel] on
 * every rank; afterwards every rank holds (and steps) only its own range, and tree rebuilds fetch the other
 * ranges from their owners.  nbco_peer_gather leaves the full [pos | vel | acc] on every rank again. */
int nbco_peer_export(nbco_ctx *ctx, int64_t n, void *h_handles192);
int nbco_peer_attach(nbco_ctx *ctx, int32_t peer_rank, const void *h_handles192);
/* same process instead of IPC: wire rank `peer_rank` = `other` directly (several GPUs driven by one process, or ranks
 * emulated on one device from different host threads) */
int nbco_peer_attach_local(nbco_ctx *ctx, int32_t peer_rank, nbco_ctx *other);
int nbco_peer_commit(nbco_ctx *ctx);
int nbco_peer_barrier(nbco_ctx *ctx);   /* barrier + host synchronisation; NBCO_ERR_CUDA if a rank did not arrive */
int nbco_peer_gather(nbco_ctx *ctx, void *d_buf, int64_t n);
int nbco_peer_detach(nbco_ctx *ctx);

/* ---- initial conditions and state files (host side, byte-compatible with main3.cu) ---- */
/* initGA with the reference's fixed seed (main3.cu:114-137,662-664): h_pos_vel = 6n floats */
int nbco_init_ga(float *h_pos_vel, int64_t n, const float *sigma_x3, const float *sigma_u3);
/* the "-test" uniform cube [-1,1]^3 drawn after initGA from the same generator (main3.cu:94-112,665-666) */
int nbco_init_test_cube(float *h_pos_vel, int64_t n, const float *sigma_x3, const float *sigma_u3);
int nbco_state_read(const char *path, float **h_pos_vel, int64_t *n);   /* caller frees with nbco_free */
int nbco_state_write(const char *path, const float *h_pos_vel, int64_t n);
void nbco_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* NBCO_H */
