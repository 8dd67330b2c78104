// api.cu -- the C ABI of include/nbco.h: context, configuration, evaluator dispatch, integrators.
//
// Host-side logic mirrored from the reference (paths relative to reference Simulation/):
//   compute_force / symplectic_euler / leapfrog / forestruth / pefrl   integrator.cuh:22-167
//   coulombOscillatorDirect / coulombOscillatorFMMKD3                  main3.cu:47-63
// Everything that touches particle data is a CUDA kernel (direct.cu, integrate.cu, fmm3.cu);
// there is no CPU path.

#include "common.cuh"
#include <cstdarg>
#include <cmath>

namespace nbco {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
	// same information as the reference's gpuAssert (kernel.cuh:52-65) but returned, not exit()ed
	set_error("GPUassert: %s %s %d (%s)", cudaGetErrorString(e), file, line, what);
	return e == cudaErrorMemoryAllocation ? NBCO_ERR_NOMEM : NBCO_ERR_CUDA;
}

int DevBuf::reserve(size_t need)
{
	if (need <= bytes) return NBCO_OK;
	if (p) { cudaFree(p); p = nullptr; bytes = 0; }
	size_t want = need + need / 8 + 256;
	cudaError_t e = cudaMalloc(&p, want);
	if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
	bytes = want;
	return NBCO_OK;
}

void DevBuf::release()
{
	if (p) cudaFree(p);
	p = nullptr; bytes = 0;
}

static int check_cfg(const nbco_config *c)
{
	if (c->order < 1 || c->order > NBCO2_MAX_ORDER) { set_error("order %d outside 1..%d", c->order, NBCO2_MAX_ORDER); return NBCO_ERR_INVALID; }
	if (!(c->radius > 0.f)) { set_error("radius must be > 0"); return NBCO_ERR_INVALID; }
	if (!(c->eps2 > 0.f)) { set_error("eps2 must be > 0 (the i = j term is 0 * rsqrt(eps2))"); return NBCO_ERR_INVALID; }
	if (!(c->dens_inhom > 0.f)) { set_error("dens_inhom must be > 0"); return NBCO_ERR_INVALID; }
	if (c->tree_steps < 1) { set_error("tree_steps must be >= 1"); return NBCO_ERR_INVALID; }
	if (c->world < 1 || c->rank < 0 || c->rank >= c->world) { set_error("bad rank/world %d/%d", c->rank, c->world); return NBCO_ERR_INVALID; }
	if (c->max_level < 0 || c->max_level > 30) { set_error("max_level outside 0..30"); return NBCO_ERR_INVALID; }
	return NBCO_OK;
}

static int eval_dispatch(nbco_ctx *ctx, int evaluator, float *pos, float *acc, int64_t n, const float *param)
{
	switch (evaluator)
	{
		case NBCO_EVAL_DIRECT3:
			return direct3_launch(ctx, pos, acc, n, param);
		case NBCO_EVAL_FMM3_KD:
			return fmm3_kd_launch(ctx, pos, acc, n, param, false);
		case NBCO_EVAL_COULOMB_DIRECT3:
			NBCO_TRY(direct3_launch(ctx, pos, acc, n, param));
			return add_elastic_launch(ctx, pos, acc, n, param ? param + 3 : nullptr);
		case NBCO_EVAL_COULOMB_FMM3_KD:
			return fmm3_kd_launch(ctx, pos, acc, n, param, true);
		default:
			set_error("unknown evaluator %d", evaluator);
			return NBCO_ERR_INVALID;
	}
}

static int sync(nbco_ctx *ctx)
{
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	return fmm3_harvest(ctx, nullptr); // evaluations that were only enqueued: sticky overflow / barrier flags, phase timers
}

// One step of a scheme, enqueued on the context stream (no host synchronisation inside).
// [lo, hi): the particles this context advances (peer mode: the rank's own tree-order range; else everything)
// d_energy != nullptr: the LAST update of the step (a kick for leapfrog, a drift for Forest-Ruth / PEFRL) also accumulates the
// kinetic and elastic energy of the state it produces (step_energy_launch); Euler ends with the force evaluation, its last
// update is the drift
static int scheme_step(nbco_ctx *ctx, int scheme, int evaluator, float *buf, int64_t n, const float *param, float dtf, int64_t lo, int64_t hi,
                       double *d_energy = nullptr)
{
	float *pos = buf, *vel = buf + 3*n, *acc = buf + 6*n;
	const long double dt = dtf, ds = dt * 1.0L; // scale = 1 everywhere in the reference
	auto K = [&](long double c, bool last = false)
	{
		if (last && d_energy) return step_energy_launch(ctx, vel + 3*lo, acc + 3*lo, (float)c, pos + 3*lo, true, param, hi - lo, d_energy);
		return step_launch(ctx, vel + 3*lo, acc + 3*lo, (float)c, hi - lo);
	};
	auto D = [&](long double c, bool last = false)
	{
		if (ctx->peer.active) ctx->peer.have_full = false; // the other ranges of this rank's arrays are stale from here on
		if (last && d_energy) return step_energy_launch(ctx, pos + 3*lo, vel + 3*lo, (float)c, nullptr, false, param, hi - lo, d_energy);
		return step_launch(ctx, pos + 3*lo, vel + 3*lo, (float)c, hi - lo);
	};
	auto F = [&]() { return eval_dispatch(ctx, evaluator, pos, acc, n, param); };
	switch (scheme)
	{
		case NBCO_EULER: // integrator.cuh:32-48
			NBCO_TRY(K(ds)); NBCO_TRY(D(dt, true)); NBCO_TRY(F());
			return NBCO_OK;
		case NBCO_LEAPFROG: // integrator.cuh:68-96
			NBCO_TRY(K(ds * 0.5L)); NBCO_TRY(D(dt)); NBCO_TRY(F()); NBCO_TRY(K(ds * 0.5L, true));
			return NBCO_OK;
		case NBCO_FORESTRUTH: // integrator.cuh:98-128
		{
			const long double th = 1.3512071919596576340476878089715L; // 1 / (2 - cbrt(2))
			NBCO_TRY(D(dt * th / 2)); NBCO_TRY(F());
			NBCO_TRY(K(ds * th)); NBCO_TRY(D(dt * (1 - th) / 2)); NBCO_TRY(F());
			NBCO_TRY(K(ds * (1 - 2*th))); NBCO_TRY(D(dt * (1 - th) / 2)); NBCO_TRY(F());
			NBCO_TRY(K(ds * th)); NBCO_TRY(D(dt * th / 2, true));
			return NBCO_OK;
		}
		case NBCO_PEFRL: // integrator.cuh:130-167
		{
			const long double xi = +0.1786178958448091E+00L, la = -0.2123418310626054E+00L, ch = -0.6626458266981849E-01L;
			NBCO_TRY(D(dt * xi)); NBCO_TRY(F());
			NBCO_TRY(K(ds * (1 - 2*la) / 2)); NBCO_TRY(D(dt * ch)); NBCO_TRY(F());
			NBCO_TRY(K(ds * la)); NBCO_TRY(D(dt * (1 - 2*(ch + xi)))); NBCO_TRY(F());
			NBCO_TRY(K(ds * la)); NBCO_TRY(D(dt * ch)); NBCO_TRY(F());
			NBCO_TRY(K(ds * (1 - 2*la) / 2)); NBCO_TRY(D(dt * xi, true));
			return NBCO_OK;
		}
		default:
			set_error("unknown scheme %d", scheme);
			return NBCO_ERR_INVALID;
	}
}

// nsteps steps of a scheme.  Leapfrog runs fused: K(1/2) D | F | [K(1/2) K(1/2) D | F]* | K(1/2) -- the
// same fma sequence per element as integrator.cuh:68-96 applied step by step, in fewer passes.
static int run_steps(nbco_ctx *ctx, int scheme, int evaluator, float *buf, int64_t n, const float *param, float dtf, int64_t nsteps,
                     cudaEvent_t ev_last_drift = nullptr, double *d_energy = nullptr)
{
	// peer mode (peer.cu): every rank steps its own tree-order range; the evaluator exchanges what it needs
	int64_t lo = 0, hi = n;
	const bool fmm = evaluator == NBCO_EVAL_FMM3_KD || evaluator == NBCO_EVAL_COULOMB_FMM3_KD;
	if (ctx->peer.active && fmm) nbco_shard_range(n, ctx->cfg.rank, ctx->cfg.world, &lo, &hi);
	if (scheme != NBCO_LEAPFROG || nsteps <= 0)
	{
		for (int64_t s = 0; s < nsteps; ++s)
			NBCO_TRY(scheme_step(ctx, scheme, evaluator, buf, n, param, dtf, lo, hi, s + 1 == nsteps ? d_energy : nullptr));
		return NBCO_OK;
	}
	float *pos = buf, *vel = buf + 3*n, *acc = buf + 6*n;
	const long double dt = dtf;
	const float h = (float)(dt * 1.0L * 0.5L);
	for (int64_t s = 0; s < nsteps; ++s)
	{
		NBCO_TRY(kick_drift_launch(ctx, pos + 3*lo, vel + 3*lo, acc + 3*lo, h, h, s > 0, dtf, hi - lo));
		if (ev_last_drift && s + 1 == nsteps) NBCO_CUDA(cudaEventRecord(ev_last_drift, ctx->stream));
		// only the own range moved: the other ranges of this rank's arrays are stale from here on (a rebuild must
		// fetch them from their owners, even right after nbco_peer_gather)
		if (ctx->peer.active) ctx->peer.have_full = false;
		NBCO_TRY(eval_dispatch(ctx, evaluator, pos, acc, n, param));
	}
	if (ctx->peer.active) ctx->peer.have_full = false;
	if (d_energy) return step_energy_launch(ctx, vel + 3*lo, acc + 3*lo, h, pos + 3*lo, true, param, hi - lo, d_energy);
	return step_launch(ctx, vel + 3*lo, acc + 3*lo, h, hi - lo);
}

} // namespace nbco

using namespace nbco;

extern "C" {

void nbco_default_config(nbco_config *cfg)
{
	memset(cfg, 0, sizeof(*cfg));
	cfg->device = 0;
	cfg->order = 3;          // constants.cuh:42
	cfg->radius = 1.f;       // constants.cuh:43
	cfg->eps2 = 1.e-18f;     // constants.cuh:39
	cfg->dens_inhom = 1.f;   // constants.cuh:52
	cfg->max_level = 0;      // constants.cuh:44
	cfg->tree_steps = 8;     // constants.cuh:45
	cfg->coll = 1;
	cfg->unsort = 1;         // constants.cuh:50
	cfg->m2l_first = 1;      // what the reference GPU path launches (fmm_cart3_kdtree.cuh:1668)
	cfg->rank = 0;
	cfg->world = 1;
	cfg->eps2_d = 1.e-18;    // EPS2 with SCAL = double (2D path)
	cfg->reproducible = 0;
}

int nbco_abi_version(void) { return NBCO_ABI_VERSION; }
const char *nbco_last_error(void) { return g_err; }

int nbco_create(const nbco_config *cfg, nbco_ctx **out)
{
	if (!cfg || !out) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_TRY(check_cfg(cfg));
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
	{
		set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
		return NBCO_ERR_CUDA;
	}
	if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d of %d", cfg->device, ndev); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(cfg->device));
	nbco_ctx *ctx = new nbco_ctx();
	ctx->cfg = *cfg;
	int sms = 0;
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && sms > 0)
		ctx->sm_count = sms;
	e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { delete ctx; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
	*out = ctx;
	return NBCO_OK;
}

void nbco_destroy(nbco_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->cfg.device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	peer_release(ctx, true);
	fmm3_destroy(ctx);
	fmm2_destroy(ctx);
	ctx->pos4.release(); ctx->dpart.release(); ctx->red.release(); ctx->h_state.release(); ctx->h_param.release();
	if (ctx->pinned) cudaFreeHost(ctx->pinned);
	if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_drift); cudaEventDestroy(ctx->ev_copied); }
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

int nbco_set_config(nbco_ctx *ctx, const nbco_config *cfg)
{
	if (!ctx || !cfg) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_TRY(check_cfg(cfg));
	if (cfg->device != ctx->cfg.device) { set_error("the device of a context cannot change"); return NBCO_ERR_INVALID; }
	if (ctx->peer.active && (cfg->rank != ctx->cfg.rank || cfg->world != ctx->cfg.world || cfg->unsort || cfg->order != ctx->cfg.order
	                         || cfg->max_level != ctx->cfg.max_level || cfg->dens_inhom != ctx->cfg.dens_inhom))
	{
		set_error("rank, world, order, levels and unsort = 0 are fixed while peers are attached (nbco_peer_detach first)");
		return NBCO_ERR_INVALID;
	}
	ctx->cfg = *cfg;
	return NBCO_OK;
}

int nbco_get_config(const nbco_ctx *ctx, nbco_config *cfg)
{
	if (!ctx || !cfg) { set_error("null argument"); return NBCO_ERR_INVALID; }
	*cfg = ctx->cfg;
	return NBCO_OK;
}

void *nbco_stream(nbco_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int nbco_track_ids(nbco_ctx *ctx, int32_t *d_ids)
{
	if (!ctx) { set_error("null context"); return NBCO_ERR_INVALID; }
	if (d_ids && ctx->peer.active) { set_error("nbco_track_ids is not available while peers are attached"); return NBCO_ERR_INVALID; }
	ctx->ids = d_ids;
	return NBCO_OK;
}

#define ENTER(ctx)                                                         \
	if (!(ctx)) { set_error("null context"); return NBCO_ERR_INVALID; }    \
	NBCO_CUDA(cudaSetDevice((ctx)->cfg.device));

int nbco_force_direct3(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER(ctx);
	NBCO_TRY(direct3_launch(ctx, (const float *)d_pos, (float *)d_acc, n, (const float *)d_param));
	return sync(ctx);
}

int nbco_force_fmm3_kd(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER(ctx);
	NBCO_TRY(fmm3_kd_launch(ctx, (float *)d_pos, (float *)d_acc, n, (const float *)d_param, false));
	return sync(ctx);
}

int nbco_coulomb_direct3(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER(ctx);
	NBCO_TRY(eval_dispatch(ctx, NBCO_EVAL_COULOMB_DIRECT3, (float *)d_pos, (float *)d_acc, n, (const float *)d_param));
	return sync(ctx);
}

int nbco_coulomb_fmm3_kd(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER(ctx);
	NBCO_TRY(eval_dispatch(ctx, NBCO_EVAL_COULOMB_FMM3_KD, (float *)d_pos, (float *)d_acc, n, (const float *)d_param));
	return sync(ctx);
}

int nbco_add_elastic(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_k3)
{
	ENTER(ctx);
	NBCO_TRY(add_elastic_launch(ctx, (const float *)d_pos, (float *)d_acc, n, (const float *)d_k3));
	return sync(ctx);
}

int nbco_step(nbco_ctx *ctx, void *d_b, const void *d_a, float ds, int64_t n)
{
	ENTER(ctx);
	NBCO_TRY(step_launch(ctx, (float *)d_b, (const float *)d_a, ds, n));
	ctx->peer.have_full = false; // peer mode: the caller advances (part of) the state behind the evaluator's back
	return sync(ctx);
}

int nbco_compute_force(nbco_ctx *ctx, int evaluator, void *d_buf, int64_t n, const void *d_param)
{
	ENTER(ctx);
	float *buf = (float *)d_buf;
	NBCO_TRY(eval_dispatch(ctx, evaluator, buf, buf + 6*n, n, (const float *)d_param));
	return sync(ctx);
}

int nbco_integrate(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                   const void *d_param, double dt, int64_t nsteps)
{
	ENTER(ctx);
	const float dtf = (float)dt; // main3.cu:231
	NBCO_TRY(run_steps(ctx, scheme, evaluator, (float *)d_buf, n, (const float *)d_param, dtf, nsteps));
	return sync(ctx);
}

int nbco_integrate_energy(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                          const void *d_param, double dt, int64_t nsteps, double *h_kin_el)
{
	ENTER(ctx);
	if (!h_kin_el || nsteps < 1) { set_error("nbco_integrate_energy: needs an output and at least one step"); return NBCO_ERR_INVALID; }
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>();
	NBCO_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(double), ctx->stream));
	NBCO_TRY(run_steps(ctx, scheme, evaluator, (float *)d_buf, n, (const float *)d_param, (float)dt, nsteps, nullptr, d));
	NBCO_CUDA(cudaMemcpyAsync(h_kin_el, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	return sync(ctx);
}

int nbco_mean_rel_err(nbco_ctx *ctx, const void *d_a, const void *d_ref, int64_t n, double *h_mean, double *h_max)
{
	ENTER(ctx);
	return rel_err_launch(ctx, (const float *)d_a, (const float *)d_ref, n, h_mean, h_max);
}

int nbco_energy(nbco_ctx *ctx, const void *d_buf, int64_t n, const void *d_param, double *h_out3)
{
	ENTER(ctx);
	if (!h_out3) { set_error("null output"); return NBCO_ERR_INVALID; }
	NBCO_TRY(kinetic_elastic_launch(ctx, (const float *)d_buf, n, (const float *)d_param, h_out3));
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>() + 2;
	NBCO_TRY(pair_energy_launch(ctx, (const float *)d_buf, n, d));
	double pe = 0.0;
	float scale = 1.f;
	if (d_param) NBCO_CUDA(cudaMemcpyAsync(&scale, d_param, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaMemcpyAsync(&pe, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_TRY(sync(ctx));
	h_out3[2] = (double)scale * pe;
	return NBCO_OK;
}

// ---- host-buffer entry points ----

static int stage(nbco_ctx *ctx, int64_t n, const float *h_param, float **d_buf, float **d_param)
{
	NBCO_TRY(ctx->h_state.reserve(sizeof(float) * 9 * (size_t)n));
	NBCO_TRY(ctx->h_param.reserve(sizeof(float) * 8));
	*d_buf = ctx->h_state.as<float>();
	*d_param = nullptr;
	if (h_param)
	{
		*d_param = ctx->h_param.as<float>();
		NBCO_CUDA(cudaMemcpyAsync(*d_param, h_param, 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	}
	return NBCO_OK;
}

int nbco_eval_host(nbco_ctx *ctx, int evaluator, float *h_pos, float *h_vel, float *h_acc,
                   int64_t n, const float *h_param)
{
	ENTER(ctx);
	if (!h_pos || !h_acc || n <= 0) { set_error("bad host buffers"); return NBCO_ERR_INVALID; }
	float *d_buf, *d_param;
	NBCO_TRY(stage(ctx, n, h_param, &d_buf, &d_param));
	const size_t vb = sizeof(float) * 3 * (size_t)n;
	NBCO_CUDA(cudaMemcpyAsync(d_buf, h_pos, vb, cudaMemcpyHostToDevice, ctx->stream));
	if (h_vel) NBCO_CUDA(cudaMemcpyAsync(d_buf + 3*n, h_vel, vb, cudaMemcpyHostToDevice, ctx->stream));
	else NBCO_CUDA(cudaMemsetAsync(d_buf + 3*n, 0, vb, ctx->stream));
	NBCO_TRY(eval_dispatch(ctx, evaluator, d_buf, d_buf + 6*n, n, d_param));
	NBCO_CUDA(cudaMemcpyAsync(h_acc, d_buf + 6*n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	const bool fmm = evaluator == NBCO_EVAL_FMM3_KD || evaluator == NBCO_EVAL_COULOMB_FMM3_KD;
	if (fmm && !ctx->cfg.unsort)
	{
		NBCO_CUDA(cudaMemcpyAsync(h_pos, d_buf, vb, cudaMemcpyDeviceToHost, ctx->stream));
		if (h_vel) NBCO_CUDA(cudaMemcpyAsync(h_vel, d_buf + 3*n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	}
	return sync(ctx);
}

static int ensure_copy_stream(nbco_ctx *ctx)
{
	if (ctx->copy_stream) return NBCO_OK;
	NBCO_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	NBCO_CUDA(cudaEventCreateWithFlags(&ctx->ev_drift, cudaEventDisableTiming));
	NBCO_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
	return NBCO_OK;
}

int nbco_run_host(nbco_ctx *ctx, int scheme, int evaluator, float *h_pos_vel, float *h_acc, int64_t n,
                  const float *h_param, double dt, int64_t nsteps)
// The reference's resume-from-snapshot flow (main3.cu:629-658,835-858) for a host-resident state: upload [pos | vel],
// a = f(x), nsteps steps, read [pos | vel] back.  The copies overlap the force evaluations where the data flow allows:
// the first evaluation needs the positions only, so the velocities are uploaded on a second stream meanwhile (unless
// that evaluation rebuilds the tree and therefore permutes the velocities); after the last drift the positions are
// final, so they are read back while the last evaluation runs (unless it permutes them).
{
	ENTER(ctx);
	if (!h_pos_vel || n <= 0) { set_error("bad host buffers"); return NBCO_ERR_INVALID; }
	float *d_buf, *d_param;
	NBCO_TRY(stage(ctx, n, h_param, &d_buf, &d_param));
	NBCO_TRY(ensure_copy_stream(ctx));
	const size_t vb = sizeof(float) * 3 * (size_t)n;
	const bool fmm = evaluator == NBCO_EVAL_FMM3_KD || evaluator == NBCO_EVAL_COULOMB_FMM3_KD;
	const bool peer = ctx->peer.active;
	// uploads: positions on the compute stream, velocities on the copy stream
	NBCO_CUDA(cudaEventRecord(ctx->ev_drift, ctx->stream));                 // d_buf is free (previous call finished on this stream)
	NBCO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_drift, 0));
	NBCO_CUDA(cudaMemcpyAsync(d_buf, h_pos_vel, vb, cudaMemcpyHostToDevice, ctx->stream));
	NBCO_CUDA(cudaMemcpyAsync(d_buf + 3*n, h_pos_vel + 3*n, vb, cudaMemcpyHostToDevice, ctx->copy_stream));
	NBCO_CUDA(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
	const bool first_permutes = fmm && (peer || fmm3_next_rebuilds(ctx, n));
	if (first_permutes) NBCO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
	NBCO_TRY(eval_dispatch(ctx, evaluator, d_buf, d_buf + 6*n, n, d_param)); // main3.cu:835-839
	if (!first_permutes) NBCO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
	const float dtf = (float)dt;
	// will the LAST evaluation of the steps permute the state?  (evaluation index nsteps - 1 from here)
	const bool early_pos = scheme == NBCO_LEAPFROG && nsteps >= 1 && !peer && (!fmm || !fmm3_rebuilds_in(ctx, n, nsteps - 1));
	NBCO_TRY(run_steps(ctx, scheme, evaluator, d_buf, n, d_param, dtf, nsteps, early_pos ? ctx->ev_drift : nullptr));
	if (early_pos)
	{
		NBCO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_drift, 0));
		NBCO_CUDA(cudaMemcpyAsync(h_pos_vel, d_buf, vb, cudaMemcpyDeviceToHost, ctx->copy_stream));
		NBCO_CUDA(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
		NBCO_CUDA(cudaMemcpyAsync(h_pos_vel + 3*n, d_buf + 3*n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	}
	else
		NBCO_CUDA(cudaMemcpyAsync(h_pos_vel, d_buf, 2 * vb, cudaMemcpyDeviceToHost, ctx->stream));
	if (h_acc) NBCO_CUDA(cudaMemcpyAsync(h_acc, d_buf + 6*n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	if (early_pos) NBCO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
	return sync(ctx);
}

int nbco_step_host(nbco_ctx *ctx, int scheme, int evaluator, float *h_buf, int64_t n,
                   const float *h_param, double dt, int64_t nsteps)
{
	ENTER(ctx);
	if (!h_buf || n <= 0) { set_error("bad host buffers"); return NBCO_ERR_INVALID; }
	float *d_buf, *d_param;
	NBCO_TRY(stage(ctx, n, h_param, &d_buf, &d_param));
	const size_t vb = sizeof(float) * 3 * (size_t)n;
	NBCO_CUDA(cudaMemcpyAsync(d_buf, h_buf, 3 * vb, cudaMemcpyHostToDevice, ctx->stream));
	const float dtf = (float)dt;
	// One leapfrog step whose evaluation does not permute the state: the positions are final after the drift, so
	// their read-back (a third of the D2H traffic) runs on a second stream while the forces are computed.
	const bool fmm = evaluator == NBCO_EVAL_FMM3_KD || evaluator == NBCO_EVAL_COULOMB_FMM3_KD;
	const bool overlap = scheme == NBCO_LEAPFROG && nsteps == 1 && !ctx->peer.active && (!fmm || !fmm3_next_rebuilds(ctx, n));
	if (overlap)
	{
		NBCO_TRY(ensure_copy_stream(ctx));
		float *pos = d_buf, *vel = d_buf + 3*n, *acc = d_buf + 6*n;
		const float h = (float)((long double)dtf * 0.5L);
		NBCO_TRY(kick_drift_launch(ctx, pos, vel, acc, h, h, false, dtf, n));      // K(1/2) D, like run_steps
		NBCO_CUDA(cudaEventRecord(ctx->ev_drift, ctx->stream));
		NBCO_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_drift, 0));
		NBCO_CUDA(cudaMemcpyAsync(h_buf, pos, vb, cudaMemcpyDeviceToHost, ctx->copy_stream));
		NBCO_CUDA(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
		NBCO_TRY(eval_dispatch(ctx, evaluator, pos, acc, n, d_param));             // F (reads pos only)
		NBCO_TRY(step_launch(ctx, vel, acc, h, n));                                 // K(1/2)
		NBCO_CUDA(cudaMemcpyAsync(h_buf + 3*n, vel, 2 * vb, cudaMemcpyDeviceToHost, ctx->stream));
		NBCO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
		return sync(ctx);
	}
	NBCO_TRY(run_steps(ctx, scheme, evaluator, d_buf, n, d_param, dtf, nsteps));
	NBCO_CUDA(cudaMemcpyAsync(h_buf, d_buf, 3 * vb, cudaMemcpyDeviceToHost, ctx->stream));
	return sync(ctx);
}

void nbco_shard_range(int64_t n, int32_t rank, int32_t world, int64_t *begin, int64_t *end)
{
	// ceil(n*i/w), the split rule of evalBox (fmm_cart3_kdtree.cuh:117-118)
	auto cut = [&](int64_t i) -> int64_t { return i <= 0 ? 0 : (i >= world ? n : (n * i - 1) / world + 1); };
	if (begin) *begin = cut(rank);
	if (end) *end = cut((int64_t)rank + 1);
}

} // extern "C"
