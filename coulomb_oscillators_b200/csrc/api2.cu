// api2.cu -- C ABI of the 2D fp64 path (include/nbco.h, "2D fp64 path"): evaluator dispatch,
// integrators, host-buffer entry points.  Host logic mirrored from the reference:
//   coulombOscillatorDirect / coulombOscillatorFMM                     main.cu:69-89
//   compute_force / symplectic_euler / leapfrog / forestruth / pefrl   integrator.cuh:22-167 (SCAL = double)
// All particle data is touched by CUDA kernels only (fmm2.cu).

#include "common.cuh"

namespace nbco {

static int eval2(nbco_ctx *ctx, int evaluator, double *pos, double *acc, int64_t n, const double *param)
{
	switch (evaluator)
	{
		case NBCO_EVAL_DIRECT2:
			return direct2_launch(ctx, pos, acc, n, param);
		case NBCO_EVAL_FMM2:
			return fmm2_launch(ctx, pos, acc, n, param, false);
		case NBCO_EVAL_COULOMB_DIRECT2:
			NBCO_TRY(direct2_launch(ctx, pos, acc, n, param));
			return add_elastic2_launch(ctx, pos, acc, n, param ? param + 2 : nullptr);
		case NBCO_EVAL_COULOMB_FMM2:
			return fmm2_launch(ctx, pos, acc, n, param, true);
		default:
			set_error("unknown 2D evaluator %d", evaluator);
			return NBCO_ERR_INVALID;
	}
}

static int sync2(nbco_ctx *ctx)
{
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	return NBCO_OK;
}

// One step of a scheme.  A kick immediately followed by a drift is one fused pass (same fma per
// element as two step() calls, kernel.cuh:85-104).
static int scheme_step2(nbco_ctx *ctx, int scheme, int evaluator, double *buf, int64_t n, const double *param, double dtd)
{
	double *pos = buf, *vel = buf + 2*n, *acc = buf + 4*n;
	const long double dt = dtd;
	auto K = [&](long double c) { return step2_launch(ctx, vel, acc, (double)c, n); };
	auto D = [&](long double c) { return step2_launch(ctx, pos, vel, (double)c, n); };
	auto KD = [&](long double kc, long double dc) { return kick_drift2_launch(ctx, pos, vel, acc, (double)kc, (double)dc, n); };
	auto F = [&]() { return eval2(ctx, evaluator, pos, acc, n, param); };
	switch (scheme)
	{
		case NBCO_EULER: // integrator.cuh:32-48
			NBCO_TRY(KD(dt, dt)); NBCO_TRY(F());
			return NBCO_OK;
		case NBCO_LEAPFROG: // :68-96
			NBCO_TRY(KD(dt * 0.5L, dt)); NBCO_TRY(F()); NBCO_TRY(K(dt * 0.5L));
			return NBCO_OK;
		case NBCO_FORESTRUTH: // :98-128
		{
			const long double th = 1.3512071919596576340476878089715L;
			NBCO_TRY(D(dt * th / 2)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * th, dt * (1 - th) / 2)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * (1 - 2*th), dt * (1 - th) / 2)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * th, dt * th / 2));
			return NBCO_OK;
		}
		case NBCO_PEFRL: // :130-167
		{
			const long double xi = +0.1786178958448091E+00L, la = -0.2123418310626054E+00L, ch = -0.6626458266981849E-01L;
			NBCO_TRY(D(dt * xi)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * (1 - 2*la) / 2, dt * ch)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * la, dt * (1 - 2*(ch + xi)))); NBCO_TRY(F());
			NBCO_TRY(KD(dt * la, dt * ch)); NBCO_TRY(F());
			NBCO_TRY(KD(dt * (1 - 2*la) / 2, dt * xi));
			return NBCO_OK;
		}
		default:
			set_error("unknown scheme %d", scheme);
			return NBCO_ERR_INVALID;
	}
}

static int stage2(nbco_ctx *ctx, int64_t n, const double *h_param, double **d_buf, double **d_param)
{
	NBCO_TRY(ctx->h_state.reserve(sizeof(double) * 6 * (size_t)n));
	NBCO_TRY(ctx->h_param.reserve(sizeof(double) * 8));
	*d_buf = ctx->h_state.as<double>();
	*d_param = nullptr;
	if (h_param)
	{
		*d_param = ctx->h_param.as<double>();
		NBCO_CUDA(cudaMemcpyAsync(*d_param, h_param, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	}
	return NBCO_OK;
}

} // namespace nbco

using namespace nbco;

#define ENTER2(ctx)                                                        \
	if (!(ctx)) { set_error("null context"); return NBCO_ERR_INVALID; }    \
	NBCO_CUDA(cudaSetDevice((ctx)->cfg.device));

extern "C" {

int nbco_force_direct2(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER2(ctx);
	NBCO_TRY(direct2_launch(ctx, (const double *)d_pos, (double *)d_acc, n, (const double *)d_param));
	return sync2(ctx);
}

int nbco_force_fmm2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER2(ctx);
	NBCO_TRY(fmm2_launch(ctx, (double *)d_pos, (double *)d_acc, n, (const double *)d_param, false));
	return sync2(ctx);
}

int nbco_coulomb_direct2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER2(ctx);
	NBCO_TRY(eval2(ctx, NBCO_EVAL_COULOMB_DIRECT2, (double *)d_pos, (double *)d_acc, n, (const double *)d_param));
	return sync2(ctx);
}

int nbco_coulomb_fmm2(nbco_ctx *ctx, void *d_pos, void *d_acc, int64_t n, const void *d_param)
{
	ENTER2(ctx);
	NBCO_TRY(eval2(ctx, NBCO_EVAL_COULOMB_FMM2, (double *)d_pos, (double *)d_acc, n, (const double *)d_param));
	return sync2(ctx);
}

int nbco_add_elastic2(nbco_ctx *ctx, const void *d_pos, void *d_acc, int64_t n, const void *d_k2)
{
	ENTER2(ctx);
	NBCO_TRY(add_elastic2_launch(ctx, (const double *)d_pos, (double *)d_acc, n, (const double *)d_k2));
	return sync2(ctx);
}

int nbco_step2(nbco_ctx *ctx, void *d_b, const void *d_a, double ds, int64_t n)
{
	ENTER2(ctx);
	NBCO_TRY(step2_launch(ctx, (double *)d_b, (const double *)d_a, ds, n));
	return sync2(ctx);
}

int nbco_compute_force2(nbco_ctx *ctx, int evaluator, void *d_buf, int64_t n, const void *d_param)
{
	ENTER2(ctx);
	double *buf = (double *)d_buf;
	NBCO_TRY(eval2(ctx, evaluator, buf, buf + 4*n, n, (const double *)d_param));
	return sync2(ctx);
}

int nbco_integrate2(nbco_ctx *ctx, int scheme, int evaluator, void *d_buf, int64_t n,
                    const void *d_param, double dt, int64_t nsteps)
{
	ENTER2(ctx);
	for (int64_t s = 0; s < nsteps; ++s)
		NBCO_TRY(scheme_step2(ctx, scheme, evaluator, (double *)d_buf, n, (const double *)d_param, dt));
	return sync2(ctx);
}

int nbco_run_host2(nbco_ctx *ctx, int scheme, int evaluator, double *h_pos_vel, double *h_acc, int64_t n,
                   const double *h_param, double dt, int64_t nsteps)
{
	ENTER2(ctx);
	if (!h_pos_vel || n <= 0) { set_error("bad host buffers"); return NBCO_ERR_INVALID; }
	double *d_buf, *d_param;
	NBCO_TRY(stage2(ctx, n, h_param, &d_buf, &d_param));
	const size_t vb = sizeof(double) * 2 * (size_t)n;
	NBCO_CUDA(cudaMemcpyAsync(d_buf, h_pos_vel, 2 * vb, cudaMemcpyHostToDevice, ctx->stream));
	NBCO_TRY(eval2(ctx, evaluator, d_buf, d_buf + 4*n, n, d_param)); // main.cu:863-867
	for (int64_t s = 0; s < nsteps; ++s)
		NBCO_TRY(scheme_step2(ctx, scheme, evaluator, d_buf, n, d_param, dt));
	NBCO_CUDA(cudaMemcpyAsync(h_pos_vel, d_buf, 2 * vb, cudaMemcpyDeviceToHost, ctx->stream));
	if (h_acc) NBCO_CUDA(cudaMemcpyAsync(h_acc, d_buf + 4*n, vb, cudaMemcpyDeviceToHost, ctx->stream));
	return sync2(ctx);
}

int nbco_step_host2(nbco_ctx *ctx, int scheme, int evaluator, double *h_buf, int64_t n,
                    const double *h_param, double dt, int64_t nsteps)
{
	ENTER2(ctx);
	if (!h_buf || n <= 0) { set_error("bad host buffers"); return NBCO_ERR_INVALID; }
	double *d_buf, *d_param;
	NBCO_TRY(stage2(ctx, n, h_param, &d_buf, &d_param));
	const size_t vb = sizeof(double) * 2 * (size_t)n;
	NBCO_CUDA(cudaMemcpyAsync(d_buf, h_buf, 3 * vb, cudaMemcpyHostToDevice, ctx->stream));
	for (int64_t s = 0; s < nsteps; ++s)
		NBCO_TRY(scheme_step2(ctx, scheme, evaluator, d_buf, n, d_param, dt));
	NBCO_CUDA(cudaMemcpyAsync(h_buf, d_buf, 3 * vb, cudaMemcpyDeviceToHost, ctx->stream));
	return sync2(ctx);
}

} // extern "C"
