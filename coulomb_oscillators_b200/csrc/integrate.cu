// integrate.cu -- streaming kernels of the time-stepping path for sm_100a.
//
// Replaces step/step_krnl and add_elastic/add_elastic_krnl (reference Simulation/kernel.cuh:85-152),
// which the reference launches with at most MAX_GRID_SIZE = 10 blocks of 128 threads
// (constants.cuh:36-37): 1280 threads for 16M particles.  Here every pass is a plain HBM-bound
// stream sized to the machine: grid = SMs x resident CTAs, 128-bit accesses over the flat float
// array (3n floats; float3 AoS has no per-particle structure these kernels need), grid-stride.
// Arithmetic is one fma per element exactly like the reference device code (fma(ds, a, b),
// kernel.cuh:94; fma(-k, x, a), kernel.cuh:130).
// Also: the mean-relative-error metric (rel_diff1, reductions.cuh:37-42) and the kinetic /
// elastic energy sums (the reference has no energy code).

#include "common.cuh"

namespace nbco {

namespace {

constexpr int kBlock = 256;

// b[i] += a[i] * ds over m floats
__global__ void __launch_bounds__(kBlock) axpy_kernel(float *__restrict__ b, const float *__restrict__ a, float ds, int64_t m)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool aligned = ((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(a)) & 15) == 0;
	if (aligned)
	{
		const int64_t m4 = m >> 2;
		float4 *b4 = reinterpret_cast<float4 *>(b);
		const float4 *a4 = reinterpret_cast<const float4 *>(a);
		for (int64_t i = t; i < m4; i += stride)
		{
			float4 x = a4[i], y = b4[i];
			y.x = fmaf(ds, x.x, y.x); y.y = fmaf(ds, x.y, y.y);
			y.z = fmaf(ds, x.z, y.z); y.w = fmaf(ds, x.w, y.w);
			b4[i] = y;
		}
		for (int64_t i = (m4 << 2) + t; i < m; i += stride)
			b[i] = fmaf(ds, a[i], b[i]);
	}
	else
		for (int64_t i = t; i < m; i += stride)
			b[i] = fmaf(ds, a[i], b[i]);
}

// fused kick(s) + drift over m floats: v = fma(k1, a, v); [v = fma(k2, a, v);] x = fma(dt, v, x).
// Element for element the same fma sequence as separate step() launches (kernel.cuh:85-98), in one pass:
// 60 B/particle instead of 72 (kick + drift) or 108 (kick + kick + drift).
template <bool TWO>
__global__ void __launch_bounds__(kBlock)
kick_drift_kernel(float *__restrict__ x, float *__restrict__ v, const float *__restrict__ a, float k1, float k2, float dt, int64_t m)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(a)) & 15) == 0;
	const int64_t m4 = aligned ? (m >> 2) : 0;
	float4 *x4 = reinterpret_cast<float4 *>(x), *v4 = reinterpret_cast<float4 *>(v);
	const float4 *a4 = reinterpret_cast<const float4 *>(a);
	for (int64_t i = t; i < m4; i += stride)
	{
		float4 aa = a4[i], vv = v4[i], xx = x4[i];
		vv.x = fmaf(k1, aa.x, vv.x); vv.y = fmaf(k1, aa.y, vv.y); vv.z = fmaf(k1, aa.z, vv.z); vv.w = fmaf(k1, aa.w, vv.w);
		if (TWO) { vv.x = fmaf(k2, aa.x, vv.x); vv.y = fmaf(k2, aa.y, vv.y); vv.z = fmaf(k2, aa.z, vv.z); vv.w = fmaf(k2, aa.w, vv.w); }
		xx.x = fmaf(dt, vv.x, xx.x); xx.y = fmaf(dt, vv.y, xx.y); xx.z = fmaf(dt, vv.z, xx.z); xx.w = fmaf(dt, vv.w, xx.w);
		v4[i] = vv; x4[i] = xx;
	}
	for (int64_t i = (m4 << 2) + t; i < m; i += stride)
	{
		float vv = fmaf(k1, a[i], v[i]);
		if (TWO) vv = fmaf(k2, a[i], vv);
		v[i] = vv;
		x[i] = fmaf(dt, vv, x[i]);
	}
}

// a[3i+c] -= k[c] * x[3i+c]
__global__ void __launch_bounds__(kBlock) elastic_kernel(const float *__restrict__ x, float *__restrict__ a,
                                                        const float *__restrict__ k3, int64_t m)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	float k[3] = {1.f, 1.f, 1.f};
	if (k3) { k[0] = k3[0]; k[1] = k3[1]; k[2] = k3[2]; }
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
		a[i] = fmaf(-k[i % 3], x[i], a[i]);
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) sh[w] = v;
	__syncthreads();
	double r = 0.0;
	if (w == 0)
	{
		r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
		for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
	}
	__syncthreads();
	return r;
}
__device__ __forceinline__ double block_max(double v, double *sh)
{
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
	int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) sh[w] = v;
	__syncthreads();
	double r = 0.0;
	if (w == 0)
	{
		r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
		for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
	}
	__syncthreads();
	return r;
}

// out[0] += sum rel_diff1, out[1] = max rel_diff1 (as a non-negative double -> ordered as uint64)
__global__ void __launch_bounds__(kBlock) rel_err_kernel(const float *__restrict__ a, const float *__restrict__ ref,
                                                        int64_t n, double *__restrict__ out)
{
	__shared__ double sh[32];
	double s = 0.0, mx = 0.0;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
	{
		float dx = a[3*i] - ref[3*i], dy = a[3*i+1] - ref[3*i+1], dz = a[3*i+2] - ref[3*i+2];
		float d2 = dx*dx + dy*dy + dz*dz;
		float r2 = ref[3*i]*ref[3*i] + ref[3*i+1]*ref[3*i+1] + ref[3*i+2]*ref[3*i+2] + 1.e-18f;
		float e = sqrtf(fmaxf(d2 / r2, 0.f));
		s += (double)e;
		mx = fmax(mx, (double)e);
	}
	s = block_sum(s, sh);
	mx = block_max(mx, sh);
	if (threadIdx.x == 0)
	{
		atomicAdd(out, s);
		atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)__double_as_longlong(mx));
	}
}

// out[0] += sum 1/2 v^2 ; out[1] += 1/2 sum k o x^2 ; buf = [pos|vel|acc]
__global__ void __launch_bounds__(kBlock) kin_el_kernel(const float *__restrict__ buf, int64_t n, const float *__restrict__ param,
                                                       double *__restrict__ out)
{
	__shared__ double sh[32];
	double ke = 0.0, el = 0.0;
	float k[3] = {1.f, 1.f, 1.f};
	if (param) { k[0] = param[3]; k[1] = param[4]; k[2] = param[5]; }
	const float *x = buf, *v = buf + 3*n;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
	{
		double vx = v[3*i], vy = v[3*i+1], vz = v[3*i+2];
		double px = x[3*i], py = x[3*i+1], pz = x[3*i+2];
		ke += 0.5 * (vx*vx + vy*vy + vz*vz);
		el += 0.5 * ((double)k[0]*px*px + (double)k[1]*py*py + (double)k[2]*pz*pz);
	}
	ke = block_sum(ke, sh);
	el = block_sum(el, sh);
	if (threadIdx.x == 0) { atomicAdd(out, ke); atomicAdd(out + 1, el); }
}

// The LAST update of an integration fused with the energy reduction (north star: "fused integrator kick/drift plus the
// energy reduction"): b = fma(ds, a, b) exactly like axpy_kernel, and in the same pass out[0] += sum 1/2 v^2,
// out[1] += 1/2 sum k o x^2 of the state just produced.  b_is_vel: the update is a kick (b = velocities, a =
// accelerations, other = positions); else a drift (b = positions, a = velocities; other unused).
__global__ void __launch_bounds__(kBlock)
axpy_energy_kernel(float *__restrict__ b, const float *__restrict__ a, float ds, const float *__restrict__ other, int b_is_vel,
                   const float *__restrict__ param, int64_t m, double *__restrict__ out)
{
	__shared__ double sh[32];
	double ke = 0.0, el = 0.0;
	float k[3] = {1.f, 1.f, 1.f};
	if (param) { k[0] = param[3]; k[1] = param[4]; k[2] = param[5]; }
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
	{
		const float nb = fmaf(ds, a[i], b[i]);
		b[i] = nb;
		const double v = b_is_vel ? (double)nb : (double)a[i], x = b_is_vel ? (double)other[i] : (double)nb;
		ke += 0.5 * v * v;
		el += 0.5 * (double)k[i % 3] * x * x;
	}
	ke = block_sum(ke, sh);
	el = block_sum(el, sh);
	if (threadIdx.x == 0) { atomicAdd(out, ke); atomicAdd(out + 1, el); }
}

} // namespace

int step_energy_launch(nbco_ctx *ctx, float *d_b, const float *d_a, float ds, const float *d_other, bool b_is_vel,
                       const float *d_param, int64_t n, double *d_out2)
{
	if (n <= 0) return NBCO_OK;
	axpy_energy_kernel<<<grid_for(3 * n, kBlock, ctx->sm_count, 8), kBlock, 0, ctx->stream>>>(d_b, d_a, ds, d_other, b_is_vel ? 1 : 0, d_param, 3 * n, d_out2);
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int step_launch(nbco_ctx *ctx, float *d_b, const float *d_a, float ds, int64_t n)
{
	if (n <= 0) return NBCO_OK;
	axpy_kernel<<<grid_for((3*n + 3) / 4, kBlock, ctx->sm_count, 8), kBlock, 0, ctx->stream>>>(d_b, d_a, ds, 3*n);
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int kick_drift_launch(nbco_ctx *ctx, float *d_pos, float *d_vel, const float *d_acc, float k1, float k2, bool two, float dt, int64_t n)
{
	if (n <= 0) return NBCO_OK;
	const int grid = grid_for((3*n + 3) / 4, kBlock, ctx->sm_count, 8);
	if (two) kick_drift_kernel<true><<<grid, kBlock, 0, ctx->stream>>>(d_pos, d_vel, d_acc, k1, k2, dt, 3*n);
	else kick_drift_kernel<false><<<grid, kBlock, 0, ctx->stream>>>(d_pos, d_vel, d_acc, k1, k2, dt, 3*n);
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int add_elastic_launch(nbco_ctx *ctx, const float *d_pos, float *d_acc, int64_t n, const float *d_k3)
{
	if (n <= 0) return NBCO_OK;
	elastic_kernel<<<grid_for(3*n, kBlock, ctx->sm_count, 8), kBlock, 0, ctx->stream>>>(d_pos, d_acc, d_k3, 3*n);
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int rel_err_launch(nbco_ctx *ctx, const float *d_a, const float *d_ref, int64_t n, double *h_mean, double *h_max)
{
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>();
	NBCO_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(double), ctx->stream));
	if (n > 0)
	{
		rel_err_kernel<<<grid_for(n, kBlock, ctx->sm_count, 4), kBlock, 0, ctx->stream>>>(d_a, d_ref, n, d);
		++ctx->launches;
	}
	double h[2];
	NBCO_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (h_mean) *h_mean = n > 0 ? h[0] / (double)n : 0.0;
	if (h_max) *h_max = h[1];
	return NBCO_OK;
}

int kinetic_elastic_launch(nbco_ctx *ctx, const float *d_buf, int64_t n, const float *d_param, double *h_out2)
{
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>();
	NBCO_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(double), ctx->stream));
	if (n > 0)
	{
		kin_el_kernel<<<grid_for(n, kBlock, ctx->sm_count, 4), kBlock, 0, ctx->stream>>>(d_buf, n, d_param, d);
		++ctx->launches;
	}
	NBCO_CUDA(cudaMemcpyAsync(h_out2, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	return NBCO_OK;
}

} // namespace nbco
