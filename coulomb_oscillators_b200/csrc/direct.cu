// direct.cu -- O(N^2) softened-Coulomb direct sum for sm_100a.
//
// Replaces direct3 / direct3_krnl (reference Simulation/direct.cuh:192-245): same sum
//     a_i = param[0] * sum_j d (|d|^2 + eps2)^(-3/2),  d = x_i - x_j   (i = j contributes exactly 0)
// but organised for the B200 FP32 pipe instead of one-thread-per-i over global memory:
//   * sources are staged once per call as float4 (16-B aligned; the reference's float3 AoS cannot
//     be loaded with one instruction) and streamed through shared memory in tiles; every thread
//     of a warp reads the same source -> one broadcast LDS.128 per source per warp;
//   * each thread keeps IPT targets in registers, so one LDS feeds IPT*32 interactions;
//   * per interaction: 3 FADD + 3 FFMA + MUFU.RSQ + 2 FMUL + 3 FFMA (11 FP32-pipe ops + 1 SFU),
//     the count SURVEY.md section 8(d) uses for the FMA roofline.  rsqrt.approx.ftz replaces the
//     reference's IEEE 1/x and sqrt (<= 2 ulp per term);
//   * optional packed-FP32 path (add/mul/fma.f32x2, new on sm_100): two targets per instruction,
//     halving the issue slots taken by the FP32 pipe;
//   * accuracy: fp32 partial sums inside a tile of sources, Kahan-compensated accumulation
//     across tiles (the reference compensates every term, direct.cuh:204-222; compensating per
//     tile keeps the error at the 1e-7 level for 11 instead of 23 ops per interaction).
// Multi-GPU: targets are sharded by rank (nbco_shard_range), sources are the full set.

#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace nbco {

namespace {

constexpr int kBlock = 128;   // threads per CTA
constexpr int kTileJ = 1024;  // sources per shared-memory tile (16 KB)

__device__ __forceinline__ float rsqrt_approx(float x)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// ---- packed fp32x2 helpers (sm_100+: one instruction, two lanes of a 64-bit register) ----
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
	f32x2 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
	f32x2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}

__global__ void __launch_bounds__(256) to_float4_kernel(const float *__restrict__ pos, float4 *__restrict__ out, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride)
		out[i] = make_float4(pos[3*i], pos[3*i+1], pos[3*i+2], 0.f);
}

struct Kahan3
{
	float sx = 0.f, sy = 0.f, sz = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
	__device__ __forceinline__ void add(float x, float y, float z)
	{
		float yx = x - cx, tx = sx + yx; cx = (tx - sx) - yx; sx = tx;
		float yy = y - cy, ty = sy + yy; cy = (ty - sy) - yy; sy = ty;
		float yz = z - cz, tz = sz + yz; cz = (tz - sz) - yz; sz = tz;
	}
};

// Scalar FP32 variant.
template <int IPT>
__global__ void __launch_bounds__(kBlock)
direct3_scalar_kernel(const float4 *__restrict__ src, int64_t n_src, int64_t i_begin, int64_t i_end,
                      float *__restrict__ acc, const float *__restrict__ param, float eps2)
{
	__shared__ float4 tile[kTileJ];
	const int tid = threadIdx.x;
	const int64_t i0 = i_begin + (int64_t)blockIdx.x * (kBlock * IPT) + tid;

	float xi[IPT], yi[IPT], zi[IPT];
	Kahan3 sum[IPT];
#pragma unroll
	for (int k = 0; k < IPT; ++k)
	{
		int64_t i = i0 + (int64_t)k * kBlock;
		float4 q = src[i < i_end ? i : i_begin];
		xi[k] = q.x; yi[k] = q.y; zi[k] = q.z;
	}

	for (int64_t base = 0; base < n_src; base += kTileJ)
	{
		int cnt = (int)((n_src - base < kTileJ) ? (n_src - base) : kTileJ);
		__syncthreads();
		for (int j = tid; j < cnt; j += kBlock)
			tile[j] = src[base + j];
		__syncthreads();

		float ax[IPT], ay[IPT], az[IPT];
#pragma unroll
		for (int k = 0; k < IPT; ++k) { ax[k] = 0.f; ay[k] = 0.f; az[k] = 0.f; }

#pragma unroll 4
		for (int j = 0; j < cnt; ++j)
		{
			float4 s = tile[j];
#pragma unroll
			for (int k = 0; k < IPT; ++k)
			{
				float dx = xi[k] - s.x, dy = yi[k] - s.y, dz = zi[k] - s.z;
				float r2 = fmaf(dx, dx, eps2);
				r2 = fmaf(dy, dy, r2);
				r2 = fmaf(dz, dz, r2);
				float w = rsqrt_approx(r2);
				float w3 = (w * w) * w;
				ax[k] = fmaf(dx, w3, ax[k]);
				ay[k] = fmaf(dy, w3, ay[k]);
				az[k] = fmaf(dz, w3, az[k]);
			}
		}
#pragma unroll
		for (int k = 0; k < IPT; ++k)
			sum[k].add(ax[k], ay[k], az[k]);
	}

	const float scale = param ? param[0] : 1.f;
#pragma unroll
	for (int k = 0; k < IPT; ++k)
	{
		int64_t i = i0 + (int64_t)k * kBlock;
		if (i < i_end)
		{
			acc[3*i]   = scale * sum[k].sx;
			acc[3*i+1] = scale * sum[k].sy;
			acc[3*i+2] = scale * sum[k].sz;
		}
	}
}

// Packed FP32x2 variant: targets are processed in pairs (IPT even).
template <int IPT, int BLOCK = kBlock, int UNROLL = 4>
__global__ void __launch_bounds__(BLOCK)
direct3_packed_kernel(const float4 *__restrict__ src, int64_t n_src, int64_t i_begin, int64_t i_end,
                      float *__restrict__ acc, const float *__restrict__ param, float eps2, float *__restrict__ part = nullptr)
{
	static_assert(IPT % 2 == 0, "packed variant handles targets in pairs");
	constexpr int NP = IPT / 2;
	// sources are stored pre-duplicated (x,x,y,y | z,z,-,-) so that no per-source MOVs are needed
	__shared__ float4 tile_xy[kTileJ];
	__shared__ float2 tile_z[kTileJ];
	const int tid = threadIdx.x;
	const int64_t i0 = i_begin + (int64_t)blockIdx.x * (BLOCK * IPT) + tid;

	f32x2 xi[NP], yi[NP], zi[NP];
	Kahan3 sum[IPT];
#pragma unroll
	for (int k = 0; k < NP; ++k)
	{
		int64_t ia = i0 + (int64_t)(2*k) * BLOCK, ib = ia + BLOCK;
		float4 qa = src[ia < i_end ? ia : i_begin];
		float4 qb = src[ib < i_end ? ib : i_begin];
		xi[k] = pack2(qa.x, qb.x); yi[k] = pack2(qa.y, qb.y); zi[k] = pack2(qa.z, qb.z);
	}
	const f32x2 eps2p = pack2(eps2, eps2);

	// split-j: blockIdx.y takes the y-th of gridDim.y equal runs of source tiles (fills the last wave of the grid when
	// the target blocks alone do not divide the SM slots; partial sums are then combined with atomicAdd)
	const int64_t tiles = (n_src + kTileJ - 1) / kTileJ;
	const int64_t j_lo = (tiles * blockIdx.y / gridDim.y) * kTileJ;
	const int64_t j_hi_ = (tiles * (blockIdx.y + 1) / gridDim.y) * kTileJ;
	const int64_t j_hi = j_hi_ < n_src ? j_hi_ : n_src;
	for (int64_t base = j_lo; base < j_hi; base += kTileJ)
	{
		int cnt = (int)((j_hi - base < kTileJ) ? (j_hi - base) : kTileJ);
		__syncthreads();
		for (int j = tid; j < cnt; j += BLOCK)
		{
			float4 s = src[base + j];
			tile_xy[j] = make_float4(s.x, s.x, s.y, s.y);
			tile_z[j] = make_float2(s.z, s.z);
		}
		__syncthreads();

		f32x2 ax[NP], ay[NP], az[NP];
#pragma unroll
		for (int k = 0; k < NP; ++k) { ax[k] = 0ull; ay[k] = 0ull; az[k] = 0ull; }

#pragma unroll UNROLL
		for (int j = 0; j < cnt; ++j)
		{
			const ulonglong2 sxy = *reinterpret_cast<const ulonglong2 *>(&tile_xy[j]);
			const f32x2 sz = *reinterpret_cast<const f32x2 *>(&tile_z[j]);
#pragma unroll
			for (int k = 0; k < NP; ++k)
			{
				f32x2 dx = sub2(xi[k], sxy.x), dy = sub2(yi[k], sxy.y), dz = sub2(zi[k], sz);
				f32x2 r2 = fma2(dx, dx, eps2p);
				r2 = fma2(dy, dy, r2);
				r2 = fma2(dz, dz, r2);
				float r2a, r2b;
				unpack2(r2, r2a, r2b);
				f32x2 w = pack2(rsqrt_approx(r2a), rsqrt_approx(r2b));
				f32x2 w3 = mul2(mul2(w, w), w);
				ax[k] = fma2(dx, w3, ax[k]);
				ay[k] = fma2(dy, w3, ay[k]);
				az[k] = fma2(dz, w3, az[k]);
			}
		}
#pragma unroll
		for (int k = 0; k < NP; ++k)
		{
			float xa, xb, ya, yb, za, zb;
			unpack2(ax[k], xa, xb); unpack2(ay[k], ya, yb); unpack2(az[k], za, zb);
			sum[2*k].add(xa, ya, za);
			sum[2*k+1].add(xb, yb, zb);
		}
	}

	const float scale = param ? param[0] : 1.f;
#pragma unroll
	for (int k = 0; k < IPT; ++k)
	{
		int64_t i = i0 + (int64_t)k * BLOCK;
		if (i < i_end)
		{
			if (gridDim.y == 1)
			{
				acc[3*i]   = scale * sum[k].sx;
				acc[3*i+1] = scale * sum[k].sy;
				acc[3*i+2] = scale * sum[k].sz;
			}
			else
			{
				// partial sum of this run of source tiles; combined in a fixed order by direct3_combine_kernel
				float *o = part + 3 * ((int64_t)blockIdx.y * (i_end - i_begin) + (i - i_begin));
				o[0] = sum[k].sx; o[1] = sum[k].sy; o[2] = sum[k].sz;
			}
		}
	}
}

// acc[i] = scale * (partial sums of the source runs, added in run order with Kahan compensation): deterministic, and the
// same for every sharding of the targets (the number of runs depends on the number of SOURCES only)
__global__ void __launch_bounds__(256)
direct3_combine_kernel(const float *__restrict__ part, int parts, int64_t i_begin, int64_t i_end, float *__restrict__ acc, const float *__restrict__ param)
{
	const int64_t cnt = i_end - i_begin;
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= cnt) return;
	const float scale = param ? param[0] : 1.f;
	Kahan3 s;
	for (int y = 0; y < parts; ++y)
	{
		const float *o = part + 3 * ((int64_t)y * cnt + t);
		s.add(o[0], o[1], o[2]);
	}
	float *a = acc + 3 * (i_begin + t);
	a[0] = scale * s.sx; a[1] = scale * s.sy; a[2] = scale * s.sz;
}

// Pair potential sum_{i<j} (d^2+eps2)^(-1/2), accumulated in double (diagnostic, not a hot path).
__global__ void __launch_bounds__(kBlock)
pair_energy_kernel(const float4 *__restrict__ src, int64_t n, int64_t i_begin, int64_t i_end, float eps2, double *__restrict__ out)
{
	__shared__ float4 tile[kTileJ];
	__shared__ double red[kBlock];
	const int tid = threadIdx.x;
	const int64_t i = i_begin + (int64_t)blockIdx.x * kBlock + tid;
	float4 q = src[i < i_end ? i : i_begin];
	double phi = 0.0;
	for (int64_t base = 0; base < n; base += kTileJ)
	{
		int cnt = (int)((n - base < kTileJ) ? (n - base) : kTileJ);
		__syncthreads();
		for (int j = tid; j < cnt; j += kBlock)
			tile[j] = src[base + j];
		__syncthreads();
		float part = 0.f;
		for (int j = 0; j < cnt; ++j)
		{
			float4 s = tile[j];
			float dx = q.x - s.x, dy = q.y - s.y, dz = q.z - s.z;
			float r2 = fmaf(dx, dx, eps2);
			r2 = fmaf(dy, dy, r2);
			r2 = fmaf(dz, dz, r2);
			float w = rsqrtf(r2);
			part += (base + j != i) ? w : 0.f;
		}
		phi += (double)part;
	}
	red[tid] = (i < i_end) ? phi : 0.0;
	__syncthreads();
	for (int s = kBlock / 2; s > 0; s >>= 1)
	{
		if (tid < s) red[tid] += red[tid + s];
		__syncthreads();
	}
	if (tid == 0) atomicAdd(out, 0.5 * red[0]);
}

int g_direct_variant = -1; // -1: default; set through NBCO_DIRECT_VARIANT for experiments

} // namespace

int direct3_launch(nbco_ctx *ctx, const float *d_pos, float *d_acc, int64_t n, const float *d_param)
{
	if (n <= 0) return NBCO_OK;
	NBCO_TRY(ctx->pos4.reserve(sizeof(float4) * (size_t)n));
	float4 *src = ctx->pos4.as<float4>();
	to_float4_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_pos, src, n);
	++ctx->launches;

	int64_t ib, ie;
	nbco_shard_range(n, ctx->cfg.rank, ctx->cfg.world, &ib, &ie);
	if (ie <= ib) return NBCO_OK;

	if (g_direct_variant < 0)
	{
		const char *e = getenv("NBCO_DIRECT_VARIANT");
		g_direct_variant = e ? atoi(e) : 0;
	}
	const int64_t cnt = ie - ib;
	const float eps2 = ctx->cfg.eps2;
#define LAUNCH(KERNEL, IPT)                                                              \
	do {                                                                                 \
		int64_t blocks = (cnt + (int64_t)kBlock * IPT - 1) / ((int64_t)kBlock * IPT);    \
		KERNEL<IPT><<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(src, n, ib, ie, d_acc, d_param, eps2); \
	} while (0)
	switch (g_direct_variant)
	{
		case 1: LAUNCH(direct3_scalar_kernel, 2); break;
		case 2: LAUNCH(direct3_scalar_kernel, 8); break;
		case 3: LAUNCH(direct3_scalar_kernel, 4); break;
		case 4: LAUNCH(direct3_packed_kernel, 8); break;
		case 5: LAUNCH(direct3_packed_kernel, 2); break;
		case 6:
		{
			int64_t blocks = (cnt + 256 * 4 - 1) / (256 * 4);
			direct3_packed_kernel<4, 256><<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, n, ib, ie, d_acc, d_param, eps2);
			break;
		}
		case 7: LAUNCH(direct3_packed_kernel, 6); break;
#define LAUNCH_P(IPT, BLK, UNR)                                                                          \
		{                                                                                                \
			int64_t blocks = (cnt + (int64_t)BLK * IPT - 1) / ((int64_t)BLK * IPT);                      \
			direct3_packed_kernel<IPT, BLK, UNR><<<(unsigned)blocks, BLK, 0, ctx->stream>>>(src, n, ib, ie, d_acc, d_param, eps2); \
			break;                                                                                       \
		}
		case 8: LAUNCH_P(4, 256, 8)
		case 9: LAUNCH_P(4, 512, 4)
		case 10: LAUNCH_P(4, 256, 2)
		case 11: LAUNCH_P(2, 256, 8)
		case 12: LAUNCH_P(4, 128, 8)
		case 13: LAUNCH(direct3_packed_kernel, 4); break;
		default: // <4, 256, 4>: measured fastest on B200 (profiles/r01_notes.md)
		{
			// grid: target blocks x source runs.  The target blocks alone can leave most of the chip idle (a rank of 8 at
			// N = 2^20: 128 blocks on 148 SMs; N = 8192: 8 blocks), so the source tiles are cut into `parts` runs, a
			// function of the number of SOURCES only: every sharding of the targets adds the same partial sums in the same
			// order (bit-identical shards, tests/test_direct_gpu.py), and there are no atomics.
			const int64_t bx = (cnt + 256 * 4 - 1) / (256 * 4);
			const int64_t tiles = (n + kTileJ - 1) / kTileJ;
			const int parts = (int)std::min<int64_t>(16, tiles);
			float *part = nullptr;
			if (parts > 1)
			{
				NBCO_TRY(ctx->dpart.reserve(sizeof(float) * 3 * (size_t)parts * (size_t)cnt));
				part = ctx->dpart.as<float>();
			}
			direct3_packed_kernel<4, 256, 4><<<dim3((unsigned)bx, (unsigned)parts), 256, 0, ctx->stream>>>(src, n, ib, ie, d_acc, d_param, eps2, part);
			if (parts > 1)
			{
				direct3_combine_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(part, parts, ib, ie, d_acc, d_param);
				++ctx->launches;
			}
			break;
		}
#undef LAUNCH_P
	}
#undef LAUNCH
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int pair_energy_launch(nbco_ctx *ctx, const float *d_pos, int64_t n, double *d_out)
{
	NBCO_TRY(ctx->pos4.reserve(sizeof(float4) * (size_t)n));
	float4 *src = ctx->pos4.as<float4>();
	to_float4_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_pos, src, n);
	int64_t ib, ie;
	nbco_shard_range(n, ctx->cfg.rank, ctx->cfg.world, &ib, &ie);
	NBCO_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), ctx->stream));
	if (ie > ib)
	{
		int64_t blocks = (ie - ib + kBlock - 1) / kBlock;
		pair_energy_kernel<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(src, n, ib, ie, ctx->cfg.eps2, d_out);
	}
	ctx->launches += 2;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

} // namespace nbco
