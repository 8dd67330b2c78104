// fmm2.cu -- the 2D fp64 path for sm_100a: uniform-quadtree Cartesian FMM, direct sum, step kernels.
//
// Replaces (reference paths relative to Simulation/):
//   fmm_cart                     fmm_cart.cuh:395-545   (GPU driver; CPU twin :546-680)
//   evalKeys / indexLeaves / multLeaves / centerLeaves / p2p2   appel.cuh:44-55,141-212,226-243,260-318
//   fmm_multipoleLeaves / fmm_buildTree2 / fmm_c2c2 / fmm_pushl / fmm_pushLeaves   fmm_cart.cuh:68-378
//   operator algebra             fmm_cart_base.cuh (static_p2m/m2m/m2l/l2l/l2p in Cartesian tensors)
//   cub::DeviceRadixSort + gather_krnl/copy_krnl            fmm_cart.cuh:479-505
//   direct2                      direct.cuh:140-190
//   step / add_elastic / rescale kernel.cuh:85-152, appel.cuh:506-527
//
// Same algorithm (same grid, same levels L, same interaction lists, same expansion order, same
// summation order inside a cell run), different formulation and kernels:
//   * 2D harmonic tensors are complex numbers.  A symmetric order-q multipole enters every
//     contraction with the traceless gradient only through Z_q = sum_k binom(q,k) i^k M_q[k]
//     = sum_j (-(z_j - c))^q / q!, and a traceless order-n local is L_n = L_n[0] + i L_n[1]; so
//        P2M  Z_q  = sum_j (-(z_j - c))^q / q!
//        M2M  Z'_n = sum_m Z_(n-m) d^m / m!,                  d = c_parent - c_child
//        M2L  L_n += 1/n! sum_q conj(Z_q) g_(n+q),            g_m = (-1)^m (m-1)! r^-m (T_m(d0) + i d1 U_(m-1)(d0)),
//                                                             r^2 = |dz|^2 + eps2, d = dz / r (the reference's softened form)
//        L2L  L'_q = sum_(m>=q) binom(m,q) L_m conj(d)^(m-q), d = c_child - c_parent
//        L2P  f    = -sum_n n L_n conj(d)^(n-1)
//     (tests/cx2d.py states these in numpy and tests/test_ops2d_host.py pins them to the oracle's
//     Cartesian form).  Per node that is 2(p+1) doubles instead of (p+1)(p+2)/2 + 2p+1, and an M2L
//     is (p+1)^2 complex FMAs instead of the O(p^3) tensor contraction;
//   * stable LSD radix sort of the 2L-bit cell keys written here (warp match ranking), one gather;
//   * M2L and L2L of a level fused into one kernel (thread per target node, no atomics);
//   * near field: thread per sorted particle, sources of the 2r+1 row runs read through L1
//     (lanes of one cell broadcast), fp64 reciprocal = MUFU.RCP64H + 6 DFMA, with L2P, the
//     xi/N rescale and the elastic term fused into the same kernel (one write of acc).
// There is no CPU path.

#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace nbco {

namespace {

typedef unsigned int u32;

constexpr int kB = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kB * kRsItems; // keys per radix tile
constexpr int kRsBits = 9, kRsBins = 1 << kRsBits; // digit width of one pass
constexpr int kMaxP2 = NBCO2_MAX_ORDER;

struct Grid2 { double minx, miny, rdelta, pad; };

__host__ __device__ __forceinline__ int tbeg(int l) { return (int)((((int64_t)1 << (2 * l)) - 1) / 3); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
	return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void cfma(double2 &s, double2 a, double2 b) // s += a*b
{
	s.x = fma(a.x, b.x, s.x); s.x = fma(-a.y, b.y, s.x);
	s.y = fma(a.x, b.y, s.y); s.y = fma(a.y, b.x, s.y);
}
__device__ __forceinline__ void cfma_conj(double2 &s, double2 a, double2 b) // s += conj(a)*b
{
	s.x = fma(a.x, b.x, s.x); s.x = fma(a.y, b.y, s.x);
	s.y = fma(a.x, b.y, s.y); s.y = fma(-a.y, b.x, s.y);
}

// 1/x for normal positive x: MUFU.RCP64H seed (measured: only ~2^-10 accurate) refined by two
// third-order steps y <- y (1 + e + e^2), e = 1 - x y: relative error e^9, i.e. below 1 ulp
__device__ __forceinline__ double rcp_nr(double x)
{
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double e = fma(-x, y, 1.0);
	y = fma(y, fma(e, e, e), y);
	e = fma(-x, y, 1.0);
	y = fma(y, fma(e, e, e), y);
	return y;
}

// Four softened pair terms with ONE reciprocal (Montgomery's trick): 1/r_k = (prod of the others) / (r0 r1 r2 r3).
// 9 DMUL + 6 DFMA + 1 MUFU for four inverses instead of 4 x (6 DFMA + MUFU); products stay far inside the
// double range (r_k >= eps2).  Accumulation order is the source order, like the reference's loop.
__device__ __forceinline__ void pair4(double px, double py, double2 q0, double2 q1, double2 q2, double2 q3, double eps2,
                                      double &ax, double &ay)
{
	const double dx0 = px - q0.x, dy0 = py - q0.y, dx1 = px - q1.x, dy1 = py - q1.y;
	const double dx2 = px - q2.x, dy2 = py - q2.y, dx3 = px - q3.x, dy3 = py - q3.y;
	const double r0 = fma(dy0, dy0, fma(dx0, dx0, eps2)), r1 = fma(dy1, dy1, fma(dx1, dx1, eps2));
	const double r2 = fma(dy2, dy2, fma(dx2, dx2, eps2)), r3 = fma(dy3, dy3, fma(dx3, dx3, eps2));
	const double p01 = r0 * r1, p23 = r2 * r3;
	const double I = rcp_nr(p01 * p23);
	const double i01 = I * p23, i23 = I * p01;
	const double v0 = i01 * r1, v1 = i01 * r0, v2 = i23 * r3, v3 = i23 * r2;
	ax = fma(v0, dx0, ax); ay = fma(v0, dy0, ay);
	ax = fma(v1, dx1, ax); ay = fma(v1, dy1, ay);
	ax = fma(v2, dx2, ax); ay = fma(v2, dy2, ay);
	ax = fma(v3, dx3, ax); ay = fma(v3, dy3, ay);
}
__device__ __forceinline__ void pair1(double px, double py, double2 q, double eps2, double &ax, double &ay)
{
	const double dx = px - q.x, dy = py - q.y;
	const double inv = rcp_nr(fma(dy, dy, fma(dx, dx, eps2)));
	ax = fma(inv, dx, ax); ay = fma(inv, dy, ay);
}

__host__ __device__ constexpr double binom_c(int n, int k)
{
	double r = 1.0;
	for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
	return r;
}
__host__ __device__ constexpr double fact_c(int n)
{
	double r = 1.0;
	for (int i = 2; i <= n; ++i) r *= (double)i;
	return r;
}

// ---------------------------------------------------------------------------------------------
//  bounding box (minmaxReduce, reductions.cuh:67-80) and grid parameters (fmm_cart.cuh:459-476)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kB) bbox2_kernel(const double2 *__restrict__ p, int64_t n, double4 *__restrict__ part)
{
	double lx = INFINITY, ly = INFINITY, hx = -INFINITY, hy = -INFINITY;
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const double2 v = p[i];
		lx = fmin(lx, v.x); ly = fmin(ly, v.y); hx = fmax(hx, v.x); hy = fmax(hy, v.y);
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		lx = fmin(lx, __shfl_down_sync(0xffffffffu, lx, o)); ly = fmin(ly, __shfl_down_sync(0xffffffffu, ly, o));
		hx = fmax(hx, __shfl_down_sync(0xffffffffu, hx, o)); hy = fmax(hy, __shfl_down_sync(0xffffffffu, hy, o));
	}
	__shared__ double4 sh[kB / 32];
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) sh[w] = make_double4(lx, ly, hx, hy);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double4 r = sh[0];
		for (int k = 1; k < kB / 32; ++k) { r.x = fmin(r.x, sh[k].x); r.y = fmin(r.y, sh[k].y); r.z = fmax(r.z, sh[k].z); r.w = fmax(r.w, sh[k].w); }
		part[blockIdx.x] = r;
	}
}

__global__ void __launch_bounds__(kB) grid2_kernel(const double4 *__restrict__ part, int nb, int side, double eps, Grid2 *__restrict__ g)
{
	double lx = INFINITY, ly = INFINITY, hx = -INFINITY, hy = -INFINITY;
	for (int k = threadIdx.x; k < nb; k += kB)
	{
		const double4 v = part[k];
		lx = fmin(lx, v.x); ly = fmin(ly, v.y); hx = fmax(hx, v.z); hy = fmax(hy, v.w);
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		lx = fmin(lx, __shfl_down_sync(0xffffffffu, lx, o)); ly = fmin(ly, __shfl_down_sync(0xffffffffu, ly, o));
		hx = fmax(hx, __shfl_down_sync(0xffffffffu, hx, o)); hy = fmax(hy, __shfl_down_sync(0xffffffffu, hy, o));
	}
	__shared__ double4 sh[kB / 32];
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = make_double4(lx, ly, hx, hy);
	__syncthreads();
	if (threadIdx.x != 0) return;
	double4 r = sh[0];
	for (int k = 1; k < kB / 32; ++k) { r.x = fmin(r.x, sh[k].x); r.y = fmin(r.y, sh[k].y); r.z = fmax(r.z, sh[k].z); r.w = fmax(r.w, sh[k].w); }
	double delta = __ddiv_rn(fmax(__dsub_rn(r.z, r.x), __dsub_rn(r.w, r.y)), (double)side);
	if (delta < eps) delta = eps;
	g->minx = r.x; g->miny = r.y; g->rdelta = __ddiv_rn(1.0, delta); g->pad = delta;
}

// evalKeys (appel.cuh:44-55): truncate, clip, x-major flatten
__global__ void __launch_bounds__(kB) keys2_kernel(const double2 *__restrict__ p, int64_t n, const Grid2 *__restrict__ g,
                                                   int side, u32 *__restrict__ keys)
{
	const double mx = g->minx, my = g->miny, rd = g->rdelta;
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const double2 v = p[i];
		int ix = __double2int_rz(__dmul_rn(__dsub_rn(v.x, mx), rd));
		int iy = __double2int_rz(__dmul_rn(__dsub_rn(v.y, my), rd));
		ix = min(max(ix, 0), side - 1);
		iy = min(max(iy, 0), side - 1);
		keys[i] = (u32)(ix * side + iy);
	}
}

// ---------------------------------------------------------------------------------------------
//  stable LSD radix sort of (key, id) pairs; digit of `bits` <= kRsBits bits per pass
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kB) rs_hist_kernel(const u32 *__restrict__ keys, int64_t n, int shift, int bits,
                                                     u32 *__restrict__ hist, int ntiles)
{
	__shared__ u32 h[kRsBins];
	const int bins = 1 << bits;
	const u32 mask = (u32)bins - 1u;
	for (int t = threadIdx.x; t < bins; t += kB) h[t] = 0;
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
	for (int k = 0; k < kRsItems; ++k)
	{
		const int64_t i = base + k * kB + threadIdx.x;
		if (i < n) atomicAdd(&h[(keys[i] >> shift) & mask], 1u);
	}
	__syncthreads();
	for (int t = threadIdx.x; t < bins; t += kB) hist[(size_t)t * ntiles + blockIdx.x] = h[t];
}

// hist[d][tile] -> exclusive prefix over the tiles of digit d (in place), tot[d] = count of digit d.
// One CTA per digit; the rows are contiguous.
__global__ void __launch_bounds__(kB) rs_scan_kernel(u32 *__restrict__ hist, int ntiles, u32 *__restrict__ tot)
{
	__shared__ u32 wsum[kB / 32];
	__shared__ u32 carry;
	u32 *row = hist + (size_t)blockIdx.x * ntiles;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < ntiles; base += kB)
	{
		const int i = base + threadIdx.x;
		const u32 v = (i < ntiles) ? row[i] : 0u;
		u32 x = v;
		for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
		if (lane == 31) wsum[w] = x;
		__syncthreads();
		u32 off = carry;
		for (int k = 0; k < w; ++k) off += wsum[k];
		if (i < ntiles) row[i] = off + x - v;
		__syncthreads();
		if (threadIdx.x == kB - 1) carry = off + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}

// Every warp owns kRsItems*32 consecutive keys of the tile and ranks them in order: stable.
__global__ void __launch_bounds__(kB) rs_scatter_kernel(const u32 *__restrict__ keys_in, const u32 *__restrict__ ids_in,
                                                        u32 *__restrict__ keys_out, u32 *__restrict__ ids_out, int64_t n,
                                                        int shift, int bits, const u32 *__restrict__ hist, int ntiles,
                                                        const u32 *__restrict__ tot)
{
	__shared__ u32 wh[kB / 32][kRsBins];
	__shared__ u32 dbase[kRsBins];
	__shared__ u32 wtot[kB / 32];
	const int bins = 1 << bits;
	const u32 mask = (u32)bins - 1u;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = threadIdx.x; t < (kB / 32) * kRsBins; t += kB) (&wh[0][0])[t] = 0;
	// exclusive scan of the digit totals (bins <= 2 * kB): every thread owns two consecutive digits
	{
		const int d0 = 2 * threadIdx.x;
		const u32 a = (d0 < bins) ? tot[d0] : 0u, b = (d0 + 1 < bins) ? tot[d0 + 1] : 0u;
		u32 x = a + b;
		for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
		if (lane == 31) wtot[warp] = x;
		__syncthreads();
		u32 off = 0;
		for (int k = 0; k < warp; ++k) off += wtot[k];
		const u32 ex = off + x - (a + b);
		if (d0 < bins) dbase[d0] = ex;
		if (d0 + 1 < bins) dbase[d0 + 1] = ex + a;
	}
	__syncthreads();
	const int64_t base = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * (kRsItems * 32);
	u32 key[kRsItems];
#pragma unroll
	for (int k = 0; k < kRsItems; ++k)
	{
		const int64_t i = base + k * 32 + lane;
		key[k] = (i < n) ? keys_in[i] : 0xFFFFFFFFu;
		if (i < n) atomicAdd(&wh[warp][(key[k] >> shift) & mask], 1u);
	}
	__syncthreads();
	for (int d = threadIdx.x; d < bins; d += kB)
	{
		u32 run = dbase[d] + hist[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
		for (int w = 0; w < kB / 32; ++w) { const u32 c = wh[w][d]; wh[w][d] = run; run += c; }
	}
	__syncthreads();
	const u32 lt = (1u << lane) - 1u;
#pragma unroll
	for (int k = 0; k < kRsItems; ++k)
	{
		const int64_t i = base + k * 32 + lane;
		const bool valid = i < n;
		const u32 d = valid ? ((key[k] >> shift) & mask) : (u32)bins;
		const u32 m = __match_any_sync(0xffffffffu, d);
		const u32 rank = __popc(m & lt);
		u32 pos = 0;
		if (valid) pos = wh[warp][d];
		__syncwarp();
		if (valid && rank == 0) wh[warp][d] = pos + __popc(m);
		__syncwarp();
		if (valid)
		{
			keys_out[pos + rank] = key[k];
			ids_out[pos + rank] = ids_in ? ids_in[i] : (u32)i;
		}
	}
}

// gather_krnl + copy_krnl (fmm_cart.cuh:500-505): positions and velocities into cell order
__global__ void __launch_bounds__(kB) gather2_kernel(const double2 *__restrict__ pv, const u32 *__restrict__ ids, int64_t n,
                                                     double2 *__restrict__ out)
{
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const u32 s = ids[i];
		out[i] = pv[s];
		out[n + i] = pv[n + s];
	}
}

// indexLeaves (appel.cuh:141-170): first particle of every cell; lindex has m + 1 entries
__global__ void __launch_bounds__(kB) index2_kernel(const u32 *__restrict__ skeys, int64_t n, int m, int *__restrict__ lindex)
{
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const int k = (int)skeys[i];
		const int kp = (i == 0) ? -1 : (int)skeys[i - 1];
		for (int j = kp + 1; j <= k; ++j) lindex[j] = (int)i;
		if (i == n - 1)
			for (int j = k + 1; j <= m; ++j) lindex[j] = (int)n;
	}
}

// ---------------------------------------------------------------------------------------------
//  upward pass
// ---------------------------------------------------------------------------------------------
struct Tree2
{
	double2 *center;  // per node
	double2 *Z;       // (P+1) per node: harmonic multipole moments
	double2 *Lc;      // (P+1) per node: harmonic local coefficients
	int *mult;        // per node
	const int *lindex; // m + 1, leaf level
	int L;
};

// multLeaves + centerLeaves (appel.cuh:184-243) + P2M (fmm_cart.cuh:68-98): thread per leaf cell
template <int P>
__global__ void __launch_bounds__(kB) leaf2_kernel(Tree2 t, const double2 *__restrict__ sp, int m)
{
	const int c = blockIdx.x * kB + threadIdx.x;
	if (c >= m) return;
	const int node = tbeg(t.L) + c;
	const int b = t.lindex[c], e = t.lindex[c + 1], ml = e - b;
	double sx = 0.0, sy = 0.0;
	for (int j = b; j < e; ++j) { const double2 v = sp[j]; sx = __dadd_rn(sx, v.x); sy = __dadd_rn(sy, v.y); }
	if (ml > 0) { sx = __ddiv_rn(sx, (double)ml); sy = __ddiv_rn(sy, (double)ml); }
	double2 Z[P + 1];
#pragma unroll
	for (int q = 0; q <= P; ++q) Z[q] = make_double2(0.0, 0.0);
	if (P >= 2)
		for (int j = b; j < e; ++j)
		{
			const double2 v = sp[j];
			const double2 d = make_double2(sx - v.x, sy - v.y); // -(z - c)
			double2 pw = d;
#pragma unroll
			for (int q = 2; q <= P; ++q) { pw = cmul(pw, d); Z[q].x += pw.x; Z[q].y += pw.y; }
		}
	t.center[node] = make_double2(sx, sy);
	t.mult[node] = ml;
	double2 *out = t.Z + (size_t)node * (P + 1);
	out[0] = make_double2((double)ml, 0.0);
#pragma unroll
	for (int q = 1; q <= P; ++q)
	{
		constexpr double one = 1.0;
		const double f = one / fact_c(q);
		out[q] = make_double2(Z[q].x * f, Z[q].y * f);
	}
}

// fmm_buildTree2 (fmm_cart.cuh:115-187): thread per parent of level l
template <int P>
__global__ void __launch_bounds__(kB) m2m2_kernel(Tree2 t, int l)
{
	const int sl = 1 << l, slp = sl << 1;
	const int ij0 = blockIdx.x * kB + threadIdx.x;
	if (ij0 >= sl * sl) return;
	const int i = ij0 >> l, j = ij0 & (sl - 1);
	const int node = tbeg(l) + ij0, ijp = tbeg(l + 1) + 2 * (i * slp + j);
	const int ch[4] = {ijp, ijp + 1, ijp + slp, ijp + slp + 1};
	int mk[4], ml = 0;
	double2 ck[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) { mk[k] = t.mult[ch[k]]; ck[k] = t.center[ch[k]]; ml += mk[k]; }
	double cx = 0.0, cy = 0.0;
	double2 Z[P + 1];
#pragma unroll
	for (int q = 0; q <= P; ++q) Z[q] = make_double2(0.0, 0.0);
	if (ml > 0)
	{
#pragma unroll
		for (int k = 0; k < 4; ++k) { cx = __dadd_rn(cx, __dmul_rn((double)mk[k], ck[k].x)); cy = __dadd_rn(cy, __dmul_rn((double)mk[k], ck[k].y)); }
		cx = __ddiv_rn(cx, (double)ml); cy = __ddiv_rn(cy, (double)ml);
		if (P >= 2)
		{
#pragma unroll
			for (int k = 0; k < 4; ++k)
			{
				if (mk[k] == 0) continue;
				const double2 *zc = t.Z + (size_t)ch[k] * (P + 1);
				double2 Zc[P + 1], pw[P + 1];
#pragma unroll
				for (int q = 0; q <= P; ++q) Zc[q] = zc[q];
				const double2 d = make_double2(cx - ck[k].x, cy - ck[k].y);
				pw[0] = make_double2(1.0, 0.0);
#pragma unroll
				for (int q = 1; q <= P; ++q)
				{
					const double2 v = cmul(pw[q - 1], d);
					const double f = 1.0 / (double)q;
					pw[q] = make_double2(v.x * f, v.y * f);
				}
#pragma unroll
				for (int n = 2; n <= P; ++n)
#pragma unroll
					for (int mm = 0; mm <= n; ++mm)
						if (n - mm != 1) cfma(Z[n], Zc[n - mm], pw[mm]);
			}
		}
		Z[0] = make_double2((double)ml, 0.0);
	}
	t.center[node] = make_double2(cx, cy);
	t.mult[node] = ml;
	double2 *out = t.Z + (size_t)node * (P + 1);
#pragma unroll
	for (int q = 0; q <= P; ++q) out[q] = Z[q];
}

// ---------------------------------------------------------------------------------------------
//  M2L of level l (fmm_c2c2, fmm_cart.cuh:214-262) fused with the L2L from level l-1
//  (fmm_pushl, :288-334): thread per node of level l, each node's local written once.
// ---------------------------------------------------------------------------------------------
template <int P, int LANES>
__global__ void __launch_bounds__(128) m2l_l2l2_kernel(Tree2 t, int l, int radius, double eps2)
// LANES lanes share one target node and split its source cells (LANES = 32 on the small levels, where a
// thread per node would serialise ~27 dependent M2L evaluations on a nearly empty machine)
{
	const int sl = 1 << l;
	const int gt = blockIdx.x * 128 + threadIdx.x;
	const int ij = gt / LANES, sub = gt % LANES;
	const bool live = ij < sl * sl;
	const int i = ij >> l, j = ij & (sl - 1);
	const int beg = tbeg(l), node = beg + ij;
	double2 ct = make_double2(0.0, 0.0);
	double2 Lo[P + 1];
#pragma unroll
	for (int n = 0; n <= P; ++n) Lo[n] = make_double2(0.0, 0.0);
	const bool nonempty = live && t.mult[node] > 0;
	if (live) ct = t.center[node];

	if (nonempty)
	{
		const int im = (i >> 1) << 1, jm = (j >> 1) << 1;
		const int kmin = max(im - 2 * radius, 0), kmax = min(im + 2 * radius + 1, sl - 1);
		const int gmin = max(jm - 2 * radius, 0), gmax = min(jm + 2 * radius + 1, sl - 1);
		const int gw = gmax - gmin + 1, ncand = (kmax - kmin + 1) * gw;
		for (int c = sub; c < ncand; c += LANES)
		{
			const int k = kmin + c / gw, g = gmin + c % gw;
			if (!(k > i + radius || k < i - radius || g > j + radius || g < j - radius)) continue;
			const int s = beg + k * sl + g;
			if (t.mult[s] <= 0) continue; // an empty source has Z = 0
			const double2 cs = t.center[s];
			const double dx = ct.x - cs.x, dy = ct.y - cs.y;
			// The reference normalises by r = sqrt(|dz|^2 + eps2) (fmm_cart.cuh:241-249), so its direction
			// d = dz / r is not a unit vector and its gradient tuple is the Chebyshev pair
			// g_m = (-1)^m (m-1)! r^-m (T_m(d0), d1 U_(m-1)(d0)); both obey E_(m+1) = 2 d0 E_m - E_(m-1).
			const double r2 = dx * dx + dy * dy + eps2;
			const double ri = rsqrt(r2);
			const double d0 = dx * ri, d1 = dy * ri, two_d0 = 2.0 * d0;
			double2 G[2 * P + 1];
			G[0] = make_double2(0.0, 0.0);
			{
				double2 Em = make_double2(1.0, 0.0), E = make_double2(d0, d1);
				double sc = -ri; // (-1)^m (m-1)! r^-m
				G[1] = make_double2(sc * E.x, sc * E.y);
#pragma unroll
				for (int mm = 2; mm <= 2 * P; ++mm)
				{
					const double2 En = make_double2(fma(two_d0, E.x, -Em.x), fma(two_d0, E.y, -Em.y));
					Em = E; E = En;
					sc *= -(double)(mm - 1) * ri;
					G[mm] = make_double2(sc * E.x, sc * E.y);
				}
			}
			const double2 *zs = t.Z + (size_t)s * (P + 1);
#pragma unroll
			for (int q = 0; q <= P; ++q)
			{
				if (q == 1) continue; // the dipole about the centre of charge is identically 0
				const double2 Zq = zs[q];
#pragma unroll
				for (int n = 0; n <= P; ++n)
					if (n + q >= 1) cfma_conj(Lo[n], Zq, G[n + q]);
			}
		}
	}
	if (LANES > 1)
	{
#pragma unroll
		for (int n = 0; n <= P; ++n)
#pragma unroll
			for (int o = LANES / 2; o > 0; o >>= 1)
			{
				Lo[n].x += __shfl_xor_sync(0xffffffffu, Lo[n].x, o);
				Lo[n].y += __shfl_xor_sync(0xffffffffu, Lo[n].y, o);
			}
	}
	if (!live || sub != 0) return;
#pragma unroll
	for (int n = 2; n <= P; ++n)
	{
		const double f = 1.0 / fact_c(n);
		Lo[n].x *= f; Lo[n].y *= f;
	}
	if (l > 2)
	{
		const int par = tbeg(l - 1) + (i >> 1) * (sl >> 1) + (j >> 1);
		const double2 cp = t.center[par];
		const double2 cd = make_double2(ct.x - cp.x, -(ct.y - cp.y)); // conj(d)
		const double2 *lp = t.Lc + (size_t)par * (P + 1);
		double2 Lp[P + 1], pw[P + 1];
#pragma unroll
		for (int n = 0; n <= P; ++n) Lp[n] = lp[n];
		pw[0] = make_double2(1.0, 0.0);
#pragma unroll
		for (int n = 1; n <= P; ++n) pw[n] = cmul(pw[n - 1], cd);
#pragma unroll
		for (int q = 0; q <= P; ++q)
#pragma unroll
			for (int mm = q; mm <= P; ++mm)
			{
				const double b = binom_c(mm, q);
				const double2 v = make_double2(Lp[mm].x * b, Lp[mm].y * b);
				cfma(Lo[q], v, pw[mm - q]);
			}
	}
	Lo[0].y = 0.0; // the order-0 local has one (real) entry
	double2 *out = t.Lc + (size_t)node * (P + 1);
#pragma unroll
	for (int n = 0; n <= P; ++n) out[n] = Lo[n];
}

// ---------------------------------------------------------------------------------------------
//  near field p2p2 (appel.cuh:260-318) + L2P (fmm_pushLeaves, fmm_cart.cuh:353-378) + rescale
//  (appel.cuh:506-527) + elastic term (kernel.cuh:119-152): thread per sorted particle
// ---------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kB) near_l2p2_kernel(Tree2 t, const double2 *__restrict__ sp, const u32 *__restrict__ skeys,
                                                       double2 *__restrict__ acc, int64_t i_lo, int64_t i_hi, int radius, int coll, double eps2,
                                                       const double *__restrict__ param, int elastic)
{
	// [i_lo, i_hi): this rank's range of the cell-sorted particles (multi-GPU: cfg.rank / cfg.world; else everything)
	const int64_t i = i_lo + (int64_t)blockIdx.x * kB + threadIdx.x;
	if (i >= i_hi) return;
	const int side = 1 << t.L;
	const int key = (int)skeys[i];
	const int ix = key >> t.L, iy = key & (side - 1);
	const double2 p = sp[i];
	double ax = 0.0, ay = 0.0;
	if (coll)
	{
		const int kmin = max(ix - radius, 0), kmax = min(ix + radius, side - 1);
		const int lmin = max(iy - radius, 0), lmax = min(iy + radius, side - 1);
		for (int k = kmin; k <= kmax; ++k)
		{
			const int b = t.lindex[k * side + lmin], e = t.lindex[k * side + lmax + 1];
			int s = b;
			for (; s + 4 <= e; s += 4)
				pair4(p.x, p.y, __ldg(sp + s), __ldg(sp + s + 1), __ldg(sp + s + 2), __ldg(sp + s + 3), eps2, ax, ay);
			for (; s < e; ++s) pair1(p.x, p.y, __ldg(sp + s), eps2, ax, ay);
		}
	}
	// L2P: f = -sum_n n L_n conj(d)^(n-1), Horner from the top
	const int node = tbeg(t.L) + key;
	const double2 c = t.center[node];
	const double2 cd = make_double2(p.x - c.x, -(p.y - c.y));
	const double2 *lc = t.Lc + (size_t)node * (P + 1);
	double2 h = lc[P];
	h.x *= (double)P; h.y *= (double)P;
#pragma unroll
	for (int q = P - 1; q >= 1; --q)
	{
		const double2 Lq = lc[q];
		const double2 v = cmul(h, cd);
		h = make_double2(fma((double)q, Lq.x, v.x), fma((double)q, Lq.y, v.y));
	}
	ax -= h.x; ay -= h.y;
	if (param) { const double s = param[0]; ax *= s; ay *= s; }
	if (elastic)
	{
		const double kx = param ? param[2] : 1.0, ky = param ? param[3] : 1.0;
		ax = fma(-kx, p.x, ax); ay = fma(-ky, p.y, ay);
	}
	acc[i] = make_double2(ax, ay);
}

// ---------------------------------------------------------------------------------------------
//  direct2 (direct.cuh:140-190): a_i = param[0] * sum_j d / (|d|^2 + eps2)
// ---------------------------------------------------------------------------------------------
constexpr int kDTile = 512;
template <int IPT>
__global__ void __launch_bounds__(kB) direct2_kernel(const double2 *__restrict__ p, double2 *__restrict__ a, int64_t n,
                                                     int64_t ib, int64_t ie, const double *__restrict__ param, double eps2)
{
	__shared__ double2 sm[kDTile];
	double2 pi[IPT];
	double ax[IPT], ay[IPT];
	const int64_t i0 = ib + ((int64_t)blockIdx.x * kB + threadIdx.x);
	const int64_t tstride = (int64_t)gridDim.x * kB;
#pragma unroll
	for (int u = 0; u < IPT; ++u)
	{
		const int64_t i = i0 + u * tstride;
		pi[u] = (i < ie) ? p[i] : make_double2(0.0, 0.0);
		ax[u] = 0.0; ay[u] = 0.0;
	}
	for (int64_t tile = 0; tile < n; tile += kDTile)
	{
		__syncthreads();
		for (int k = threadIdx.x; k < kDTile; k += kB)
			sm[k] = (tile + k < n) ? p[tile + k] : make_double2(0.0, 0.0);
		__syncthreads();
		const int cnt = (int)min((int64_t)kDTile, n - tile);
		int j = 0;
		for (; j + 4 <= cnt; j += 4)
		{
			const double2 q0 = sm[j], q1 = sm[j + 1], q2 = sm[j + 2], q3 = sm[j + 3];
#pragma unroll
			for (int u = 0; u < IPT; ++u) pair4(pi[u].x, pi[u].y, q0, q1, q2, q3, eps2, ax[u], ay[u]);
		}
		for (; j < cnt; ++j)
		{
			const double2 q = sm[j];
#pragma unroll
			for (int u = 0; u < IPT; ++u) pair1(pi[u].x, pi[u].y, q, eps2, ax[u], ay[u]);
		}
	}
	const double s = param ? param[0] : 1.0;
#pragma unroll
	for (int u = 0; u < IPT; ++u)
	{
		const int64_t i = i0 + u * tstride;
		if (i < ie) a[i] = make_double2(s * ax[u], s * ay[u]);
	}
}

// ---------------------------------------------------------------------------------------------
//  streaming kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kB) step2_kernel(double2 *__restrict__ b, const double2 *__restrict__ a, double ds, int64_t n)
{
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		double2 x = b[i];
		const double2 v = a[i];
		x.x = fma(v.x, ds, x.x); x.y = fma(v.y, ds, x.y);
		b[i] = x;
	}
}

// K(kc) then D(dc) in one pass: v += a*kc; x += v*dc
__global__ void __launch_bounds__(kB) kick_drift2_kernel(double2 *__restrict__ x, double2 *__restrict__ v,
                                                         const double2 *__restrict__ a, double kc, double dc, int64_t n)
{
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		double2 vv = v[i], xx = x[i];
		const double2 aa = a[i];
		vv.x = fma(aa.x, kc, vv.x); vv.y = fma(aa.y, kc, vv.y);
		xx.x = fma(vv.x, dc, xx.x); xx.y = fma(vv.y, dc, xx.y);
		v[i] = vv; x[i] = xx;
	}
}

__global__ void __launch_bounds__(kB) elastic2_kernel(const double2 *__restrict__ x, double2 *__restrict__ a,
                                                      const double *__restrict__ k2, int64_t n)
{
	const double kx = k2 ? k2[0] : 1.0, ky = k2 ? k2[1] : 1.0;
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		double2 aa = a[i];
		const double2 xx = x[i];
		aa.x = fma(-kx, xx.x, aa.x); aa.y = fma(-ky, xx.y, aa.y);
		a[i] = aa;
	}
}

__device__ __forceinline__ double block_sum2(double v, double *sh)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	__syncthreads();
	if (l == 0) sh[w] = v;
	__syncthreads();
	double r = 0.0;
	if (w == 0)
	{
		r = (l < (kB >> 5)) ? sh[l] : 0.0;
		for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
	}
	return r;
}

// out[0] += sum 1/2 v^2, out[1] += 1/2 sum k o x^2
__global__ void __launch_bounds__(kB) kin_el2_kernel(const double2 *__restrict__ x, const double2 *__restrict__ v, int64_t n,
                                                     const double *__restrict__ param, double *__restrict__ out)
{
	__shared__ double sh[kB / 32];
	const double kx = param ? param[2] : 1.0, ky = param ? param[3] : 1.0;
	double ke = 0.0, el = 0.0;
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const double2 vv = v[i], xx = x[i];
		ke += 0.5 * (vv.x * vv.x + vv.y * vv.y);
		el += 0.5 * (kx * xx.x * xx.x + ky * xx.y * xx.y);
	}
	ke = block_sum2(ke, sh);
	if (threadIdx.x == 0) atomicAdd(out, ke);
	el = block_sum2(el, sh);
	if (threadIdx.x == 0) atomicAdd(out + 1, el);
}

// out[0] += sum_{i<j} -1/2 log(d^2 + eps2)  (pair term of the 2D Hamiltonian)
__global__ void __launch_bounds__(kB) pair_energy2_kernel(const double2 *__restrict__ p, int64_t n, double eps2, double *__restrict__ out)
{
	__shared__ double2 sm[kB];
	__shared__ double sh[kB / 32];
	const int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x;
	const double2 pi = (i < n) ? p[i] : make_double2(0.0, 0.0);
	double e = 0.0;
	// tiles of sources with index > the first target of this CTA
	for (int64_t tile = (int64_t)blockIdx.x * kB; tile < n; tile += kB)
	{
		__syncthreads();
		sm[threadIdx.x] = (tile + threadIdx.x < n) ? p[tile + threadIdx.x] : make_double2(0.0, 0.0);
		__syncthreads();
		const int cnt = (int)min((int64_t)kB, n - tile);
		if (i < n)
			for (int j = 0; j < cnt; ++j)
				if (tile + j > i)
				{
					const double dx = pi.x - sm[j].x, dy = pi.y - sm[j].y;
					e -= 0.5 * log(dx * dx + dy * dy + eps2);
				}
	}
	e = block_sum2(e, sh);
	if (threadIdx.x == 0) atomicAdd(out, e);
}

// out[0] += sum_i |a-ref| / sqrt(|ref|^2 + 1e-18); out[1] = max (bits of a non-negative double)
__global__ void __launch_bounds__(kB) rel_err2_kernel(const double2 *__restrict__ a, const double2 *__restrict__ r, int64_t n,
                                                      double *__restrict__ out)
{
	__shared__ double sh[kB / 32];
	double s = 0.0, mx = 0.0;
	const int64_t stride = (int64_t)gridDim.x * kB;
	for (int64_t i = (int64_t)blockIdx.x * kB + threadIdx.x; i < n; i += stride)
	{
		const double2 x = a[i], y = r[i];
		const double dx = x.x - y.x, dy = x.y - y.y;
		const double e = sqrt((dx * dx + dy * dy) / (y.x * y.x + y.y * y.y + 1.e-18));
		s += e; mx = fmax(mx, e);
	}
	for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
	if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)(out + 1), (unsigned long long)__double_as_longlong(mx));
	s = block_sum2(s, sh);
	if (threadIdx.x == 0) atomicAdd(out, s);
}

} // namespace

// ---------------------------------------------------------------------------------------------
//  host side
// ---------------------------------------------------------------------------------------------
enum Phase2 { P2_KEYS = 0, P2_SORT, P2_PERMUTE, P2_LEAVES, P2_M2M, P2_M2L_L2L, P2_NEAR_L2P, P2_COUNT };
static const char *kPhase2Names[P2_COUNT] = {"bbox_keys", "sort", "permute", "leaves_p2m", "m2m", "m2l_l2l", "near_l2p"};

struct Fmm2Plan
{
	int64_t n = 0;
	int order = 0, L = 0;
	DevBuf keysA, keysB, idsA, idsB, hist, tmp, part, grid, center, Z, Lc, mult, lindex, red;
	cudaEvent_t ev[P2_COUNT + 1] = {};
	bool have_events = false, timed = false;
	float ms[P2_COUNT] = {};
	double tot_ms[P2_COUNT] = {};
	int64_t evals = 0;
	const u32 *sorted_keys = nullptr, *sorted_ids = nullptr;
};

int fmm2_levels(int64_t n, int order, double dens)
// fmm_cart.cuh:416-418
{
	const double s = (double)order * std::sqrt((double)order);
	int L = (int)std::round(std::log2(dens * (double)n / s) / 2);
	return std::max(L, 2);
}

static double eps2_of(const nbco_ctx *ctx) { return ctx->cfg.eps2_d > 0.0 ? ctx->cfg.eps2_d : (double)ctx->cfg.eps2; }
// the near field multiplies four softened squared distances before taking one reciprocal
static int check_eps2(double eps2)
{
	if (!(eps2 >= 1.e-70)) { set_error("2D path: eps2 = %g is below 1e-70 (products of four squared distances would underflow)", eps2); return NBCO_ERR_INVALID; }
	return NBCO_OK;
}

template <int P>
static void run_order(nbco_ctx *ctx, Fmm2Plan &pl, Tree2 t, double2 *d_pos, double2 *d_acc, int64_t n, const double *d_param,
                      bool elastic, int radius, double eps2)
{
	cudaStream_t st = ctx->stream;
	const int L = pl.L, m = 1 << (2 * L);
	leaf2_kernel<P><<<(m + kB - 1) / kB, kB, 0, st>>>(t, d_pos, m);
	ctx->launches++;
	cudaEventRecord(pl.ev[P2_M2M], st);
	for (int l = L - 1; l >= 2; --l)
	{
		const int cnt = 1 << (2 * l);
		m2m2_kernel<P><<<(cnt + kB - 1) / kB, kB, 0, st>>>(t, l);
		ctx->launches++;
	}
	cudaEventRecord(pl.ev[P2_M2L_L2L], st);
	for (int l = 2; l <= L; ++l)
	{
		const int cnt = 1 << (2 * l);
		if (l <= 7) m2l_l2l2_kernel<P, 32><<<(cnt * 32 + 127) / 128, 128, 0, st>>>(t, l, radius, eps2);
		else m2l_l2l2_kernel<P, 1><<<(cnt + 127) / 128, 128, 0, st>>>(t, l, radius, eps2);
		ctx->launches++;
	}
	cudaEventRecord(pl.ev[P2_NEAR_L2P], st);
	// multi-GPU (cfg.world > 1): the tree is replicated (every rank sorts and summarises all particles: 0.56 of the 1.49 ms of
	// an evaluation at N = 2^22), the dominant kernel -- near field + L2P -- runs on the rank's own range of the cell-sorted
	// particles; the caller all-gathers the accelerations (coulomb_oscillators_b200/parallel.py: fmm2_integrate_sharded)
	int64_t i_lo = 0, i_hi = n;
	if (ctx->cfg.world > 1) nbco_shard_range(n, ctx->cfg.rank, ctx->cfg.world, &i_lo, &i_hi);
	if (i_hi > i_lo)
	{
		near_l2p2_kernel<P><<<(unsigned)((i_hi - i_lo + kB - 1) / kB), kB, 0, st>>>(t, d_pos, pl.sorted_keys, d_acc, i_lo, i_hi, radius,
		                                                                             ctx->cfg.coll, eps2, d_param, elastic ? 1 : 0);
		ctx->launches++;
	}
	cudaEventRecord(pl.ev[P2_COUNT], st);
}

int fmm2_launch(nbco_ctx *ctx, double *d_pos, double *d_acc, int64_t n, const double *d_param, bool fuse_elastic)
{
	if (!d_pos || !d_acc || n < 1) { set_error("fmm2: bad arguments"); return NBCO_ERR_INVALID; }
	if (n > 0x7FFFFFF0LL) { set_error("fmm2: n too large"); return NBCO_ERR_INVALID; }
	const int P = ctx->cfg.order;
	if (P < 1 || P > kMaxP2) { set_error("fmm2: order %d outside 1..%d", P, kMaxP2); return NBCO_ERR_INVALID; }
	NBCO_TRY(check_eps2(eps2_of(ctx)));
	const int radius = (int)ctx->cfg.radius; // int radius = tree_radius (fmm_cart.cuh:398)
	if (radius < 1) { set_error("fmm2: radius must be >= 1"); return NBCO_ERR_INVALID; }
	if (!ctx->fmm2) ctx->fmm2 = new Fmm2Plan();
	Fmm2Plan &pl = *ctx->fmm2;
	const int L = ctx->cfg.max_level > 0 ? std::max(ctx->cfg.max_level, 2) : fmm2_levels(n, P, (double)ctx->cfg.dens_inhom);
	if (L > 12) { set_error("fmm2: %d levels need 4^%d cells", L, L); return NBCO_ERR_INVALID; }
	pl.n = n; pl.order = P; pl.L = L;
	const int side = 1 << L, m = side * side;
	const int64_t ntot = (((int64_t)1 << (2 * (L + 1))) - 1) / 3;
	const int ntiles = (int)((n + kRsTile - 1) / kRsTile);
	const int nbits = 2 * L, npass = (nbits + kRsBits - 1) / kRsBits, bits = (nbits + npass - 1) / npass;
	const int nbb = grid_for(n, kB, ctx->sm_count, 8);
	NBCO_TRY(pl.keysA.reserve(4 * (size_t)n)); NBCO_TRY(pl.keysB.reserve(4 * (size_t)n));
	NBCO_TRY(pl.idsA.reserve(4 * (size_t)n)); NBCO_TRY(pl.idsB.reserve(4 * (size_t)n));
	NBCO_TRY(pl.hist.reserve(4 * ((size_t)ntiles + 1) * kRsBins));
	NBCO_TRY(pl.tmp.reserve(32 * (size_t)n));
	NBCO_TRY(pl.part.reserve(32 * (size_t)nbb));
	NBCO_TRY(pl.grid.reserve(sizeof(Grid2)));
	NBCO_TRY(pl.center.reserve(16 * (size_t)ntot));
	NBCO_TRY(pl.Z.reserve(16 * (size_t)ntot * (P + 1)));
	NBCO_TRY(pl.Lc.reserve(16 * (size_t)ntot * (P + 1)));
	NBCO_TRY(pl.mult.reserve(4 * (size_t)ntot));
	NBCO_TRY(pl.lindex.reserve(4 * (size_t)(m + 1)));
	if (!pl.have_events)
	{
		for (int k = 0; k <= P2_COUNT; ++k) NBCO_CUDA(cudaEventCreate(&pl.ev[k]));
		pl.have_events = true;
	}
	cudaStream_t st = ctx->stream;
	double2 *pos = (double2 *)d_pos;

	NBCO_CUDA(cudaEventRecord(pl.ev[P2_KEYS], st));
	bbox2_kernel<<<nbb, kB, 0, st>>>(pos, n, pl.part.as<double4>());
	grid2_kernel<<<1, kB, 0, st>>>(pl.part.as<double4>(), nbb, side, std::sqrt(eps2_of(ctx)), pl.grid.as<Grid2>());
	keys2_kernel<<<nbb, kB, 0, st>>>(pos, n, pl.grid.as<Grid2>(), side, pl.keysA.as<u32>());
	ctx->launches += 3;

	NBCO_CUDA(cudaEventRecord(pl.ev[P2_SORT], st));
	u32 *kin = pl.keysA.as<u32>(), *kout = pl.keysB.as<u32>(), *iin = pl.idsA.as<u32>(), *iout = pl.idsB.as<u32>();
	for (int pass = 0; pass < npass; ++pass)
	{
		const int shift = pass * bits, b = std::min(bits, nbits - shift);
		rs_hist_kernel<<<ntiles, kB, 0, st>>>(kin, n, shift, b, pl.hist.as<u32>(), ntiles);
		u32 *tot = pl.hist.as<u32>() + (size_t)ntiles * kRsBins;
		rs_scan_kernel<<<1 << b, kB, 0, st>>>(pl.hist.as<u32>(), ntiles, tot);
		rs_scatter_kernel<<<ntiles, kB, 0, st>>>(kin, pass == 0 ? nullptr : iin, kout, iout, n, shift, b, pl.hist.as<u32>(), ntiles, tot);
		ctx->launches += 3;
		std::swap(kin, kout); std::swap(iin, iout);
	}
	pl.sorted_keys = kin; pl.sorted_ids = iin;

	NBCO_CUDA(cudaEventRecord(pl.ev[P2_PERMUTE], st));
	gather2_kernel<<<nbb, kB, 0, st>>>(pos, pl.sorted_ids, n, pl.tmp.as<double2>());
	ctx->launches++;
	NBCO_CUDA(cudaMemcpyAsync(pos, pl.tmp.p, 32 * (size_t)n, cudaMemcpyDeviceToDevice, st));

	NBCO_CUDA(cudaEventRecord(pl.ev[P2_LEAVES], st));
	index2_kernel<<<nbb, kB, 0, st>>>(pl.sorted_keys, n, m, pl.lindex.as<int>());
	ctx->launches++;

	Tree2 t;
	t.center = pl.center.as<double2>(); t.Z = pl.Z.as<double2>(); t.Lc = pl.Lc.as<double2>();
	t.mult = pl.mult.as<int>(); t.lindex = pl.lindex.as<int>(); t.L = L;
	const double eps2 = eps2_of(ctx);
	double2 *acc = (double2 *)d_acc;
	switch (P)
	{
#define NBCO2_CASE(K) case K: run_order<K>(ctx, pl, t, pos, acc, n, d_param, fuse_elastic, radius, eps2); break;
		NBCO2_CASE(1) NBCO2_CASE(2) NBCO2_CASE(3) NBCO2_CASE(4) NBCO2_CASE(5)
		NBCO2_CASE(6) NBCO2_CASE(7) NBCO2_CASE(8) NBCO2_CASE(9) NBCO2_CASE(10)
#undef NBCO2_CASE
		default: set_error("fmm2: order %d", P); return NBCO_ERR_INVALID;
	}
	NBCO_CUDA(cudaGetLastError());
	pl.timed = true;
	pl.evals++;
	return NBCO_OK;
}

static int collect_phase_ms(nbco_ctx *ctx)
{
	Fmm2Plan &pl = *ctx->fmm2;
	if (!pl.timed) return NBCO_OK;
	NBCO_CUDA(cudaEventSynchronize(pl.ev[P2_COUNT]));
	for (int k = 0; k < P2_COUNT; ++k)
	{
		float v = 0.f;
		cudaEventElapsedTime(&v, pl.ev[k], pl.ev[k + 1]);
		pl.ms[k] = v; pl.tot_ms[k] += v;
	}
	pl.timed = false;
	return NBCO_OK;
}

int direct2_launch(nbco_ctx *ctx, const double *d_pos, double *d_acc, int64_t n, const double *d_param)
{
	if (!d_pos || !d_acc || n < 1) { set_error("direct2: bad arguments"); return NBCO_ERR_INVALID; }
	int64_t ib, ie;
	nbco_shard_range(n, ctx->cfg.rank, ctx->cfg.world, &ib, &ie);
	const int64_t cnt = ie - ib;
	if (cnt <= 0) return NBCO_OK;
	const double eps2 = eps2_of(ctx);
	NBCO_TRY(check_eps2(eps2));
	// two targets per thread once there are enough targets to fill the machine twice
	if (cnt >= (int64_t)ctx->sm_count * kB * 4)
	{
		const unsigned g = (unsigned)((cnt + 2 * kB - 1) / (2 * kB));
		direct2_kernel<2><<<g, kB, 0, ctx->stream>>>((const double2 *)d_pos, (double2 *)d_acc, n, ib, ie, d_param, eps2);
	}
	else
	{
		const unsigned g = (unsigned)((cnt + kB - 1) / kB);
		direct2_kernel<1><<<g, kB, 0, ctx->stream>>>((const double2 *)d_pos, (double2 *)d_acc, n, ib, ie, d_param, eps2);
	}
	ctx->launches++;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int step2_launch(nbco_ctx *ctx, double *d_b, const double *d_a, double ds, int64_t n)
{
	step2_kernel<<<grid_for(n, kB, ctx->sm_count, 8), kB, 0, ctx->stream>>>((double2 *)d_b, (const double2 *)d_a, ds, n);
	ctx->launches++;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int kick_drift2_launch(nbco_ctx *ctx, double *d_pos, double *d_vel, const double *d_acc, double kc, double dc, int64_t n)
{
	kick_drift2_kernel<<<grid_for(n, kB, ctx->sm_count, 8), kB, 0, ctx->stream>>>((double2 *)d_pos, (double2 *)d_vel,
	                                                                             (const double2 *)d_acc, kc, dc, n);
	ctx->launches++;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int add_elastic2_launch(nbco_ctx *ctx, const double *d_pos, double *d_acc, int64_t n, const double *d_k2)
{
	elastic2_kernel<<<grid_for(n, kB, ctx->sm_count, 8), kB, 0, ctx->stream>>>((const double2 *)d_pos, (double2 *)d_acc, d_k2, n);
	ctx->launches++;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

void fmm2_destroy(nbco_ctx *ctx)
{
	if (!ctx->fmm2) return;
	Fmm2Plan &pl = *ctx->fmm2;
	DevBuf *all[] = {&pl.keysA, &pl.keysB, &pl.idsA, &pl.idsB, &pl.hist, &pl.tmp, &pl.part, &pl.grid, &pl.center, &pl.Z, &pl.Lc,
	                 &pl.mult, &pl.lindex, &pl.red};
	for (DevBuf *b : all) b->release();
	if (pl.have_events) for (int k = 0; k <= P2_COUNT; ++k) cudaEventDestroy(pl.ev[k]);
	delete ctx->fmm2;
	ctx->fmm2 = nullptr;
}

} // namespace nbco

using namespace nbco;

extern "C" {

int nbco_fmm2_levels(int64_t n, int32_t order, double dens_inhom) { return fmm2_levels(n, order, dens_inhom); }

int nbco_energy2(nbco_ctx *ctx, const void *d_buf, int64_t n, const void *d_param, double *h_out3)
{
	if (!ctx || !d_buf || !h_out3 || n < 1) { set_error("energy2: bad arguments"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>();
	NBCO_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(double), ctx->stream));
	const double2 *x = (const double2 *)d_buf;
	kin_el2_kernel<<<grid_for(n, kB, ctx->sm_count, 4), kB, 0, ctx->stream>>>(x, x + n, n, (const double *)d_param, d);
	pair_energy2_kernel<<<(unsigned)((n + kB - 1) / kB), kB, 0, ctx->stream>>>(x, n, eps2_of(ctx), d + 2);
	ctx->launches += 2;
	double h[4], scale = 1.0;
	NBCO_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	if (d_param) NBCO_CUDA(cudaMemcpyAsync(&scale, d_param, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	h_out3[0] = h[0]; h_out3[1] = h[1]; h_out3[2] = scale * h[2];
	return NBCO_OK;
}

int nbco_mean_rel_err2(nbco_ctx *ctx, const void *d_a, const void *d_ref, int64_t n, double *h_mean, double *h_max)
{
	if (!ctx || !d_a || !d_ref || n < 1) { set_error("mean_rel_err2: bad arguments"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	NBCO_TRY(ctx->red.reserve(4 * sizeof(double)));
	double *d = ctx->red.as<double>();
	NBCO_CUDA(cudaMemsetAsync(d, 0, 4 * sizeof(double), ctx->stream));
	rel_err2_kernel<<<grid_for(n, kB, ctx->sm_count, 4), kB, 0, ctx->stream>>>((const double2 *)d_a, (const double2 *)d_ref, n, d);
	ctx->launches++;
	double h[2];
	NBCO_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (h_mean) *h_mean = h[0] / (double)n;
	if (h_max) *h_max = h[1];
	return NBCO_OK;
}

int nbco_fmm2_get_info(nbco_ctx *ctx, nbco_fmm2_info *info)
{
	if (!ctx || !info || !ctx->fmm2) { set_error("no 2D FMM evaluation yet"); return NBCO_ERR_INVALID; }
	const Fmm2Plan &pl = *ctx->fmm2;
	memset(info, 0, sizeof(*info));
	info->levels = pl.L; info->order = pl.order; info->n = pl.n;
	info->nodes = (((int64_t)1 << (2 * (pl.L + 1))) - 1) / 3;
	info->coeffs = pl.order + 1;
	info->kernel_launches = ctx->launches;
	info->evals = pl.evals;
	return NBCO_OK;
}

int nbco_fmm2_get_tree(nbco_ctx *ctx, double *h_center, double *h_mpole, double *h_local, int32_t *h_mult,
                       int32_t *h_leaf_index, int32_t *h_perm)
{
	if (!ctx || !ctx->fmm2) { set_error("no 2D FMM evaluation yet"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	Fmm2Plan &pl = *ctx->fmm2;
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	const size_t ntot = (size_t)((((int64_t)1 << (2 * (pl.L + 1))) - 1) / 3), c = (size_t)pl.order + 1, m = (size_t)1 << (2 * pl.L);
	if (h_center) NBCO_CUDA(cudaMemcpy(h_center, pl.center.p, 16 * ntot, cudaMemcpyDeviceToHost));
	if (h_mpole) NBCO_CUDA(cudaMemcpy(h_mpole, pl.Z.p, 16 * ntot * c, cudaMemcpyDeviceToHost));
	if (h_local) NBCO_CUDA(cudaMemcpy(h_local, pl.Lc.p, 16 * ntot * c, cudaMemcpyDeviceToHost));
	if (h_mult) NBCO_CUDA(cudaMemcpy(h_mult, pl.mult.p, 4 * ntot, cudaMemcpyDeviceToHost));
	if (h_leaf_index) NBCO_CUDA(cudaMemcpy(h_leaf_index, pl.lindex.p, 4 * (m + 1), cudaMemcpyDeviceToHost));
	if (h_perm) NBCO_CUDA(cudaMemcpy(h_perm, pl.sorted_ids, 4 * (size_t)pl.n, cudaMemcpyDeviceToHost));
	return NBCO_OK;
}

int nbco_fmm2_get_phase_ms(nbco_ctx *ctx, const char **names, float *ms, int cap)
{
	if (!ctx || !ctx->fmm2) return 0;
	if (collect_phase_ms(ctx) != NBCO_OK) return 0;
	int k = 0;
	for (; k < P2_COUNT && k < cap; ++k) { names[k] = kPhase2Names[k]; ms[k] = ctx->fmm2->ms[k]; }
	return k;
}

} // extern "C"
