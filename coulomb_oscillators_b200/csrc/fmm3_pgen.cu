// fmm3_pgen.cu -- runtime-order version of the FMM operator passes, used for the orders that
// have no unrolled instantiation (fmm3_p<N>.cu).  Same maths as fmm_ops.cuh with the order as a
// loop bound and the tensors in local memory; the reference makes the same split between its
// static_* templates (orders <= 5) and generic loops (fmm_cart_base3.cuh:1338-1345,1446-1452).
#include "fmm3_common.cuh"

namespace nbco {

namespace {

constexpr int kMaxP = NBCO_MAX_ORDER;
constexpr int kMaxSym = (kMaxP + 1) * (kMaxP + 2) * (kMaxP + 3) / 6; // orders 0..kMaxP
constexpr int kMaxTrl = (kMaxP + 1) * (kMaxP + 1);
constexpr int kMaxElems = (kMaxP + 1) * (kMaxP + 2) / 2;

__constant__ float c_fact[13] = {1.f, 1.f, 2.f, 6.f, 24.f, 120.f, 720.f, 5040.f, 40320.f, 362880.f, 3628800.f, 39916800.f, 479001600.f};
// (2k-1)!! for k = 0..12
__constant__ float c_odf[13] = {1.f, 1.f, 3.f, 15.f, 105.f, 945.f, 10395.f, 135135.f, 2027025.f, 34459425.f, 654729075.f, 13749310575.f, 316234143225.f};

__device__ __forceinline__ int g_sym_off(int p) { return p * (p + 1) * (p + 2) / 6; }
__device__ __forceinline__ int g_trl_off(int p) { return p * p; }
__device__ __forceinline__ int g_sym_idx(int x, int z, int n) { return (n * (n + 1) - (n - z) * (n - z + 1)) / 2 + n - x; }
__device__ __forceinline__ float g_binom(int n, int k) { return (k < 0 || k > n) ? 0.f : c_fact[n] / (c_fact[k] * c_fact[n - k]); }
__device__ __forceinline__ float g_trinom(int n, int kx, int kz) { return c_fact[n] / (c_fact[kx] * c_fact[n - kx - kz] * c_fact[kz]); }
__device__ __forceinline__ float g_coeff13(int n, int m) { return ((m & 1) ? -1.f : 1.f) * c_odf[n - m]; } // (-1)^m (2(n-m)-1)!!
__device__ __forceinline__ float g_coeff2(int n, int k) { return c_fact[n] / ((float)(1 << k) * c_fact[k] * c_fact[n - 2 * k]); }

struct Pows { float x[kMaxP + 1], y[kMaxP + 1], z[kMaxP + 1]; };
__device__ __forceinline__ void make_pows(Pows &pw, int N, float dx, float dy, float dz)
{
	pw.x[0] = pw.y[0] = pw.z[0] = 1.f;
	for (int i = 1; i <= N; ++i) { pw.x[i] = pw.x[i-1] * dx; pw.y[i] = pw.y[i-1] * dy; pw.z[i] = pw.z[i-1] * dz; }
}

__device__ void g_refine(float *A, int n)
{
	for (int z = 2; z <= n; ++z)
		for (int x = n - z; x >= 0; --x)
			A[g_sym_idx(x, z, n)] = -A[g_sym_idx(x + 2, z - 2, n)] - A[g_sym_idx(x, z - 2, n)];
}

__device__ void g_contract_trl_ma(float *C, const float *A, const float *B, float c, int nA, int nB)
{
	const int nC = nA - nB;
	int i = 0;
	for (int z = 0; z <= (nC < 1 ? nC : 1); ++z)
		for (int x = nC - z; x >= 0; --x)
		{
			float t = 0.f;
			for (int kz = 0; kz <= nB; ++kz)
				for (int kx = 0; kx <= nB - kz; ++kx)
					t += g_trinom(nB, kx, kz) * A[g_sym_idx(x + kx, z + kz, nA)] * B[g_sym_idx(kx, kz, nB)];
			C[i++] += c * t;
		}
}

__device__ void g_p2m_acc(float *M, int P, float dx, float dy, float dz)
{
	Pows pw; make_pows(pw, P - 1, dx, dy, dz);
	for (int q = 2; q <= P - 1; ++q)
	{
		const float C = ((q & 1) ? -1.f : 1.f) / c_fact[q];
		float *Mq = M + g_sym_off(q);
		int i = 0;
		for (int z = 0; z <= q; ++z)
			for (int x = q - z; x >= 0; --x)
				Mq[i++] += C * pw.x[x] * pw.y[q - x - z] * pw.z[z];
	}
}

__device__ void g_m2m_acc(float *Mout, const float *Min, int P, float dx, float dy, float dz)
{
	Pows pw; make_pows(pw, P - 1, dx, dy, dz);
	for (int n = 2; n <= P - 1; ++n)
	{
		int i = 0;
		for (int z = 0; z <= n; ++z)
			for (int x = n - z; x >= 0; --x)
			{
				const int y = n - x - z;
				float t = 0.f;
				for (int m = 0; m <= n; ++m)
				{
					if (n - m == 1) continue;
					const float *Mo = Min + g_sym_off(n - m);
					float c = 0.f;
					for (int k1 = 0; k1 <= (x < m ? x : m); ++k1)
					{
						float c2 = 0.f;
						const int lo = m - k1 - y > 0 ? m - k1 - y : 0, hi = z < m - k1 ? z : m - k1;
						for (int k3 = lo; k3 <= hi; ++k3)
						{
							const int k2 = m - k1 - k3;
							c2 += g_binom(y, k2) * g_binom(z, k3) * pw.y[k2] * pw.z[k3] * Mo[g_sym_idx(x - k1, z - k3, n - m)];
						}
						c += c2 * g_binom(x, k1) * pw.x[k1];
					}
					t += c * c_fact[n - m];
				}
				Mout[g_sym_off(n) + i++] += t / c_fact[n];
			}
	}
}

__device__ void g_m2l_acc(float *L, const float *M, int P, float ux, float uy, float uz, float rinv)
{
	Pows pw; make_pows(pw, P, ux, uy, uz);
	float g[kMaxElems];
	float cm = rinv;
	for (int m = 1; m <= P; ++m)
	{
		cm *= rinv; // 1 / r^(m+1)
		const float sgn = (m & 1) ? -1.f : 1.f;
		for (int z = 0; z <= 1; ++z)
			for (int x = m - z; x >= 0; --x)
			{
				const int y = m - x - z;
				float t1 = 0.f;
				for (int k1 = 0; k1 <= x / 2; ++k1)
				{
					float t2 = 0.f;
					for (int k2 = 0; k2 <= y / 2; ++k2)
						t2 += g_coeff13(m, k1 + k2) * g_coeff2(y, k2) * pw.y[y - 2 * k2];
					t1 += t2 * g_coeff2(x, k1) * pw.x[x - 2 * k1];
				}
				g[g_sym_idx(x, z, m)] = sgn * t1 * pw.z[z];
			}
		g_refine(g, m);
		for (int n = 1; n <= m; ++n)
		{
			const int k = m - n;
			if (k == 1) continue;
			g_contract_trl_ma(L + g_trl_off(n), g, M + g_sym_off(k), cm / c_fact[n], m, k);
		}
	}
}

__device__ void g_local_expand(float *S, const float *Ltrl, int P)
{
	S[0] = 0.f;
	for (int q = 1; q <= P; ++q)
	{
		float *dst = S + g_sym_off(q);
		for (int j = 0; j < 2 * q + 1; ++j) dst[j] = Ltrl[g_trl_off(q) + j];
		g_refine(dst, q);
	}
}

__device__ void g_tensor_pow(float *pwt, int q, const Pows &pw)
{
	int i = 0;
	for (int z = 0; z <= q; ++z)
		for (int x = q - z; x >= 0; --x)
			pwt[i++] = pw.x[x] * pw.y[q - x - z] * pw.z[z];
}

__device__ void g_l2l_acc(float *Lc, const float *S, int P, float dx, float dy, float dz)
{
	Pows pw; make_pows(pw, P - 1, dx, dy, dz);
	float pwt[kMaxElems];
	for (int n = 1; n <= P; ++n)
		for (int m = n; m <= P; ++m)
		{
			g_tensor_pow(pwt, m - n, pw);
			g_contract_trl_ma(Lc + g_trl_off(n), S + g_sym_off(m), pwt, g_binom(m, m - n), m, m - n);
		}
}

__device__ void g_l2p_field(float *f, const float *S, int P, float dx, float dy, float dz)
{
	Pows pw; make_pows(pw, P - 1, dx, dy, dz);
	float pwt[kMaxElems];
	float t[3] = {0.f, 0.f, 0.f};
	for (int n = 1; n <= P; ++n)
	{
		g_tensor_pow(pwt, n - 1, pw);
		g_contract_trl_ma(t, S + g_sym_off(n), pwt, (float)n, n, n - 1);
	}
	f[0] = -t[0]; f[1] = -t[1]; f[2] = -t[2];
}

// ---- kernels: same structure as fmm3_order.cuh ----
__global__ void __launch_bounds__(128) g_leaf_p2m_kernel(TreeData t, const float *__restrict__ spos, int64_t n, int L, int P, int first, int count)
{
	const int beg = kd_beg(L), offM = g_sym_off(P);
	for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < first + count; i += gridDim.x * blockDim.x)
	{
		int64_t st = seg_start(n, i, L);
		int cnt = (int)(seg_start(n, i + 1, L) - st);
		const float *p = spos + 3 * st;
		float cx = 0.f, cy = 0.f, cz = 0.f;
		for (int j = 0; j < cnt; ++j) { cx += p[3*j]; cy += p[3*j+1]; cz += p[3*j+2]; }
		if (cnt > 0) { float f = (float)cnt; cx = __fdiv_rn(cx, f); cy = __fdiv_rn(cy, f); cz = __fdiv_rn(cz, f); }
		t.center[beg + i] = make_float4(cx, cy, cz, t.size2[beg + i]);
		float M[kMaxSym];
		for (int k = 0; k < offM; ++k) M[k] = 0.f;
		if (P >= 3)
			for (int j = 0; j < cnt; ++j)
				g_p2m_acc(M, P, p[3*j] - cx, p[3*j+1] - cy, p[3*j+2] - cz);
		M[0] = (float)cnt;
		float *out = t.mpole + (int64_t)(beg + i) * t.sM;
		for (int k = 0; k < offM; ++k) out[k] = M[k];
	}
}

__device__ void g_m2m_node(const TreeData &t, int node, int64_t n, int l, int i, int P)
{
	const int c0 = 2*node + 1, c1 = c0 + 1, offM = g_sym_off(P);
	const float m0 = (float)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
	const float m1 = (float)(seg_start(n, 2*i + 2, l + 1) - seg_start(n, 2*i + 1, l + 1));
	const float mt = (float)(seg_start(n, i + 1, l) - seg_start(n, i, l));
	const float4 a = node_center(t, c0), b = node_center(t, c1);
	float cx = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.x), __fmul_rn(m1, b.x)), mt);
	float cy = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.y), __fmul_rn(m1, b.y)), mt);
	float cz = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.z), __fmul_rn(m1, b.z)), mt);
	float M[kMaxSym];
	for (int k = 0; k < offM; ++k) M[k] = 0.f;
	if (P >= 3)
	{
		g_m2m_acc(M, node_mpole(t, c0), P, cx - a.x, cy - a.y, cz - a.z);
		g_m2m_acc(M, node_mpole(t, c1), P, cx - b.x, cy - b.y, cz - b.z);
	}
	M[0] = mt;
	float *out = t.mpole + (int64_t)node * t.sM;
	for (int k = 0; k < offM; ++k) out[k] = M[k];
	t.center[node] = make_float4(cx, cy, cz, t.size2[node]);
}

__global__ void __launch_bounds__(128) g_m2m_level_kernel(TreeData t, int64_t n, int l, int P, int first, int count)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) g_m2m_node(t, kd_beg(l) + first + i, n, l, first + i, P);
}

__global__ void __launch_bounds__(256) g_m2m_top_kernel(TreeData t, int64_t n, int lhi, int llo, int P, int r, int g)
{
	for (int l = lhi; l >= llo; --l)
	{
		const int first = l >= g ? r << (l - g) : 0, count = l >= g ? 1 << (l - g) : 1 << l;
		for (int i = threadIdx.x; i < count; i += blockDim.x)
			g_m2m_node(t, kd_beg(l) + first + i, n, l, first + i, P);
		__syncthreads();
	}
}

__global__ void __launch_bounds__(128)
g_m2l_kernel(TreeData t, const int2 *__restrict__ list, const unsigned *__restrict__ count, unsigned cap, float eps2, int P)
{
	const unsigned npairs = min(*count, cap);
	const int offL = g_trl_off(P + 1);
	for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < npairs; w += gridDim.x * blockDim.x)
	{
		int2 np = list[w];
		const int flags = (np.x >> kFlagShift) & 3;
		np.x &= kNodeMask;
		const float4 c1 = node_center(t, np.x), c2 = node_center(t, np.y);
		float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
		const float rinv = 1.f / sqrtf(dx*dx + dy*dy + dz*dz + eps2);
		dx *= rinv; dy *= rinv; dz *= rinv;
		float Lq[kMaxTrl];
		for (int dir = 0; dir < 2; ++dir)
		{
			if (!((flags >> dir) & 1)) continue;
			const int tgt = dir ? np.y : np.x, src = dir ? np.x : np.y;
			const float s = dir ? -1.f : 1.f;
			for (int k = 0; k < offL; ++k) Lq[k] = 0.f;
			g_m2l_acc(Lq, node_mpole(t, src), P, s * dx, s * dy, s * dz, rinv);
			float *dst = t.local + (int64_t)tgt * t.sL;
			for (int k = 1; k < offL; ++k) atomicAdd(dst + k, Lq[k]);
		}
	}
}

__device__ void g_l2l_node(const TreeData &t, int child, int P)
{
	const int parent = (child - 1) >> 1, offL = g_trl_off(P + 1);
	const float4 cp = t.center[parent], cc = t.center[child];
	float S[kMaxSym], Lc[kMaxTrl];
	float *dst = t.local + (int64_t)child * t.sL;
	g_local_expand(S, t.local + (int64_t)parent * t.sL, P);
	for (int k = 0; k < offL; ++k) Lc[k] = dst[k];
	g_l2l_acc(Lc, S, P, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
	for (int k = 1; k < offL; ++k) dst[k] = Lc[k];
}

__global__ void __launch_bounds__(128) g_l2l_level_kernel(TreeData t, int lchild, int first, int count, int P)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) g_l2l_node(t, kd_beg(lchild) + first + i, P);
}

__global__ void __launch_bounds__(256) g_l2l_top_kernel(TreeData t, int lfirst, int llast, int P)
{
	for (int l = lfirst; l <= llast; ++l)
	{
		for (int i = threadIdx.x; i < (1 << l); i += blockDim.x)
			g_l2l_node(t, kd_beg(l) + i, P);
		__syncthreads();
	}
}

__global__ void __launch_bounds__(128)
g_l2p_kernel(TreeData t, const float *__restrict__ spos, const float *__restrict__ acc_near, float *__restrict__ acc_out,
             const int *__restrict__ perm_or_null, const float *__restrict__ param, int fuse_elastic, int64_t n, int L, int P, int64_t j_lo, int64_t j_hi, float eps2, int coll)
{
	const float scale = param ? param[0] : 1.f;
	float k3[3] = {1.f, 1.f, 1.f};
	if (fuse_elastic && param) { k3[0] = param[3]; k3[1] = param[4]; k3[2] = param[5]; }
	const int beg = kd_beg(L);
	const unsigned long long magic = ~0ull / (unsigned long long)n; // floor((2^64 - 1) / n) <= 2^64 / n: never overshoots
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = j_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_hi; j += stride)
	{
		const int leaf = owner_of(j, n, L, magic);
		const float4 c = t.center[beg + leaf];
		float S[kMaxSym];
		g_local_expand(S, t.local + (int64_t)(beg + leaf) * t.sL, P);
		const float x = spos[3*j], y = spos[3*j+1], z = spos[3*j+2];
		float f[3];
		g_l2p_field(f, S, P, x - c.x, y - c.y, z - c.z);
		if (coll) self_p2p(f, spos, j, leaf, x, y, z, n, L, eps2);
		float ax = (acc_near[3*j] + f[0]) * scale, ay = (acc_near[3*j+1] + f[1]) * scale, az = (acc_near[3*j+2] + f[2]) * scale;
		if (fuse_elastic) { ax = fmaf(-k3[0], x, ax); ay = fmaf(-k3[1], y, ay); az = fmaf(-k3[2], z, az); }
		const int64_t o = perm_or_null ? (int64_t)perm_or_null[j] : j;
		acc_out[3*o] = ax; acc_out[3*o+1] = ay; acc_out[3*o+2] = az;
	}
}

constexpr int kTopLevels = 7;

template <int P>
struct GenImpl
{
	static void upward(nbco_ctx *ctx, TreeData t, const float *spos, int64_t n, int L, int r, int g, int part)
	{
		cudaStream_t st = ctx->stream;
		if (part == 1)
		{
			if (g > 0) { g_m2m_top_kernel<<<1, 256, 0, st>>>(t, n, g - 1, 0, P, r, g); ++ctx->launches; }
			return;
		}
		const int first = r << (L - g), count = 1 << (L - g);
		g_leaf_p2m_kernel<<<grid_for(count, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, n, L, P, first, count); ++ctx->launches;
		for (int l = L - 1; l > kTopLevels && l >= g; --l)
		{
			const int cnt = 1 << (l - g);
			g_m2m_level_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(t, n, l, P, r << (l - g), cnt); ++ctx->launches;
		}
		if (std::min(L - 1, kTopLevels) >= g) { g_m2m_top_kernel<<<1, 256, 0, st>>>(t, n, std::min(L - 1, kTopLevels), g, P, r, g); ++ctx->launches; }
	}
	static void m2l(nbco_ctx *ctx, TreeData t, const int2 *list, const unsigned *count, unsigned cap, float eps2)
	{
		g_m2l_kernel<<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(t, list, count, cap, eps2, P); ++ctx->launches;
	}
	static void downward(nbco_ctx *ctx, TreeData t, const float *spos, float *acc_near, float *acc_out,
	                     const int *perm_or_null, const float *param, int fuse_elastic, int64_t n, int L, int r, int g, float eps2, int coll, cudaEvent_t ev_l2p,
	                     const CsrView *)
	{
		cudaStream_t st = ctx->stream;
		if (L >= 2)
		{
			g_l2l_top_kernel<<<1, 256, 0, st>>>(t, 2, std::min(L, kTopLevels + 1), P); ++ctx->launches;
			for (int l = kTopLevels + 2; l <= L; ++l)
			{
				const int first = l >= g ? r << (l - g) : r >> (g - l), count = l >= g ? 1 << (l - g) : 1;
				g_l2l_level_kernel<<<(count + 127) / 128, 128, 0, st>>>(t, l, first, count, P); ++ctx->launches;
			}
		}
		if (ev_l2p) cudaEventRecord(ev_l2p, st);
		const int64_t j_lo = seg_start(n, r, g), j_hi = seg_start(n, r + 1, g);
		g_l2p_kernel<<<grid_for(j_hi - j_lo, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, acc_near, acc_out, perm_or_null,
		                                                                            param, fuse_elastic, n, L, P, j_lo, j_hi, eps2, coll);
		++ctx->launches;
	}
};

} // namespace

#define NBCO_GENERIC_ORDER(P) \
	extern const OrderOps kOrderOps##P; \
	const OrderOps kOrderOps##P = {GenImpl<P>::upward, GenImpl<P>::m2l, GenImpl<P>::downward, 0, nullptr, nullptr};

#ifndef NBCO_STATIC_ORDER_MAX
#define NBCO_STATIC_ORDER_MAX 5
#endif
#if NBCO_STATIC_ORDER_MAX < 4
NBCO_GENERIC_ORDER(4)
#endif
#if NBCO_STATIC_ORDER_MAX < 5
NBCO_GENERIC_ORDER(5)
#endif
#if NBCO_STATIC_ORDER_MAX < 6
NBCO_GENERIC_ORDER(6)
#endif
// orders 7..NBCO_MAX_ORDER always run the generic loops (the reference switches to its runtime m2l_acc3 /
// *_kdtree2 kernels from p = 7 too, fmm_cart3_kdtree.cuh:376-381,673-704); `nbco3 -test` sweeps p = 1..10 like main3.cu:790-811
#if NBCO_STATIC_ORDER_MAX < 7
NBCO_GENERIC_ORDER(7)
#endif
#if NBCO_STATIC_ORDER_MAX < 8
NBCO_GENERIC_ORDER(8)
#endif
NBCO_GENERIC_ORDER(9)
NBCO_GENERIC_ORDER(10)
static_assert(NBCO_MAX_ORDER == 10, "instantiate the generic orders up to NBCO_MAX_ORDER");

} // namespace nbco
