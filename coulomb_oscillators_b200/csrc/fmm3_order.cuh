// fmm3_order.cuh -- the order-templated passes of the kd-tree FMM (P2M, M2M, M2L, L2L, L2P) and
// their launchers.  Included by fmm3_p<N>.cu, which instantiates exactly one order each.
#pragma once
#include "fmm3_common.cuh"
#include "fmm_ops.cuh"

namespace nbco {

using namespace ops;

namespace {

// node tuples are padded to a multiple of 4 floats and 16-byte aligned: move them as float4
template <int N> constexpr int pad4() { return (N + 3) & ~3; }

template <int N>
__device__ __forceinline__ void load_tuple(float *dst /* pad4<N>() */, const float *__restrict__ src)
{
#pragma unroll
	for (int k = 0; k < pad4<N>() / 4; ++k)
	{
		const float4 v = reinterpret_cast<const float4 *>(src)[k];
		dst[4*k] = v.x; dst[4*k+1] = v.y; dst[4*k+2] = v.z; dst[4*k+3] = v.w;
	}
}

template <int N>
__device__ __forceinline__ void store_tuple(float *__restrict__ dst, const float *src /* pad4<N>() */)
{
#pragma unroll
	for (int k = 0; k < pad4<N>() / 4; ++k)
		reinterpret_cast<float4 *>(dst)[k] = make_float4(src[4*k], src[4*k+1], src[4*k+2], src[4*k+3]);
}

// vector reduction into a node tuple (red.global.add.v4.f32, sm_90+)
template <int N>
__device__ __forceinline__ void atomic_add_tuple(float *__restrict__ dst, const float *src /* pad4<N>() */)
{
#pragma unroll
	for (int k = 0; k < pad4<N>() / 4; ++k)
		atomicAdd(reinterpret_cast<float4 *>(dst) + k, make_float4(src[4*k], src[4*k+1], src[4*k+2], src[4*k+3]));
}

// =====================================================================================
//  upward pass
// =====================================================================================
// leaf centres (centerLeaves_krnl, appel.cuh:226-243: sequential fp32 mean) + P2M (:231-250)
template <int P>
__global__ void __launch_bounds__(128)
leaf_p2m_kernel(TreeData t, const float *__restrict__ spos, int64_t n, int L, int first, int count)
{
	const int beg = kd_beg(L);
	for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < first + count; i += gridDim.x * blockDim.x)
	{
		int64_t st = seg_start(n, i, L);
		int cnt = (int)(seg_start(n, i + 1, L) - st);
		const float *p = spos + 3 * st;
		float cx = 0.f, cy = 0.f, cz = 0.f;
		for (int j = 0; j < cnt; ++j) { cx += p[3*j]; cy += p[3*j+1]; cz += p[3*j+2]; }
		if (cnt > 0) { float f = (float)cnt; cx = __fdiv_rn(cx, f); cy = __fdiv_rn(cy, f); cz = __fdiv_rn(cz, f); }
		t.center[beg + i] = make_float4(cx, cy, cz, t.size2[beg + i]);
		float M[pad4<sym_off(P)>()];
#pragma unroll
		for (int k = 0; k < pad4<sym_off(P)>(); ++k) M[k] = 0.f;
		if constexpr (P >= 3)
			for (int j = 0; j < cnt; ++j)
				p2m_acc<P>(M, p[3*j] - cx, p[3*j+1] - cy, p[3*j+2] - cz);
		M[0] = (float)cnt;
		store_tuple<sym_off(P)>(t.mpole + (int64_t)(beg + i) * t.sM, M);
	}
}

// Uniform leaves (n a multiple of 2^L: every power-of-two N) of exactly 8 particles: a leaf is 96 contiguous, 16-byte
// aligned bytes, moved as six float4 instead of 24 scalar loads (the scalar version is bound by L1 wavefronts: ncu,
// profiles/r01_notes.md); the centre is still the sequential fp32 sum in storage order (bit-exact).
template <int P>
__global__ void __launch_bounds__(128)
leaf_p2m_u8_kernel(TreeData t, const float *__restrict__ spos, int L, int first, int count)
{
	const int beg = kd_beg(L);
	for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < first + count; i += gridDim.x * blockDim.x)
	{
		float c[24];
		const float4 *src = reinterpret_cast<const float4 *>(spos + 24 * (int64_t)i);
#pragma unroll
		for (int k = 0; k < 6; ++k) { const float4 v = src[k]; c[4*k] = v.x; c[4*k+1] = v.y; c[4*k+2] = v.z; c[4*k+3] = v.w; }
		float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
		for (int j = 0; j < 8; ++j) { cx += c[3*j]; cy += c[3*j+1]; cz += c[3*j+2]; }
		cx = __fdiv_rn(cx, 8.f); cy = __fdiv_rn(cy, 8.f); cz = __fdiv_rn(cz, 8.f);
		t.center[beg + i] = make_float4(cx, cy, cz, t.size2[beg + i]);
		float M[pad4<sym_off(P)>()];
#pragma unroll
		for (int k = 0; k < pad4<sym_off(P)>(); ++k) M[k] = 0.f;
		if constexpr (P >= 3)
		{
#pragma unroll
			for (int j = 0; j < 8; ++j) p2m_acc<P>(M, c[3*j] - cx, c[3*j+1] - cy, c[3*j+2] - cz);
		}
		M[0] = 8.f;
		store_tuple<sym_off(P)>(t.mpole + (int64_t)(beg + i) * t.sM, M);
	}
}

// same result with G lanes per leaf, for leaves of more than 8 particles (orders >= 4): particles are
// loaded one per lane, the centre is still the SEQUENTIAL fp32 sum (lane 0 adds the shuffled values in
// storage order, bit-exact), the P2M sums are accumulated per lane and reduced by butterfly shuffles
template <int P, int G>
__global__ void __launch_bounds__(128)
leaf_p2m_group_kernel(TreeData t, const float *__restrict__ spos, int64_t n, int L, int first, int count)
{
	const int beg = kd_beg(L);
	const int lane = threadIdx.x & (G - 1);
	const int groups = (gridDim.x * blockDim.x) / G;
	const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
	for (int i = first + (blockIdx.x * blockDim.x + threadIdx.x) / G; i < first + count; i += groups)
	{
		const int64_t st = seg_start(n, i, L);
		const int cnt = (int)(seg_start(n, i + 1, L) - st);
		const float *p = spos + 3 * st;
		float cx = 0.f, cy = 0.f, cz = 0.f;
		for (int j0 = 0; j0 < cnt; j0 += G)
		{
			const int j = j0 + lane;
			const float x = j < cnt ? p[3*j] : 0.f, y = j < cnt ? p[3*j+1] : 0.f, z = j < cnt ? p[3*j+2] : 0.f;
			const int nj = min(G, cnt - j0);
			for (int k = 0; k < nj; ++k)
			{
				cx += __shfl_sync(gmask, x, k, G); cy += __shfl_sync(gmask, y, k, G); cz += __shfl_sync(gmask, z, k, G);
			}
		}
		if (cnt > 0) { const float f = (float)cnt; cx = __fdiv_rn(cx, f); cy = __fdiv_rn(cy, f); cz = __fdiv_rn(cz, f); }
		float M[pad4<sym_off(P)>()];
#pragma unroll
		for (int k = 0; k < pad4<sym_off(P)>(); ++k) M[k] = 0.f;
		if constexpr (P >= 3)
		{
			for (int j = lane; j < cnt; j += G)
				p2m_acc<P>(M, p[3*j] - cx, p[3*j+1] - cy, p[3*j+2] - cz);
#pragma unroll
			for (int k = sym_off(2); k < sym_off(P); ++k)
#pragma unroll
				for (int o = G / 2; o > 0; o >>= 1)
					M[k] += __shfl_xor_sync(gmask, M[k], o, G);
		}
		if (lane == 0)
		{
			M[0] = (float)cnt;
			t.center[beg + i] = make_float4(cx, cy, cz, t.size2[beg + i]);
			store_tuple<sym_off(P)>(t.mpole + (int64_t)(beg + i) * t.sM, M);
		}
	}
}

// one parent from its two children (fmm_buildTree3_kdtree2_krnl, :328-368)
template <int P>
__device__ __forceinline__ void m2m_node(const TreeData &t, int node, int64_t n, int l, int i)
{
	const int c0 = 2*node + 1, c1 = c0 + 1;
	const float m0 = (float)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
	const float m1 = (float)(seg_start(n, 2*i + 2, l + 1) - seg_start(n, 2*i + 1, l + 1));
	const float mt = (float)(seg_start(n, i + 1, l) - seg_start(n, i, l));
	const float4 a = node_center(t, c0), b = node_center(t, c1); // level-g children of the replicated top live on their owners
	// coord = (m0*c0 + m1*c1) / m with separately rounded products (host arithmetic, :345-348)
	float cx = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.x), __fmul_rn(m1, b.x)), mt);
	float cy = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.y), __fmul_rn(m1, b.y)), mt);
	float cz = __fdiv_rn(__fadd_rn(__fmul_rn(m0, a.z), __fmul_rn(m1, b.z)), mt);
	float M[pad4<sym_off(P)>()];
#pragma unroll
	for (int k = 0; k < pad4<sym_off(P)>(); ++k) M[k] = 0.f;
	if constexpr (P >= 3)
	{
		float Mi[pad4<sym_off(P)>()];
		load_tuple<sym_off(P)>(Mi, node_mpole(t, c0));
		m2m_acc<P>(M, Mi, cx - a.x, cy - a.y, cz - a.z);
		load_tuple<sym_off(P)>(Mi, node_mpole(t, c1));
		m2m_acc<P>(M, Mi, cx - b.x, cy - b.y, cz - b.z);
	}
	M[0] = mt;
	store_tuple<sym_off(P)>(t.mpole + (int64_t)node * t.sM, M);
	t.center[node] = make_float4(cx, cy, cz, t.size2[node]);
}

template <int P>
__global__ void __launch_bounds__(128) m2m_level_kernel(TreeData t, int64_t n, int l, int first, int count)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) m2m_node<P>(t, kd_beg(l) + first + i, n, l, first + i);
}

// levels lhi .. llo (lhi >= llo) of the subtree under node (llo, first + blockIdx.x), one CTA per subtree: the
// upward pass is a chain of dependent levels and, launched level by level, is bound by launch latency
template <int P>
__global__ void __launch_bounds__(128) m2m_sub_kernel(TreeData t, int64_t n, int lhi, int llo, int first)
{
	const int root = first + blockIdx.x;
	for (int l = lhi; l >= llo; --l)
	{
		const int cnt = 1 << (l - llo), f = root << (l - llo);
		for (int i = threadIdx.x; i < cnt; i += blockDim.x)
			m2m_node<P>(t, kd_beg(l) + f + i, n, l, f + i);
		__syncthreads();
	}
}

// levels lhi .. llo in one CTA (few nodes, dependent launches otherwise); of a level l >= g only the nodes
// of rank r's subtree
template <int P>
__global__ void __launch_bounds__(256) m2m_top_kernel(TreeData t, int64_t n, int lhi, int llo, int r, int g)
{
	for (int l = lhi; l >= llo; --l)
	{
		const int first = l >= g ? r << (l - g) : 0, count = l >= g ? 1 << (l - g) : 1 << l;
		for (int i = threadIdx.x; i < count; i += blockDim.x)
			m2m_node<P>(t, kd_beg(l) + first + i, n, l, first + i);
		__syncthreads();
	}
}

// =====================================================================================
//  M2L (replaces fmm_c2c3_kdtree / _kdtree2, :613-750)
// =====================================================================================
template <int P>
__global__ void __launch_bounds__(128)
m2l_kernel(TreeData t, const int2 *__restrict__ list, const u32 *__restrict__ count, u32 cap, float eps2)
{
	const u32 npairs = min(*count, cap);
	for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < npairs; w += gridDim.x * blockDim.x)
	{
		int2 np = list[w];
		const int flags = (np.x >> kFlagShift) & 3; // bit 0: np.x is a target of this rank, bit 1: np.y
		np.x &= kNodeMask;
		const float4 c1 = node_center(t, np.x), c2 = node_center(t, np.y);
		float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z;
		const float r2 = dx*dx + dy*dy + dz*dz + eps2;
		const float rinv = 1.f / sqrtf(r2); // IEEE sqrt and divide like the reference (:636-637); M2L is not flop-bound
		dx *= rinv; dy *= rinv; dz *= rinv;
		float M[pad4<sym_off(P)>()], Lq[pad4<trl_off(P + 1)>()];
		// target np.x, source np.y
		if (flags & 1)
		{
			load_tuple<sym_off(P)>(M, node_mpole(t, np.y));
			m2l_weight<P>(M);
#pragma unroll
			for (int k = 0; k < pad4<trl_off(P + 1)>(); ++k) Lq[k] = 0.f;
			m2l_acc_w<P>(Lq, M, dx, dy, dz, rinv);
			atomic_add_tuple<trl_off(P + 1)>(t.local + (int64_t)np.x * t.sL, Lq); // slot 0 receives +0
		}
		// target np.y, source np.x
		if (flags & 2)
		{
			load_tuple<sym_off(P)>(M, node_mpole(t, np.x));
			m2l_weight<P>(M);
#pragma unroll
			for (int k = 0; k < pad4<trl_off(P + 1)>(); ++k) Lq[k] = 0.f;
			m2l_acc_w<P>(Lq, M, -dx, -dy, -dz, rinv);
			atomic_add_tuple<trl_off(P + 1)>(t.local + (int64_t)np.y * t.sL, Lq); // slot 0 receives +0
		}
	}
}

// =====================================================================================
//  by-target downward pass (round 2): M2L gathered per TARGET node inside the L2L levels
// =====================================================================================
// Lq += sum over the sources row[sub], row[sub + G], ... of the target's M2L row (sorted by source id)
template <int P>
__device__ __forceinline__ void m2l_gather(const TreeData &t, const CsrView &csr, int node, const float4 &ct, float *Lq, int sub, int G, float eps2)
{
	const u32 e1 = min(csr.off[node + 1], csr.cap);
	for (u32 e = csr.off[node] + sub; e < e1; e += G)
	{
		const int sn = csr.src[e];
		const float4 cs = node_center(t, sn);
		float dx = ct.x - cs.x, dy = ct.y - cs.y, dz = ct.z - cs.z; // c_target - c_source
		const float r2 = dx*dx + dy*dy + dz*dz + eps2;
		const float rinv = 1.f / sqrtf(r2); // IEEE sqrt and divide like the reference (:636-637)
		dx *= rinv; dy *= rinv; dz *= rinv;
		float M[pad4<sym_off(P)>()];
		load_tuple<sym_off(P)>(M, node_mpole(t, sn));
		m2l_weight<P>(M);
		m2l_acc_w<P>(Lq, M, dx, dy, dz, rinv);
	}
}

// one node of the downward pass: local = (M2L sum of its row) + (shift of the parent's local); stored once
template <int P>
__device__ __forceinline__ void down_node(const TreeData &t, const CsrView &csr, int child, float eps2, bool has_parent)
{
	const float4 cc = t.center[child];
	float Lc[pad4<trl_off(P + 1)>()];
#pragma unroll
	for (int k = 0; k < pad4<trl_off(P + 1)>(); ++k) Lc[k] = 0.f;
	m2l_gather<P>(t, csr, child, cc, Lc, 0, 1, eps2);
	if (has_parent)
	{
		const int parent = (child - 1) >> 1;
		const float4 cp = t.center[parent];
		float Lp[pad4<trl_off(P + 1)>()], S[sym_off(P + 1)];
		load_tuple<trl_off(P + 1)>(Lp, t.local + (int64_t)parent * t.sL);
		S[0] = 0.f;
		local_expand<P>(S, Lp);
		l2l_acc<P>(Lc, S, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
	}
	Lc[0] = 0.f;
	store_tuple<trl_off(P + 1)>(t.local + (int64_t)child * t.sL, Lc);
}

template <int P>
__global__ void __launch_bounds__(128) down_level_kernel(TreeData t, CsrView csr, int lchild, int first, int count, float eps2)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) down_node<P>(t, csr, kd_beg(lchild) + first + i, eps2, lchild >= 2);
}

// wide levels, tuples of 16 floats (P <= 3): FOUR lanes per node.  The lanes share the node's M2L row (sources sub,
// sub + 4, ...), the partial sums are reduced by butterfly shuffles, every lane moves one float4 of the parent's tuple and
// stores one float4 of the result (fully used sectors)
template <int P>
__global__ void __launch_bounds__(128) down_level4_kernel(TreeData t, CsrView csr, int lchild, int first, int count, float eps2)
{
	static_assert(pad4<trl_off(P + 1)>() == 16, "four float4 per tuple");
	const int gt = blockIdx.x * blockDim.x + threadIdx.x, i = gt >> 2, sub = gt & 3;
	if (i >= count) return; // a whole group of four leaves together
	const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
	const int child = kd_beg(lchild) + first + i, parent = (child - 1) >> 1;
	const float4 cc = t.center[child];
	float Lc[16];
#pragma unroll
	for (int k = 0; k < 16; ++k) Lc[k] = 0.f;
	m2l_gather<P>(t, csr, child, cc, Lc, sub, 4, eps2);
#pragma unroll
	for (int k = 1; k < trl_off(P + 1); ++k)
	{
		Lc[k] += __shfl_xor_sync(gmask, Lc[k], 1, 4);
		Lc[k] += __shfl_xor_sync(gmask, Lc[k], 2, 4);
	}
	if (lchild >= 2)
	{
		const float4 qp = reinterpret_cast<const float4 *>(t.local + (int64_t)parent * t.sL)[sub];
		const float4 cp = t.center[parent];
		float Lp[16], S[sym_off(P + 1)];
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			Lp[4*q]   = __shfl_sync(gmask, qp.x, q, 4); Lp[4*q+1] = __shfl_sync(gmask, qp.y, q, 4);
			Lp[4*q+2] = __shfl_sync(gmask, qp.z, q, 4); Lp[4*q+3] = __shfl_sync(gmask, qp.w, q, 4);
		}
		S[0] = 0.f;
		local_expand<P>(S, Lp);
		l2l_acc<P>(Lc, S, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
	}
	Lc[0] = 0.f;
	float4 out;
	out.x = sub == 0 ? Lc[0] : (sub == 1 ? Lc[4] : (sub == 2 ? Lc[8]  : Lc[12]));
	out.y = sub == 0 ? Lc[1] : (sub == 1 ? Lc[5] : (sub == 2 ? Lc[9]  : Lc[13]));
	out.z = sub == 0 ? Lc[2] : (sub == 1 ? Lc[6] : (sub == 2 ? Lc[10] : Lc[14]));
	out.w = sub == 0 ? Lc[3] : (sub == 1 ? Lc[7] : (sub == 2 ? Lc[11] : Lc[15]));
	reinterpret_cast<float4 *>(t.local + (int64_t)child * t.sL)[sub] = out;
}

// child levels lfirst .. llast of the subtree under node (lfirst - 1, first + blockIdx.x), one CTA per subtree
template <int P>
__global__ void __launch_bounds__(128) down_sub_kernel(TreeData t, CsrView csr, int lfirst, int llast, int first, float eps2)
{
	const int root = first + blockIdx.x, lroot = lfirst - 1;
	for (int l = lfirst; l <= llast; ++l)
	{
		const int cnt = 1 << (l - lroot), f = root << (l - lroot);
		for (int i = threadIdx.x; i < cnt; i += blockDim.x)
			down_node<P>(t, csr, kd_beg(l) + f + i, eps2, l >= 2);
		__syncthreads();
	}
}

// levels 0 .. llast in one CTA (levels 0 and 1 have empty rows and no parent: their locals become zero)
template <int P>
__global__ void __launch_bounds__(256) down_top_kernel(TreeData t, CsrView csr, int llast, float eps2)
{
	for (int l = 0; l <= llast; ++l)
	{
		for (int i = threadIdx.x; i < (1 << l); i += blockDim.x)
			down_node<P>(t, csr, kd_beg(l) + i, eps2, l >= 2);
		__syncthreads();
	}
}

// L2P + near field gathered per TARGET leaf (own leaf + the leaves of its P2P row) + rescale + optional elastic term +
// optional un-sort, one pass over the particles: every acceleration is one sum in registers, stored once
template <int P>
__global__ void __launch_bounds__(128)
l2p_near_kernel(TreeData t, CsrView csr, const float *__restrict__ spos, float *__restrict__ acc_out,
                const int *__restrict__ perm_or_null, const float *__restrict__ param, int fuse_elastic, int64_t n, int L, int64_t j_lo, int64_t j_hi, float eps2, int coll)
{
	const float scale = param ? param[0] : 1.f;
	float k3[3] = {1.f, 1.f, 1.f};
	if (fuse_elastic && param) { k3[0] = param[3]; k3[1] = param[4]; k3[2] = param[5]; }
	const int beg = kd_beg(L);
	const unsigned long long magic = ~0ull / (unsigned long long)n;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = j_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_hi; j += stride)
	{
		const int leaf = owner_of(j, n, L, magic);
		const float4 c = t.center[beg + leaf];
		float Lq[pad4<trl_off(P + 1)>()], S[sym_off(P + 1)];
		load_tuple<trl_off(P + 1)>(Lq, t.local + (int64_t)(beg + leaf) * t.sL);
		S[0] = 0.f;
		local_expand<P>(S, Lq);
		const float x = spos[3*j], y = spos[3*j+1], z = spos[3*j+2];
		float f[3];
		l2p_field<P>(f, S, x - c.x, y - c.y, z - c.z);
		if (coll)
		{
			float fn[3] = {0.f, 0.f, 0.f};
			self_p2p(fn, spos, j, leaf, x, y, z, n, L, eps2);
			const u32 e1 = min(csr.off[csr.ntot + leaf + 1], csr.cap);
			for (u32 e = csr.off[csr.ntot + leaf]; e < e1; ++e)
			{
				const int sl = csr.src[e] - beg;
				const int64_t s0 = seg_start(n, sl, L);
				const int cnt = (int)(seg_start(n, sl + 1, L) - s0);
				// multi-GPU: the particles of a remote source leaf are read from their owner's published positions
				const float *__restrict__ sp = spos;
				if (t.peers.g > 0)
				{
					const int o = sl >> (L - t.peers.g);
					if (o != t.peers.me) sp = t.peers.pos[o];
				}
				leaf_p2p(fn, sp + 3 * s0, cnt, x, y, z, eps2);
			}
			f[0] += fn[0]; f[1] += fn[1]; f[2] += fn[2];
		}
		float ax = f[0] * scale, ay = f[1] * scale, az = f[2] * scale;
		if (fuse_elastic) { ax = fmaf(-k3[0], x, ax); ay = fmaf(-k3[1], y, ay); az = fmaf(-k3[2], z, az); }
		const int64_t o = perm_or_null ? (int64_t)perm_or_null[j] : j;
		acc_out[3*o] = ax; acc_out[3*o+1] = ay; acc_out[3*o+2] = az;
	}
}

// =====================================================================================
//  downward pass (replaces fmm_pushl3_kdtree*, fmm_pushLeaves3_kdtree*, rescale, add_elastic)
// =====================================================================================
template <int P>
__device__ __forceinline__ void l2l_node(const TreeData &t, int child)
{
	const int parent = (child - 1) >> 1;
	const float4 cp = t.center[parent], cc = t.center[child];
	float Lp[pad4<trl_off(P + 1)>()], S[sym_off(P + 1)], Lc[pad4<trl_off(P + 1)>()];
	float *dst = t.local + (int64_t)child * t.sL;
	load_tuple<trl_off(P + 1)>(Lp, t.local + (int64_t)parent * t.sL);
	load_tuple<trl_off(P + 1)>(Lc, dst);
	S[0] = 0.f;
	local_expand<P>(S, Lp);
	l2l_acc<P>(Lc, S, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
	store_tuple<trl_off(P + 1)>(dst, Lc);
}

// children of level l (i.e. nodes of level l+1) pull from their parents
template <int P>
__global__ void __launch_bounds__(128) l2l_level_kernel(TreeData t, int lchild, int first, int count)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) l2l_node<P>(t, kd_beg(lchild) + first + i);
}

// wide levels, tuples of 16 floats (P <= 3): FOUR lanes per node, each moves one float4 of the parent's and of the child's
// tuple (every request touches fully used sectors; with a thread per node each request touches 32 half-used ones and
// the kernel is bound by memory latency: ncu issue active 6 %, profiles/r01_notes.md); the quads are exchanged by
// shuffles and the shift is computed by all four lanes, each storing its own quad
template <int P>
__global__ void __launch_bounds__(128) l2l_level4_kernel(TreeData t, int lchild, int first, int count)
{
	static_assert(pad4<trl_off(P + 1)>() == 16, "four float4 per tuple");
	const int gt = blockIdx.x * blockDim.x + threadIdx.x, i = gt >> 2, sub = gt & 3;
	if (i >= count) return; // a whole group of four leaves together
	const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
	const int child = kd_beg(lchild) + first + i, parent = (child - 1) >> 1;
	float4 *dst = reinterpret_cast<float4 *>(t.local + (int64_t)child * t.sL);
	const float4 qp = reinterpret_cast<const float4 *>(t.local + (int64_t)parent * t.sL)[sub];
	const float4 qc = dst[sub];
	const float4 cp = t.center[parent], cc = t.center[child];
	float Lp[16], Lc[16], S[sym_off(P + 1)];
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		Lp[4*q]   = __shfl_sync(gmask, qp.x, q, 4); Lp[4*q+1] = __shfl_sync(gmask, qp.y, q, 4);
		Lp[4*q+2] = __shfl_sync(gmask, qp.z, q, 4); Lp[4*q+3] = __shfl_sync(gmask, qp.w, q, 4);
		Lc[4*q]   = __shfl_sync(gmask, qc.x, q, 4); Lc[4*q+1] = __shfl_sync(gmask, qc.y, q, 4);
		Lc[4*q+2] = __shfl_sync(gmask, qc.z, q, 4); Lc[4*q+3] = __shfl_sync(gmask, qc.w, q, 4);
	}
	S[0] = 0.f;
	local_expand<P>(S, Lp);
	l2l_acc<P>(Lc, S, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
	float4 out;
	out.x = sub == 0 ? Lc[0] : (sub == 1 ? Lc[4] : (sub == 2 ? Lc[8]  : Lc[12]));
	out.y = sub == 0 ? Lc[1] : (sub == 1 ? Lc[5] : (sub == 2 ? Lc[9]  : Lc[13]));
	out.z = sub == 0 ? Lc[2] : (sub == 1 ? Lc[6] : (sub == 2 ? Lc[10] : Lc[14]));
	out.w = sub == 0 ? Lc[3] : (sub == 1 ? Lc[7] : (sub == 2 ? Lc[11] : Lc[15]));
	dst[sub] = out;
}

// child levels lfirst .. llast of the subtree under node (lfirst - 1, first + blockIdx.x), one CTA per subtree
template <int P>
__global__ void __launch_bounds__(128) l2l_sub_kernel(TreeData t, int lfirst, int llast, int first)
{
	const int root = first + blockIdx.x, lroot = lfirst - 1;
	for (int l = lfirst; l <= llast; ++l)
	{
		const int cnt = 1 << (l - lroot), f = root << (l - lroot);
		for (int i = threadIdx.x; i < cnt; i += blockDim.x)
			l2l_node<P>(t, kd_beg(l) + f + i);
		__syncthreads();
	}
}

template <int P>
__global__ void __launch_bounds__(256) l2l_top_kernel(TreeData t, int lfirst, int llast)
{
	for (int l = lfirst; l <= llast; ++l)
	{
		for (int i = threadIdx.x; i < (1 << l); i += blockDim.x)
			l2l_node<P>(t, kd_beg(l) + i);
		__syncthreads();
	}
}

// L2P + rescale + optional elastic term + optional un-sort, one pass over the particles
template <int P>
__global__ void __launch_bounds__(128)
l2p_kernel(TreeData t, const float *__restrict__ spos, const float *__restrict__ acc_near, float *__restrict__ acc_out,
           const int *__restrict__ perm_or_null, const float *__restrict__ param, int fuse_elastic, int64_t n, int L, int64_t j_lo, int64_t j_hi, float eps2, int coll)
{
	const float scale = param ? param[0] : 1.f;
	float k3[3] = {1.f, 1.f, 1.f};
	if (fuse_elastic && param) { k3[0] = param[3]; k3[1] = param[4]; k3[2] = param[5]; }
	const int beg = kd_beg(L);
	const unsigned long long magic = ~0ull / (unsigned long long)n; // floor((2^64 - 1) / n) <= 2^64 / n: never overshoots
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = j_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_hi; j += stride)
	{
		// leaf of sorted position j: floor(2^L j / n) (:162-164)
		const int leaf = owner_of(j, n, L, magic);
		const float4 c = t.center[beg + leaf];
		float Lq[pad4<trl_off(P + 1)>()], S[sym_off(P + 1)];
		load_tuple<trl_off(P + 1)>(Lq, t.local + (int64_t)(beg + leaf) * t.sL);
		S[0] = 0.f;
		local_expand<P>(S, Lq);
		const float x = spos[3*j], y = spos[3*j+1], z = spos[3*j+2];
		float f[3];
		l2p_field<P>(f, S, x - c.x, y - c.y, z - c.z);
		if (coll) self_p2p(f, spos, j, leaf, x, y, z, n, L, eps2);
		float ax = (acc_near[3*j] + f[0]) * scale, ay = (acc_near[3*j+1] + f[1]) * scale, az = (acc_near[3*j+2] + f[2]) * scale;
		if (fuse_elastic) { ax = fmaf(-k3[0], x, ax); ay = fmaf(-k3[1], y, ay); az = fmaf(-k3[2], z, az); }
		const int64_t o = perm_or_null ? (int64_t)perm_or_null[j] : j;
		acc_out[3*o] = ax; acc_out[3*o+1] = ay; acc_out[3*o+2] = az;
	}
}


// L2P for uniform leaves of C particles (n a multiple of 2^L, C = n / 2^L in {8, 16, 32}): leaf = j / C without the
// ceil-division bookkeeping, the leaf's particles read as float4 (3 C / 4 of them, shared by the lanes of the leaf through
// L1), the intra-leaf near field fully unrolled
template <int P, int C>
__global__ void __launch_bounds__(128)
l2p_uniform_kernel(TreeData t, const float *__restrict__ spos, const float *__restrict__ acc_near, float *__restrict__ acc_out,
                   const int *__restrict__ perm_or_null, const float *__restrict__ param, int fuse_elastic, int L, int64_t j_lo, int64_t j_hi, float eps2, int coll)
{
	const float scale = param ? param[0] : 1.f;
	float k3[3] = {1.f, 1.f, 1.f};
	if (fuse_elastic && param) { k3[0] = param[3]; k3[1] = param[4]; k3[2] = param[5]; }
	const int beg = kd_beg(L);
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = j_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < j_hi; j += stride)
	{
		const int leaf = (int)(j / C), me = (int)(j % C);
		const float4 c = t.center[beg + leaf];
		float Lq[pad4<trl_off(P + 1)>()], S[sym_off(P + 1)];
		load_tuple<trl_off(P + 1)>(Lq, t.local + (int64_t)(beg + leaf) * t.sL);
		S[0] = 0.f;
		local_expand<P>(S, Lq);
		const float4 *src = reinterpret_cast<const float4 *>(spos + 3 * (int64_t)leaf * C);
		const float x = spos[3*j], y = spos[3*j+1], z = spos[3*j+2];
		float f[3];
		l2p_field<P>(f, S, x - c.x, y - c.y, z - c.z);
		if (coll)
		{
			float ax = 0.f, ay = 0.f, az = 0.f;
			// four particles per three float4
#pragma unroll
			for (int q = 0; q < C / 4; ++q)
			{
				const float4 v0 = src[3*q], v1 = src[3*q+1], v2 = src[3*q+2];
				const float sx[4] = {v0.x, v0.w, v1.z, v2.y}, sy[4] = {v0.y, v1.x, v1.w, v2.z}, sz[4] = {v0.z, v1.y, v2.x, v2.w};
#pragma unroll
				for (int k = 0; k < 4; ++k)
				{
					const float dx = x - sx[k], dy = y - sy[k], dz = z - sz[k];
					float r2 = fmaf(dx, dx, eps2);
					r2 = fmaf(dy, dy, r2);
					r2 = fmaf(dz, dz, r2);
					float w;
					asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(r2));
					w = w * fmaf(-0.5f * r2 * w, w, 1.5f);
					const float w3 = (w * w) * w;
					ax = fmaf(dx, w3, ax); ay = fmaf(dy, w3, ay); az = fmaf(dz, w3, az);
				}
			}
			(void)me;
			f[0] += ax; f[1] += ay; f[2] += az;
		}
		float ax = (acc_near[3*j] + f[0]) * scale, ay = (acc_near[3*j+1] + f[1]) * scale, az = (acc_near[3*j+2] + f[2]) * scale;
		if (fuse_elastic) { ax = fmaf(-k3[0], x, ax); ay = fmaf(-k3[1], y, ay); az = fmaf(-k3[2], z, az); }
		const int64_t o = perm_or_null ? (int64_t)perm_or_null[j] : j;
		acc_out[3*o] = ax; acc_out[3*o+1] = ay; acc_out[3*o+2] = az;
	}
}

// L2L of the LEAF level fused into L2P (uniform leaves of C particles, tuples of 16 floats, P <= 3): FOUR lanes per leaf.
// Each lane moves one float4 of the parent's and of the leaf's tuple (as l2l_level4_kernel), the four shift the parent's
// expansion to the leaf's centre in registers and every lane evaluates C / 4 of the leaf's particles.  The leaf's final
// tuple is never written: against the two-kernel form this saves one 64-byte write and one 64-byte read per leaf and
// three of every four expansions of the same tuple (nbco_fmm_get_tree finishes the leaf level on demand,
// finish_leaf_locals below).  Same operations in the same order as l2l_level4_kernel + l2p_uniform_kernel: bit-identical.
// Sparse near field: with t.nearbits the rows of acc_near are read only for the leaves the pair kernel marked, and those
// rows are zeroed again behind the output stores, so the caller never clears acc_near (OrderOps::sparse_near).
template <int P, int C>
__global__ void __launch_bounds__(128)
l2lp_uniform_kernel(TreeData t, const float *__restrict__ spos, float *__restrict__ acc_near, float *__restrict__ acc_out,
                    const int *__restrict__ perm_or_null, const float *__restrict__ param, int fuse_elastic, int L, int leaf_lo, int leaf_hi, float eps2, int coll)
{
	static_assert(pad4<trl_off(P + 1)>() == 16 && C % 4 == 0, "four float4 per tuple, C / 4 particles per lane");
	constexpr int K = C / 4; // particles per lane
	const float scale = param ? param[0] : 1.f;
	float k3[3] = {1.f, 1.f, 1.f};
	if (fuse_elastic && param) { k3[0] = param[3]; k3[1] = param[4]; k3[2] = param[5]; }
	const int beg = kd_beg(L);
	const int sub = threadIdx.x & 3;
	const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
	const int stride = (int)((gridDim.x * blockDim.x) >> 2);
	for (int leaf = leaf_lo + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 2); leaf < leaf_hi; leaf += stride) // a group of four stays together
	{
		const int child = beg + leaf, parent = (child - 1) >> 1;
		const float4 qp = reinterpret_cast<const float4 *>(t.local + (int64_t)parent * t.sL)[sub];
		const float4 qc = reinterpret_cast<const float4 *>(t.local + (int64_t)child * t.sL)[sub];
		const float4 cp = t.center[parent], cc = t.center[child];
		// my K particles: 3 K consecutive floats of the leaf's row
		const int64_t j0 = (int64_t)leaf * C + sub * K;
		float tp[3 * K], an[3 * K];
		{
			const float2 *p2 = reinterpret_cast<const float2 *>(spos + 3 * j0);
#pragma unroll
			for (int m = 0; m < 3 * K / 2; ++m) { const float2 v = p2[m]; tp[2*m] = v.x; tp[2*m+1] = v.y; }
		}
		// sparse near field: the rows of a leaf that no pair touched are zero and are neither read nor written; the rows that
		// were touched are read here and zeroed again at the END of the iteration (a store issued right behind a load of the
		// same address stalls the load/store unit until the data is back: 0.22 -> 0.62 ms, profiles/r02_notes.md)
		const bool touched = !t.nearbits || ((t.nearbits[leaf >> 5] >> (leaf & 31)) & 1u);
#pragma unroll
		for (int m = 0; m < 3 * K; ++m) an[m] = 0.f;
		if (touched)
		{
			const float2 *a2 = reinterpret_cast<const float2 *>(acc_near + 3 * j0);
#pragma unroll
			for (int m = 0; m < 3 * K / 2; ++m) { const float2 w = a2[m]; an[2*m] = w.x; an[2*m+1] = w.y; }
		}
		float Lp[16], Lc[16], S[sym_off(P + 1)];
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			Lp[4*q]   = __shfl_sync(gmask, qp.x, q, 4); Lp[4*q+1] = __shfl_sync(gmask, qp.y, q, 4);
			Lp[4*q+2] = __shfl_sync(gmask, qp.z, q, 4); Lp[4*q+3] = __shfl_sync(gmask, qp.w, q, 4);
			Lc[4*q]   = __shfl_sync(gmask, qc.x, q, 4); Lc[4*q+1] = __shfl_sync(gmask, qc.y, q, 4);
			Lc[4*q+2] = __shfl_sync(gmask, qc.z, q, 4); Lc[4*q+3] = __shfl_sync(gmask, qc.w, q, 4);
		}
		S[0] = 0.f;
		local_expand<P>(S, Lp);
		l2l_acc<P>(Lc, S, cc.x - cp.x, cc.y - cp.y, cc.z - cp.z);
		S[0] = 0.f;
		local_expand<P>(S, Lc);
		float f[K][3];
#pragma unroll
		for (int k = 0; k < K; ++k) l2p_field<P>(f[k], S, tp[3*k] - cc.x, tp[3*k+1] - cc.y, tp[3*k+2] - cc.z);
		if (coll)
		{
			float a[K][3];
#pragma unroll
			for (int k = 0; k < K; ++k) a[k][0] = a[k][1] = a[k][2] = 0.f;
			const float4 *src = reinterpret_cast<const float4 *>(spos + 3 * (int64_t)leaf * C);
#pragma unroll
			for (int q = 0; q < C / 4; ++q)
			{
				const float4 v0 = src[3*q], v1 = src[3*q+1], v2 = src[3*q+2];
				const float sx[4] = {v0.x, v0.w, v1.z, v2.y}, sy[4] = {v0.y, v1.x, v1.w, v2.z}, sz[4] = {v0.z, v1.y, v2.x, v2.w};
#pragma unroll
				for (int s = 0; s < 4; ++s)
#pragma unroll
					for (int k = 0; k < K; ++k)
					{
						const float dx = tp[3*k] - sx[s], dy = tp[3*k+1] - sy[s], dz = tp[3*k+2] - sz[s];
						float r2 = fmaf(dx, dx, eps2);
						r2 = fmaf(dy, dy, r2);
						r2 = fmaf(dz, dz, r2);
						float w;
						asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(r2));
						w = w * fmaf(-0.5f * r2 * w, w, 1.5f);
						const float w3 = (w * w) * w;
						a[k][0] = fmaf(dx, w3, a[k][0]); a[k][1] = fmaf(dy, w3, a[k][1]); a[k][2] = fmaf(dz, w3, a[k][2]);
					}
			}
#pragma unroll
			for (int k = 0; k < K; ++k) { f[k][0] += a[k][0]; f[k][1] += a[k][1]; f[k][2] += a[k][2]; }
		}
		float o3[3 * K];
#pragma unroll
		for (int k = 0; k < K; ++k)
		{
			float ax = (an[3*k] + f[k][0]) * scale, ay = (an[3*k+1] + f[k][1]) * scale, az = (an[3*k+2] + f[k][2]) * scale;
			if (fuse_elastic) { ax = fmaf(-k3[0], tp[3*k], ax); ay = fmaf(-k3[1], tp[3*k+1], ay); az = fmaf(-k3[2], tp[3*k+2], az); }
			o3[3*k] = ax; o3[3*k+1] = ay; o3[3*k+2] = az;
		}
		if (perm_or_null)
		{
#pragma unroll
			for (int k = 0; k < K; ++k)
			{
				const int64_t o = (int64_t)perm_or_null[j0 + k];
				acc_out[3*o] = o3[3*k]; acc_out[3*o+1] = o3[3*k+1]; acc_out[3*o+2] = o3[3*k+2];
			}
		}
		else
		{
			float2 *d2 = reinterpret_cast<float2 *>(acc_out + 3 * j0);
#pragma unroll
			for (int m = 0; m < 3 * K / 2; ++m) d2[m] = make_float2(o3[2*m], o3[2*m+1]);
		}
		if (touched && t.nearbits)
		{
			float2 *a2 = reinterpret_cast<float2 *>(acc_near + 3 * j0);
#pragma unroll
			for (int m = 0; m < 3 * K / 2; ++m) a2[m] = make_float2(0.f, 0.f);
		}
	}
}

constexpr int kTopLevels = 7; // levels 0..7 of the upward / 2..8 of the downward pass run in one CTA
constexpr int kWideLevel = 1 << 16; // nodes of one rank at a level from which the level gets its own launch
constexpr int kSubLevels = 7; // deeper levels: chunks of 7 levels, one CTA per subtree (128 nodes at its widest level)

template <int P>
struct OrderImpl
{
	static void upward(nbco_ctx *ctx, TreeData t, const float *spos, int64_t n, int L, int r, int g, int part)
	{
		cudaStream_t st = ctx->stream;
		if (part == 1)
		{
			if (g > 0) { m2m_top_kernel<P><<<1, 256, 0, st>>>(t, n, g - 1, 0, r, g); ++ctx->launches; }
			return;
		}
		const int mlt_max = (int)((n - 1) / (1ll << L) + 1);
		const int first = r << (L - g), count = 1 << (L - g);
		const bool uniform = (n & ((1ll << L) - 1)) == 0; // every leaf holds exactly n / 2^L particles
		if (uniform && mlt_max == 8) leaf_p2m_u8_kernel<P><<<grid_for(count, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, L, first, count);
		else if (mlt_max <= 8) leaf_p2m_kernel<P><<<grid_for(count, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, n, L, first, count);
		else if (mlt_max <= 16) leaf_p2m_group_kernel<P, 16><<<grid_for(16ll * count, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, n, L, first, count);
		else leaf_p2m_group_kernel<P, 32><<<grid_for(32ll * count, 128, ctx->sm_count, 16), 128, 0, st>>>(t, spos, n, L, first, count);
		++ctx->launches;
		// levels L-1 .. max(kTopLevels + 1, g): a level with many nodes of this rank gets its own launch (bandwidth-bound,
		// fully parallel); the smaller ones run in chunks of kSubLevels levels, one CTA per subtree of a chunk
		// (latency-bound: a launch per level would cost more than the work)
		const int lstop = std::max(kTopLevels + 1, g);
		int lhi = L - 1;
		for (; lhi >= lstop && (1 << (lhi - g)) >= kWideLevel; --lhi)
		{
			const int cnt = 1 << (lhi - g);
			m2m_level_kernel<P><<<(cnt + 127) / 128, 128, 0, st>>>(t, n, lhi, r << (lhi - g), cnt); ++ctx->launches;
		}
		for (; lhi >= lstop; lhi -= kSubLevels)
		{
			const int llo = std::max(lhi - kSubLevels + 1, lstop);
			m2m_sub_kernel<P><<<1 << (llo - g), 128, 0, st>>>(t, n, lhi, llo, r << (llo - g)); ++ctx->launches;
		}
		if (std::min(L - 1, kTopLevels) >= g) { m2m_top_kernel<P><<<1, 256, 0, st>>>(t, n, std::min(L - 1, kTopLevels), g, r, g); ++ctx->launches; }
	}
	static void m2l(nbco_ctx *ctx, TreeData t, const int2 *list, const unsigned *count, unsigned cap, float eps2)
	{
		m2l_kernel<P><<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(t, list, count, cap, eps2); ++ctx->launches; // 12 .. 24 CTAs per SM: same time (bound by the reductions in L2, profiles/r02_notes.md)
	}
	// by-target flow: same level schedule as below (top levels in one CTA, chunks of kSubLevels levels per subtree CTA
	// while a level is small, one launch per wide level), every node gathers its own M2L row
	static void downward_by_target(nbco_ctx *ctx, TreeData t, const float *spos, float *acc_out, const int *perm_or_null, const float *param,
	                               int fuse_elastic, int64_t n, int L, int r, int g, float eps2, int coll, cudaEvent_t ev_l2p, const CsrView &csr)
	{
		cudaStream_t st = ctx->stream;
		const int ltop = std::min(L, kTopLevels + 1);
		down_top_kernel<P><<<1, 256, 0, st>>>(t, csr, ltop, eps2); ++ctx->launches;
		int lf = kTopLevels + 2;
		for (; lf <= L; lf += kSubLevels)
		{
			int ll = std::min(lf + kSubLevels - 1, L);
			while (ll >= lf && ll >= g && (1 << (ll - g)) >= kWideLevel) --ll; // leave the wide levels to the loop below
			if (ll < lf) break;
			const int lroot = lf - 1;
			const int first = lroot >= g ? r << (lroot - g) : r >> (g - lroot), count = lroot >= g ? 1 << (lroot - g) : 1;
			down_sub_kernel<P><<<count, 128, 0, st>>>(t, csr, lf, ll, first, eps2); ++ctx->launches;
			if (ll < lf + kSubLevels - 1) { lf = ll + 1; break; }
		}
		for (int l = lf; l <= L; ++l)
		{
			const int first = l >= g ? r << (l - g) : r >> (g - l), count = l >= g ? 1 << (l - g) : 1;
			if constexpr (pad4<trl_off(P + 1)>() == 16)
				down_level4_kernel<P><<<(4 * count + 127) / 128, 128, 0, st>>>(t, csr, l, first, count, eps2);
			else
				down_level_kernel<P><<<(count + 127) / 128, 128, 0, st>>>(t, csr, l, first, count, eps2);
			++ctx->launches;
		}
		if (ev_l2p) cudaEventRecord(ev_l2p, st);
		const int64_t j_lo = seg_start(n, r, g), j_hi = seg_start(n, r + 1, g);
		l2p_near_kernel<P><<<grid_for(j_hi - j_lo, 128, ctx->sm_count, 16), 128, 0, st>>>(t, csr, spos, acc_out, perm_or_null, param, fuse_elastic,
		                                                                                   n, L, j_lo, j_hi, eps2, coll);
		++ctx->launches;
	}
	// the pair-list downward pass leaves the leaf level to the L2P kernel when this holds
	static bool fused_leaf_level(int64_t n, int L)
	{
		if (pad4<trl_off(P + 1)>() != 16 || L < 1 || (n & ((1ll << L) - 1)) != 0 || getenv("NBCO_NO_LEAF_FUSION")) return false; // the variable: A/B runs
		return (n >> L) == 8; // the leaf size of the reference's level rule at order 3 and power-of-two N (s = order^2); 16 and 32
		                      // particles per leaf (only reachable through max_level / dens_inhom at these orders) keep the two-kernel form
	}
	static int sparse_near(int64_t n, int L) { return fused_leaf_level(n, L) ? 1 : 0; } // l2lp_uniform_kernel honours t.nearbits
	// nbco_fmm_get_tree: push the leaf level the fused kernel kept in registers (once per evaluation, fmm3.cu keeps the flag)
	static void finish_leaf_locals(nbco_ctx *ctx, TreeData t, int64_t n, int L, int r, int g)
	{
		if (!fused_leaf_level(n, L)) return;
		if constexpr (pad4<trl_off(P + 1)>() == 16)
		{
			const int first = L >= g ? r << (L - g) : r >> (g - L), count = L >= g ? 1 << (L - g) : 1;
			l2l_level4_kernel<P><<<(4 * count + 127) / 128, 128, 0, ctx->stream>>>(t, L, first, count); ++ctx->launches;
		}
	}
	static void downward(nbco_ctx *ctx, TreeData t, const float *spos, float *acc_near, float *acc_out,
	                     const int *perm_or_null, const float *param, int fuse_elastic, int64_t n, int L, int r, int g, float eps2, int coll, cudaEvent_t ev_l2p,
	                     const CsrView *csr)
	{
		if (csr) { downward_by_target(ctx, t, spos, acc_out, perm_or_null, param, fuse_elastic, n, L, r, g, eps2, coll, ev_l2p, *csr); return; }
		cudaStream_t st = ctx->stream;
		// uniform leaves and 16-float tuples: the leaf level is pushed inside the L2P kernel (l2lp_uniform_kernel)
		const bool fused = fused_leaf_level(n, L);
		const int Lfull = L;
		if (fused) L = L - 1; // last level of the L2L schedule below
		// locals of levels 0 and 1 stay zero (nothing is ever admissible there); level l+1 pulls from level l >= 1
		if (L >= 2)
		{
			l2l_top_kernel<P><<<1, 256, 0, st>>>(t, 2, std::min(L, kTopLevels + 1)); ++ctx->launches;
			// child levels kTopLevels + 2 .. L: chunks of kSubLevels levels (one CTA per subtree of the rank's own part)
			// while a level is small, one launch per level once it is wide (see upward())
			int lf = kTopLevels + 2;
			for (; lf <= L; lf += kSubLevels)
			{
				int ll = std::min(lf + kSubLevels - 1, L);
				while (ll >= lf && ll >= g && (1 << (ll - g)) >= kWideLevel) --ll; // leave the wide levels to the loop below
				if (ll < lf) break;
				const int lroot = lf - 1;
				const int first = lroot >= g ? r << (lroot - g) : r >> (g - lroot), count = lroot >= g ? 1 << (lroot - g) : 1;
				l2l_sub_kernel<P><<<count, 128, 0, st>>>(t, lf, ll, first); ++ctx->launches;
				if (ll < lf + kSubLevels - 1) { lf = ll + 1; break; }
			}
			for (int l = lf; l <= L; ++l)
			{
				const int first = l >= g ? r << (l - g) : r >> (g - l), count = l >= g ? 1 << (l - g) : 1;
				if constexpr (pad4<trl_off(P + 1)>() == 16)
					l2l_level4_kernel<P><<<(4 * count + 127) / 128, 128, 0, st>>>(t, l, first, count);
				else
					l2l_level_kernel<P><<<(count + 127) / 128, 128, 0, st>>>(t, l, first, count);
				++ctx->launches;
			}
		}
		L = Lfull;
		if (ev_l2p) cudaEventRecord(ev_l2p, st);
		const int64_t j_lo = seg_start(n, r, g), j_hi = seg_start(n, r + 1, g);
		const bool uniform = (n & ((1ll << L) - 1)) == 0;
		const int64_t C = n >> L;
		if (fused)
		{
			if constexpr (pad4<trl_off(P + 1)>() == 16)
			{
				const int leaf_lo = (int)(j_lo / C), leaf_hi = (int)(j_hi / C);
				const int gridf = grid_for(4ll * (leaf_hi - leaf_lo), 128, ctx->sm_count, 16);
#define NBCO_L2LP_UNIFORM(CC) l2lp_uniform_kernel<P, CC><<<gridf, 128, 0, st>>>(t, spos, acc_near, acc_out, perm_or_null, param, fuse_elastic, L, leaf_lo, leaf_hi, eps2, coll)
				NBCO_L2LP_UNIFORM(8);
#undef NBCO_L2LP_UNIFORM
			}
			++ctx->launches;
			return;
		}
		const int grid = grid_for(j_hi - j_lo, 128, ctx->sm_count, 16);
#define NBCO_L2P_UNIFORM(CC) l2p_uniform_kernel<P, CC><<<grid, 128, 0, st>>>(t, spos, acc_near, acc_out, perm_or_null, param, fuse_elastic, L, j_lo, j_hi, eps2, coll)
		if (uniform && C == 8) NBCO_L2P_UNIFORM(8);
		else if (uniform && C == 16) NBCO_L2P_UNIFORM(16);
		else if (uniform && C == 32) NBCO_L2P_UNIFORM(32);
		else
			l2p_kernel<P><<<grid, 128, 0, st>>>(t, spos, acc_near, acc_out, perm_or_null, param, fuse_elastic, n, L, j_lo, j_hi, eps2, coll);
#undef NBCO_L2P_UNIFORM
		++ctx->launches;
	}
};

} // namespace

#define NBCO_INSTANTIATE_ORDER(P)                                                                          \
	extern const OrderOps kOrderOps##P;                                                                    \
	const OrderOps kOrderOps##P = {OrderImpl<P>::upward, OrderImpl<P>::m2l, OrderImpl<P>::downward, 1, OrderImpl<P>::finish_leaf_locals, OrderImpl<P>::sparse_near};

} // namespace nbco
