// fmm3_common.cuh -- types and index arithmetic shared by fmm3.cu (tree build, traversal, near
// field, host driver) and the per-order operator translation units fmm3_p*.cu.
#pragma once
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace nbco {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;

#ifndef NBCO_BOTTOM_CAP
#define NBCO_BOTTOM_CAP 8192
#endif
#ifndef NBCO_BOTTOM_THREADS
#define NBCO_BOTTOM_THREADS 1024
#endif
constexpr int kBottomCap = NBCO_BOTTOM_CAP;   // particles a bottom CTA keeps in shared memory (22 bytes each + block state)
constexpr int kBottomThreads = NBCO_BOTTOM_THREADS;
constexpr int kBottomCtasPerSm = kBottomCap <= 4096 && kBottomThreads <= 512 ? 2 : 1; // two resident CTAs need <= 64 registers and half the memory
constexpr int kNoAxis = 3;
constexpr int kFlagShift = 28;             // interaction-list entries carry two target flags above the node id
constexpr int kNodeMask = (1 << kFlagShift) - 1;

// ---- index arithmetic of the implicit tree (fmm_cart3_kdtree.cuh:33-78,117-118) ----
__host__ __device__ __forceinline__ int kd_beg(int l) { return (1 << l) - 1; }
__host__ __device__ __forceinline__ int64_t seg_start(int64_t n, int64_t i, int l)
{
	return i <= 0 ? 0 : (((n * i - 1) >> l) + 1); // ceil(n*i / 2^l)
}
__device__ __forceinline__ int node_level(int node) { return 31 - __clz(node + 1); }

__host__ __device__ __forceinline__ u32 ordered_bits(float f)
{
#ifdef __CUDA_ARCH__
	u32 u = __float_as_uint(f);
#else
	u32 u; memcpy(&u, &f, 4);
#endif
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_bits(u32 k)
{
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__device__ __forceinline__ int widest_axis(float dx, float dy, float dz)
{
	return (dx > dy) ? ((dx > dz) ? 0 : 2) : ((dy > dz) ? 1 : 2); // :92,129 (ties go to the later axis)
}

// kd_size (:395-399) with the host's operation order: (dx*dx + dy*dy) + dz*dz, no contraction
__device__ __forceinline__ float box_size2(const float *lb, const float *rb)
{
	float dx = __fsub_rn(rb[0], lb[0]), dy = __fsub_rn(rb[1], lb[1]), dz = __fsub_rn(rb[2], lb[2]);
	return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int chain_push(int axis, int parent_chain)
// axes in order of most recent use, without repeats; 2 bits each, 3 = none
{
	int c = axis, k = 1;
	for (int s = 0; s < 3; ++s)
	{
		int a = (parent_chain >> (2 * s)) & 3;
		if (a != kNoAxis && a != axis && k < 3) { c |= a << (2 * k); ++k; }
	}
	for (; k < 3; ++k) c |= kNoAxis << (2 * k);
	return c;
}

struct TreeGeom
{
	float *lbound, *rbound;  // float3 per node
	float *size2;            // kd_size per node
	int *splitdim;           // widest axis per node
	int *chain;              // tie-break chain per node
	// Axis and chain the BUILD splits a node by.  Ordinary trees: the same arrays as splitdim / chain.  Shallow trees (a
	// level-(L-1) node holds more particles than a bottom CTA): the build runs deeper than the tree, and the nodes from the
	// leaf level on ("virtual" below it) split by the axis and chain of their parent, so that the concatenation of the
	// virtual leaves is the leaf sorted along its parent's axis -- what the reference's per-level sort leaves behind.
	int *baxis, *bchain;
	int first_leaf;          // kd_beg(L) of the TREE: nodes from here on inherit when baxis != splitdim
};

// inherit: this node splits like its parent (see TreeGeom::baxis)
__device__ __forceinline__ bool kd_inherits(const TreeGeom &g, int node) { return g.baxis != g.splitdim && node >= g.first_leaf && node > 0; }

__device__ __forceinline__ void write_box(const TreeGeom &g, int node, const float *lb, const float *rb, int parent_chain, int parent_axis)
{
	g.lbound[3*node] = lb[0]; g.lbound[3*node+1] = lb[1]; g.lbound[3*node+2] = lb[2];
	g.rbound[3*node] = rb[0]; g.rbound[3*node+1] = rb[1]; g.rbound[3*node+2] = rb[2];
	int ax = widest_axis(rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2]);
	g.splitdim[node] = ax;
	g.chain[node] = chain_push(ax, parent_chain);
	g.size2[node] = box_size2(lb, rb);
	if (g.baxis != g.splitdim)
	{
		const bool inh = kd_inherits(g, node);
		g.baxis[node] = inh ? parent_axis : ax;
		g.bchain[node] = inh ? parent_chain : chain_push(ax, parent_chain);
	}
}


// Multi-GPU over peer memory (peer.cu): rank r of 2^g owns the subtree of kd node (g, r).  Every rank keeps
// full-size node arrays but fills only its own subtree (levels >= g) and the replicated top (levels < g);
// data of a remote node or leaf is read from its owner's arrays, mapped through CUDA IPC over NVLink.
// Single GPU / replicated mode: g = 0 and slot 0 holds the local arrays.
constexpr int kMaxPeers = 8;
struct PeerTab
{
	const float4 *center[kMaxPeers];
	const float *mpole[kMaxPeers];
	const float *pos[kMaxPeers];   // tree-ordered positions: every owner publishes its own range
	int g, me;
};

__device__ __forceinline__ int node_owner(const PeerTab &p, int node)
{
	if (p.g == 0) return 0;
	const int l = node_level(node);
	return l < p.g ? p.me : ((node - kd_beg(l)) >> (l - p.g));
}

struct TreeData
{
	float4 *center;   // xyz = centre of charge, w = kd_size of the node's box (read by the MAC)
	const float *size2;
	float *mpole;     // sM floats per node, symmetric tuple orders 0..P-1
	float *local;     // sL floats per node, traceless tuple orders 0..P
	int sM, sL;
	PeerTab peers;
	// sparse near field (pair-list flow, OrderOps::sparse_near): bit (leaf) is set when the pair kernel added something to
	// the rows of that leaf in acc_near; the L2P kernel reads (and re-zeroes) only those rows.  nullptr: acc_near is dense
	const u32 *nearbits = nullptr;
};

// centre / multipole tuple of a node that may live on another GPU
__device__ __forceinline__ float4 node_center(const TreeData &t, int node) { return t.peers.center[node_owner(t.peers, node)][node]; }
__device__ __forceinline__ const float *node_mpole(const TreeData &t, int node)
{
	return t.peers.mpole[node_owner(t.peers, node)] + (int64_t)node * t.sM;
}

// kd-tree geometry and build scratch (kdtree.cu)
struct KdTree
{
	int64_t n = 0;
	int L = 0, lt = 0;
	int Lb = 0;          // depth the BUILD runs to: L, or deeper for shallow trees (virtual levels, TreeGeom::baxis)
	DevBuf lbound, rbound, size2, splitdim, chain;   // per node (of the build: 2^(Lb+1) - 1; the tree's nodes are a prefix)
	DevBuf baxis, bchain; // only when Lb > L
	DevBuf payA, payB;   // (x, y, z, id) records, ping-pong between the top levels (16 B / particle each)
	DevBuf hist, seg;    // per-segment linear histograms and selection state of the current top level
	DevBuf spos, perm, bbox;
	bool bottom_attr = false;
};
int kd_reserve(nbco_ctx *ctx, KdTree &t, int64_t n, int L);
// r, g: build the top g levels over all particles and, below them, only the subtree of node (g, r)
// (g = 0: the whole tree)
int kd_build(nbco_ctx *ctx, KdTree &t, const float *pos, cudaEvent_t ev_bottom, int r, int g);
// distributed build over the ranks' published buffers (peer mode): no rank holds or partitions all particles
int kd_build_peer(nbco_ctx *ctx, KdTree &t, const float *pos, cudaEvent_t ev_bottom, int r, int g, void *const *pub);
int kd_preload_kernels();   // force the (lazy) module loads of the build kernels
void kd_release(KdTree &t);

// leaf (or node of level l) that owns sorted position j: floor(2^l j / n) (:162-164) without a 64-bit
// division: multiply by magic = floor(2^64 / n) and correct by at most one step
__device__ __forceinline__ int owner_of(int64_t j, int64_t n, int l, unsigned long long magic)
{
	int q = (int)__umul64hi(((unsigned long long)j) << l, magic);
	if (seg_start(n, q + 1, l) <= j) ++q;
	return q;
}

// Interaction lists bucketed BY TARGET (round 2): one compressed-row structure over the target ids
//   [0, ntot)            M2L: the sources of every node        (entries = source node ids)
//   [ntot, ntot + 2^L)   P2P: the source leaves of every leaf  (entries = source leaf node ids)
// Rows are sorted by source id, so every local expansion and every acceleration is ONE deterministic sum in registers,
// stored once: no float atomics, no zero-fill of the locals, no near-field accumulation buffer.
struct CsrView
{
	const u32 *off;   // ntot + 2^L + 1 row offsets
	const int *src;
	int ntot;
	u32 cap;          // entries of src (rows are clamped to it: an evaluation that does not fit is repeated by the host)
};

// near field of one particle against the particles of one source leaf (fmm_p2p_interaction, :767-792, one direction)
__device__ __forceinline__ void leaf_p2p(float *f, const float *__restrict__ lp, int cnt, float x, float y, float z, float eps2)
{
	float ax = 0.f, ay = 0.f, az = 0.f;
	auto term = [&](int k)
	{
		const float dx = x - lp[3*k], dy = y - lp[3*k+1], dz = z - lp[3*k+2];
		float r2 = fmaf(dx, dx, eps2);
		r2 = fmaf(dy, dy, r2);
		r2 = fmaf(dz, dz, r2);
		float w;
		asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(r2));
		w = w * fmaf(-0.5f * r2 * w, w, 1.5f);
		const float w3 = (w * w) * w;
		ax = fmaf(dx, w3, ax); ay = fmaf(dy, w3, ay); az = fmaf(dz, w3, az);
	};
#pragma unroll
	for (int k = 0; k < 8; ++k)
		if (k < cnt) term(k);
	for (int k = 8; k < cnt; ++k) term(k);
	f[0] += ax; f[1] += ay; f[2] += az;
}

// intra-leaf near field of one particle (fmm_p2p3_self_kdtree, :1048-1120): the other particles of its leaf, read
// through L1 (a leaf is 1-2 cache lines); i = j contributes exactly 0
__device__ __forceinline__ void self_p2p(float *f, const float *__restrict__ spos, int64_t j, int leaf,
                                         float x, float y, float z, int64_t n, int L, float eps2)
{
	const int64_t s0 = seg_start(n, leaf, L);
	const int cnt = (int)(seg_start(n, leaf + 1, L) - s0);
	(void)j;
	leaf_p2p(f, spos + 3 * s0, cnt, x, y, z, eps2);
}

// Order-specific passes (one translation unit per order: the unrolled templates are expensive
// to compile, the reference's single TU takes > 4 min).
struct OrderOps
{
	// part 0: P2M and M2M of rank r's subtree (levels L .. g); part 1: the replicated top (levels g-1 .. 0),
	// whose level-g children are read from their owners.  One rank: g = 0, part 0 is the whole pass.
	void (*upward)(nbco_ctx *ctx, TreeData t, const float *spos, int64_t n, int L, int r, int g, int part);
	void (*m2l)(nbco_ctx *ctx, TreeData t, const int2 *list, const unsigned *count, unsigned cap, float eps2);
	// rank r of 2^g ranks pushes locals down its own subtree (plus the ancestors of its root) only
	// csr == nullptr: pair lists (M2L done by m2l(), near field summed into acc_near by the pair kernel); csr != nullptr:
	// by-target flow: the levels gather their M2L sources themselves (no m2l() call, locals need no zero-fill) and the
	// L2P kernel gathers the near field (acc_near unused)
	void (*downward)(nbco_ctx *ctx, TreeData t, const float *spos, float *acc_near, float *acc_out,
	                 const int *perm_or_null, const float *param, int fuse_elastic, int64_t n, int L, int r, int g,
	                 float eps2, int coll, cudaEvent_t ev_l2p /* recorded between the L2L levels and the L2P kernel */,
	                 const CsrView *csr);
	int by_target;   // 1: the passes of this order implement the by-target flow
	// pair-list flow: downward() may leave the leaf level of the L2L pass to its L2P kernel (locals of the leaves then hold
	// the M2L sums only); this pushes that level for callers that read the tuples back.  nullptr: never fused
	void (*finish_leaf_locals)(nbco_ctx *ctx, TreeData t, int64_t n, int L, int r, int g);
	// 1: for this (n, L) downward() reads acc_near only where t.nearbits says so and leaves acc_near all zero again, so the
	// caller clears the bit map (2^L bits) instead of acc_near (12 n bytes) before the pair kernel.  nullptr: never
	int (*sparse_near)(int64_t n, int L);
};
const OrderOps *order_ops(int order);

} // namespace nbco
