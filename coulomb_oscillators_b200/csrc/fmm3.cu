// fmm3.cu -- 3D Cartesian FMM on a balanced kd-tree for sm_100a.
//
// Replaces fmm_cart3_kdtree (reference Simulation/fmm_cart3_kdtree.cuh:1478-1771) and everything it
// launches: CUB radix sort + bb_segsort + gather/copy passes per level (:1311-1364), evalBox/evalKeys
// (:109-202), centerLeaves/multLeaves (appel.cuh:184-258), P2M/M2M (:231-393), the 18-block dual
// traversal (:416-567), P2P (:767-1132), M2L (:613-765), L2L/L2P (:1134-1309), rescale
// (appel.cuh:506-527) and, for the coulombOscillator evaluator, add_elastic (kernel.cuh:119-152).
//
// Design (DESIGN.md has the long form):
//  * the tree is the reference's: node i of level l owns sorted range [ceil(n i/2^l), ceil(n (i+1)/2^l)),
//    split axis = widest box extent, child boxes cut at the boundary particles' coordinates.  index/mult
//    are pure functions of (n, l, i) and are never stored.
//  * build (kdtree.cu): every level is a median SELECTION + unordered partition, not a sort.  Levels whose
//    segments exceed kBottomCap particles: one histogram pass + one three-way partition pass over 16-byte
//    (x, y, z, id) records; all remaining levels run in ONE kernel, a CTA per subtree holding its particles in
//    shared memory.  Equal keys are ordered as a stable sort at every level would order them (ties fall back to
//    the coordinates of the previous split axes, then the input index).  One gather at the end.
//  * geometry that decides the interaction lists (centres, box sizes, MAC) is evaluated with the
//    reference's host operation order and no FMA contraction (__fmul_rn/__fadd_rn), the MAC's
//    pow() is a host-computed table (only two multiplicities exist per level), so tree arrays and
//    lists are bit-exact against the reference's CPU path.
//  * traversal is level-synchronous over all SMs (the reference uses 18 blocks), frontier in global
//    memory, warp-aggregated appends.
//  * operators are compile-time unrolled templates (fmm_ops.cuh), one instantiation per order.

#include "fmm3_common.cuh"
#include <cooperative_groups.h>

namespace nbco {

namespace {

// =====================================================================================
//  permutation glue (replaces the gather_krnl/copy_krnl pairs, kernel.cuh:228-311)
// =====================================================================================
__global__ void __launch_bounds__(256) gather3_kernel(const float *__restrict__ src, const int *__restrict__ perm, float *__restrict__ dst, int64_t n)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
	{
		int64_t s = perm[j];
		dst[3*j] = src[3*s]; dst[3*j+1] = src[3*s+1]; dst[3*j+2] = src[3*s+2];
	}
}

// peer mode: zero the locals of rank r's own subtree only (levels g .. L; the replicated top is cleared by a small
// memset).  A full-size memset costs 64 B x all nodes on every rank and does not shrink with the number of GPUs.
__global__ void __launch_bounds__(256) zero_own_locals_kernel(float4 *__restrict__ local4, int q4 /* float4 per node */, int L, int r, int g)
{
	const int64_t own_nodes = ((int64_t)1 << (L - g + 1)) - 1; // heap order inside the own subtree
	const int64_t total = own_nodes * q4, stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride)
	{
		const int64_t m = k / q4;
		const int lo = 63 - __clzll(m + 1);
		const int64_t node = kd_beg(g + lo) + ((int64_t)r << lo) + (m + 1 - ((int64_t)1 << lo));
		local4[node * q4 + (k - m * q4)] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
}

__global__ void __launch_bounds__(256) gather1_kernel(const int *__restrict__ src, const int *__restrict__ perm, int *__restrict__ dst, int64_t n)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) dst[j] = src[perm[j]];
}

// peer mode: the velocity of sorted particle j comes from whoever held its pre-rebuild index (tree-order ranges
// are fixed index ranges, so that is rank floor(2^g idx / n)); most particles stay on their rank between rebuilds
struct VelSrc { const float *v[kMaxPeers]; int g, me; u32 *dbg; };
__global__ void __launch_bounds__(256) gather3_peer_kernel(VelSrc src, const int *__restrict__ perm, float *__restrict__ dst, int64_t cnt, int64_t n)
{
	const unsigned long long magic = ~0ull / (unsigned long long)n;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += stride)
	{
		int64_t s = perm[j];
		if (s < 0 || s >= n) { if (src.dbg) { src.dbg[40] = 300u; src.dbg[41] = (u32)j; src.dbg[42] = (u32)s; } s = 0; }
		const float *v = src.v[owner_of(s, n, src.g, magic)];
		dst[3*j] = v[3*s]; dst[3*j+1] = v[3*s+1]; dst[3*j+2] = v[3*s+2];
	}
}

// =====================================================================================
//  dual tree traversal, level-synchronous (replaces fmm_dualTraversal, :429-567)
// =====================================================================================
struct TravArgs
{
	const float4 *center;
	const float *mfac;    // [level][2]: MAC factor M for the low / high multiplicity of the level
	int2 *p2p, *m2l, *front_in, *front_out;
	u32 *cnt;             // [0] p2p, [1] m2l, [2..4] frontier sizes (rotating)
	u32 cap_p2p, cap_m2l, cap_front;
	int64_t n;
	int ntot, L, m2l_first;
	float radius;
	int64_t sh_lo, sh_hi; // particles (tree order) whose accelerations this rank computes
	PeerTab peers;        // centres of remote nodes are read from their owners
};

__device__ __forceinline__ float4 trav_center(const TravArgs &a, int node) { return a.peers.center[node_owner(a.peers, node)][node]; }

__device__ __forceinline__ int node_mult(int64_t n, int node, int &level)
{
	level = node_level(node);
	int i = node - kd_beg(level);
	return (int)(seg_start(n, i + 1, level) - seg_start(n, i, level));
}

// multi-GPU: a node is a target of this rank iff its particle range meets the rank's shard
__device__ __forceinline__ bool node_mine(const TravArgs &a, int node)
{
	const int level = node_level(node);
	const int i = node - kd_beg(level);
	return seg_start(a.n, i, level) < a.sh_hi && seg_start(a.n, i + 1, level) > a.sh_lo;
}

// kd_admissible (:401-414) with the host's operation order; pow() comes from the host table
__device__ __forceinline__ bool mac_ok(const TravArgs &a, int n1, int n2)
{
	float4 c1 = trav_center(a, n1), c2 = trav_center(a, n2);
	float dx = __fsub_rn(c2.x, c1.x), dy = __fsub_rn(c2.y, c1.y), dz = __fsub_rn(c2.z, c1.z);
	float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	float sz = fmaxf(c1.w, c2.w);
	int l1, l2;
	int m1 = node_mult(a.n, n1, l1), m2 = node_mult(a.n, n2, l2);
	int lv = m1 >= m2 ? l1 : l2, mm = m1 >= m2 ? m1 : m2;
	float M = a.mfac[2 * lv + (mm != (int)(a.n >> lv))];
	float parM = __fmul_rn(a.radius, M);
	return __fmul_rn(__fmul_rn(parM, parM), sz) < dist2;
}

// append `count` items per lane with one atomic per warp; returns this lane's first slot
__device__ __forceinline__ u32 warp_append(u32 *counter, int count)
{
	const int lane = threadIdx.x & 31;
	int incl = count;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		int v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += v;
	}
	int total = __shfl_sync(0xffffffffu, incl, 31);
	u32 base = 0;
	if (lane == 31 && total > 0) base = atomicAdd(counter, (u32)total);
	base = __shfl_sync(0xffffffffu, base, 31);
	return base + (u32)(incl - count);
}

// one pair of the dual traversal (fmm_dualTraversal_cpu, :581-610; m2l_first as in the GPU kernel :504-542)
// returns 0 nothing, 1 p2p, 2 m2l, 3 self split, 4 split y, 5 split x; flags = which side is a target here
__device__ __forceinline__ int classify_pair(const TravArgs &a, int2 np, int &flags)
{
	// bit 0: np.x is a target of this rank, bit 1: np.y is; untouched pairs (and their descendants) are
	// dropped before any node data is read.  With one rank every node is a target (flags = 3).
	flags = 3;
	if (a.sh_lo > 0 || a.sh_hi < a.n)
	{
		flags = (node_mine(a, np.x) ? 1 : 0) | (node_mine(a, np.y) ? 2 : 0);
		if (!flags) return 0;
	}
	int kind;
	const bool xl = 2*np.x + 1 >= a.ntot, yl = 2*np.y + 1 >= a.ntot;
	if (!a.m2l_first && xl && yl) kind = (np.x != np.y) ? 1 : 0;
	else if (np.x == np.y && !xl) kind = 3;
	else if (np.x != np.y && mac_ok(a, np.x, np.y)) kind = 2;
	else if (xl && yl) kind = (np.x != np.y) ? 1 : 0;
	else if (xl || (!yl && trav_center(a, np.x).w <= trav_center(a, np.y).w)) kind = 4;
	else kind = 5;
	if (!kind) flags = 0;
	return kind;
}

__device__ __forceinline__ int expand_pair(int kind, int2 np, int2 *out)
{
	if (kind == 3)
	{
		out[0] = make_int2(2*np.x + 1, 2*np.x + 1);
		out[1] = make_int2(2*np.x + 1, 2*np.x + 2);
		out[2] = make_int2(2*np.x + 2, 2*np.x + 2);
		return 3;
	}
	if (kind == 4) { out[0] = make_int2(np.x, 2*np.y + 1); out[1] = make_int2(np.x, 2*np.y + 2); return 2; }
	if (kind == 5) { out[0] = make_int2(2*np.x + 1, np.y); out[1] = make_int2(2*np.x + 2, np.y); return 2; }
	return 0;
}

// level-synchronous round.  Appends are aggregated per CTA: the three output counters (p2p list, m2l
// list, next frontier) receive ONE atomic each per 256 classified pairs -- per-warp atomics on the same
// three addresses serialise in L2 and dominated the big rounds (profiles/r01_notes.md).
__device__ __forceinline__ void traverse_round(const TravArgs &a, const int2 *__restrict__ front_in, int2 *__restrict__ front_out, int round,
                                               u32 (*wtot)[3], u32 *base)
{
	u32 *cin = a.cnt + 2 + round % 3, *cout = a.cnt + 2 + (round + 1) % 3, *cnext = a.cnt + 2 + (round + 2) % 3;
	if (blockIdx.x == 0 && threadIdx.x == 0) *cnext = 0; // nobody touches it during this round
	const u32 nin = min(*(volatile u32 *)cin, a.cap_front);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (u32 w0 = blockIdx.x * blockDim.x; w0 < nin; w0 += gridDim.x * blockDim.x)
	{
		const u32 w = w0 + threadIdx.x;
		int kind = 0, flags = 0;
		int2 np = make_int2(0, 0);
		if (w < nin) { np = front_in[w]; kind = classify_pair(a, np, flags); }
		int2 kids[3];
		const int nf = expand_pair(kind, np, kids);
		// packed per-lane counts: bits 0..9 p2p, 10..19 m2l, 20..31 frontier entries
		const u32 mine = (kind == 1 ? 1u : 0u) | (kind == 2 ? 1u << 10 : 0u) | ((u32)nf << 20);
		u32 incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u32 v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += v;
		}
		if (lane == 31) { wtot[warp][0] = incl & 1023u; wtot[warp][1] = (incl >> 10) & 1023u; wtot[warp][2] = incl >> 20; }
		__syncthreads();
		if (threadIdx.x < 3)
		{
			u32 tot = 0;
			for (int i = 0; i < 8; ++i) tot += wtot[i][threadIdx.x];
			u32 *c = threadIdx.x == 0 ? a.cnt + 0 : (threadIdx.x == 1 ? a.cnt + 1 : cout);
			base[threadIdx.x] = tot ? atomicAdd(c, tot) : 0u;
		}
		__syncthreads();
		const u32 excl = incl - mine;
		u32 s1 = base[0] + (excl & 1023u), s2 = base[1] + ((excl >> 10) & 1023u), s3 = base[2] + (excl >> 20);
		for (int i = 0; i < warp; ++i) { s1 += wtot[i][0]; s2 += wtot[i][1]; s3 += wtot[i][2]; }
		const int2 tagged = make_int2(np.x | (flags << kFlagShift), np.y);
		if (kind == 1 && s1 < a.cap_p2p) a.p2p[s1] = tagged;
		if (kind == 2 && s2 < a.cap_m2l) a.m2l[s2] = tagged;
		if (nf && s3 + nf > a.cap_front) a.cnt[5] = 1u; // sticky: a frontier did not fit
		if (nf && s3 + nf <= a.cap_front)
			for (int k = 0; k < nf; ++k) front_out[s3 + k] = kids[k];
		__syncthreads(); // wtot / base are reused by the next chunk
	}
}

__global__ void __launch_bounds__(256) traverse_round_kernel(TravArgs a, int round)
{
	__shared__ u32 wtot[8][3];
	__shared__ u32 base[3];
	traverse_round(a, a.front_in, a.front_out, round, wtot, base);
}

// all rounds in one cooperative launch: a grid-wide barrier replaces ~2L kernel boundaries (most rounds
// classify a few thousand pairs and are dominated by launch and ramp-up time)
__global__ void __launch_bounds__(256) traverse_all_kernel(const TravArgs a, int2 *fa, int2 *fb, int max_rounds)
{
	__shared__ u32 wtot[8][3];
	__shared__ u32 base[3];
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	for (int r = 0; r < max_rounds; ++r)
	{
		// (the kernel parameters are never written: a modified copy would live in local memory, and the
		// owner-indexed centre pointers are read with a dynamic index)
		if (*(volatile u32 *)(a.cnt + 2 + r % 3) == 0) break; // uniform: the counter was final before the last barrier
		traverse_round(a, (r & 1) ? fb : fa, (r & 1) ? fa : fb, r, wtot, base);
		grid.sync();
	}
}

// -------------------------------------------------------------------------------------------------
// Asynchronous traversal: ONE work queue instead of level-synchronous rounds.  The breadth-first version
// pays the latency of a round (frontier read, two dependent centre loads, three counter atomics, a grid
// barrier: ~6-10 us) about 2L times although most rounds hold a handful of pairs (measured: running the
// small rounds on one CTA without grid barriers did not help, the dependent loads dominate).  Here a pair is
// processed as soon as its parent has been: queue slot k holds SENT until a producer stores the pair with
// one 64-bit write; consumers reserve slots in order (one atomic per CTA pass), poll only slots below the
// reserved tail, process the ready ones, append their children at the tail and restore SENT behind them.
// Termination: processed == tail, read in that order.
// Launched cooperatively only to guarantee that every CTA is resident (producers never starve).
// -------------------------------------------------------------------------------------------------
constexpr unsigned long long kSent = ~0ull;
// kCarry: a splitting thread keeps its last child instead of queueing it (depth-first chains).  Measured slower on
// B200 (0.12 vs 0.10 ms at 64 k particles) and it scatters the M2L list (M2L 0.27 -> 0.46 ms at 16 M): off.
constexpr bool kCarry = false;

__global__ void __launch_bounds__(256) traverse_queue_kernel(TravArgs a, unsigned long long *q)
{
	// cnt[0] p2p, cnt[1] m2l, cnt[5] overflow, cnt[8] head (slots handed out), cnt[9] tail (slots reserved by
	// producers), cnt[10] finished chains.  A thread that splits a pair keeps the LAST child in registers and
	// queues the others, so the dependent chain root -> leaf costs one classification per level instead of a
	// queue round trip; a chain starts with a pop and counts as finished when it ends without children.
	__shared__ u32 wtot[8][3], wwant[8], wready[8], wend[8];
	__shared__ u32 base[4];
	__shared__ u32 s_tail, s_done, s_any, s_end;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 slot = 0xffffffffu; // this thread's queue slot (none yet)
	bool carry = false;
	int2 cur = make_int2(0, 0);
	for (;;)
	{
		// (1) hand a fresh slot to every idle thread that has none: one atomic per CTA
		const bool want = !carry && slot == 0xffffffffu;
		const u32 wmask = __ballot_sync(0xffffffffu, want);
		if (lane == 0) wwant[warp] = __popc(wmask);
		__syncthreads();
		if (threadIdx.x == 0)
		{
			u32 tot = 0;
			for (int i = 0; i < 8; ++i) tot += wwant[i];
			base[3] = tot ? atomicAdd(a.cnt + 8, tot) : 0u;
			// finished is read BEFORE tail: finished(t1) == tail(t2 > t1) implies no chain was alive at t1
			const u32 done = *(volatile u32 *)(a.cnt + 10);
			__threadfence();
			const u32 tail = *(volatile u32 *)(a.cnt + 9);
			s_tail = tail; s_done = (done == tail) ? 1u : 0u;
		}
		__syncthreads();
		if (want)
		{
			u32 off = base[3] + __popc(wmask & ((1u << lane) - 1u));
			for (int i = 0; i < warp; ++i) off += wwant[i];
			slot = off;
		}
		const u32 tail = s_tail;
		if (s_done) break; // uniform: every queued pair was popped and every chain has ended
		// (2) the pair of this pass: the carried child, or a queued pair (only slots below the reserved tail
		// can hold, or be about to hold, one)
		bool ready = carry;
		int2 np = cur;
		if (!carry && slot < tail && slot < a.cap_front)
		{
			const unsigned long long v = *(volatile unsigned long long *)(q + slot);
			if (v != kSent)
			{
				*(volatile unsigned long long *)(q + slot) = kSent; // the queue is clean again for the next evaluation
				np = make_int2((int)(u32)v, (int)(u32)(v >> 32));
				ready = true;
				slot = 0xffffffffu;
			}
		}
		int kind = 0, flags = 0;
		if (ready) kind = classify_pair(a, np, flags);
		int2 kids[3];
		const int nf = expand_pair(kind, np, kids);
		const int npush = (kCarry && nf > 0) ? nf - 1 : nf;
		// (3) CTA-aggregated appends: p2p list, m2l list, queue tail
		const u32 mine = (kind == 1 ? 1u : 0u) | (kind == 2 ? 1u << 10 : 0u) | ((u32)npush << 20);
		u32 incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u32 x = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += x;
		}
		const bool ended = ready && (!kCarry || nf == 0);
		const u32 rmask = __ballot_sync(0xffffffffu, ready), emask = __ballot_sync(0xffffffffu, ended);
		if (lane == 31) { wtot[warp][0] = incl & 1023u; wtot[warp][1] = (incl >> 10) & 1023u; wtot[warp][2] = incl >> 20; }
		if (lane == 0) { wready[warp] = __popc(rmask); wend[warp] = __popc(emask); }
		__syncthreads();
		if (threadIdx.x < 3)
		{
			u32 tot = 0;
			for (int i = 0; i < 8; ++i) tot += wtot[i][threadIdx.x];
			u32 *c = threadIdx.x == 0 ? a.cnt + 0 : (threadIdx.x == 1 ? a.cnt + 1 : a.cnt + 9);
			base[threadIdx.x] = tot ? atomicAdd(c, tot) : 0u;
		}
		if (threadIdx.x == 32)
		{
			u32 tr = 0, te = 0;
			for (int i = 0; i < 8; ++i) { tr += wready[i]; te += wend[i]; }
			s_any = tr; s_end = te;
		}
		__syncthreads();
		const u32 nready = s_any, nend = s_end;
		if (nready)
		{
			const u32 excl = incl - mine;
			u32 s1 = base[0] + (excl & 1023u), s2 = base[1] + ((excl >> 10) & 1023u), s3 = base[2] + (excl >> 20);
			for (int i = 0; i < warp; ++i) { s1 += wtot[i][0]; s2 += wtot[i][1]; s3 += wtot[i][2]; }
			const int2 tagged = make_int2(np.x | (flags << kFlagShift), np.y);
			if (kind == 1 && s1 < a.cap_p2p) a.p2p[s1] = tagged;
			if (kind == 2 && s2 < a.cap_m2l) a.m2l[s2] = tagged;
			if (npush)
			{
				if (s3 + npush > a.cap_front) a.cnt[5] = 1u; // sticky: the queue did not fit (the host grows it and repeats)
				for (int k = 0; k < npush; ++k)
					if (s3 + k < a.cap_front)
						*(volatile unsigned long long *)(q + s3 + k) = (unsigned long long)(u32)kids[k].x | ((unsigned long long)(u32)kids[k].y << 32);
					else
						atomicAdd(a.cnt + 10, 1u); // a dropped pair counts as finished, or the queue would never drain
			}
			carry = kCarry && nf > 0;
			if (carry) cur = kids[nf - 1];
			if (nend)
			{
				__threadfence();  // queued children are visible before their parent's chain counts as finished
				__syncthreads();
				if (threadIdx.x == 0) atomicAdd(a.cnt + 10, nend);
			}
		}
		else
			__nanosleep(64);
	}
}

__global__ void traverse_queue_init_kernel(unsigned long long *q, u32 *cnt)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		q[0] = 0ull; // the pair (root, root)
		cnt[0] = 0; cnt[1] = 0; cnt[5] = 0; cnt[8] = 0; cnt[9] = 1; cnt[10] = 0;
	}
}


// -------------------------------------------------------------------------------------------------
// Incremental traversal (evaluations between two tree rebuilds).  The kd partition and the node boxes are
// fixed between rebuilds; only the centres of charge move, so the traversal of the previous evaluation is
// almost the traversal of this one: a pair changes its fate only when its MAC test flips (measured: a few
// thousand of ~6 M visited pairs per step at N = 2^24).  The full traversal RECORDS every visited pair in
// one array V (frontier after frontier; the children of a split pair are contiguous) with a record
// R = kind | flags << 3 | first child << 5.  A reuse evaluation then
//   (1) re-classifies every recorded pair in ONE data-parallel pass and lists the pairs whose kind changed,
//   (2) retires the recorded descendants of the changed pairs (tombstones; a small breadth-first walk over
//       the child links on one CTA) and marks the surviving changed pairs as seeds (they keep their record slot: parents still link to them),
//   (3) emits the interaction lists from the surviving records (data-parallel),
//   (4) runs the ordinary level-synchronous traversal from the seeds only (a few short rounds).
// The lists are the same sets as a traversal from the root: a recorded pair survives iff none of its ancestors
// changed kind, and every changed pair is re-expanded from scratch.  tests/test_fmm_gpu.py compares the lists
// with a from-the-root traversal (NBCO_TRAVERSE=rounds) and with the oracle after several steps.
// -------------------------------------------------------------------------------------------------
constexpr u32 kTomb = 7u;
__device__ __forceinline__ u32 rec_pack(int kind, int flags, u32 child) { return (u32)kind | ((u32)flags << 3) | (child << 5); }

// one round: frontier V[start, start + nin) (or, for the seeds of an update, V[idx[0 .. nin)]: a re-expanded pair keeps
// its record slot, so that its parent's child link still reaches it) -> children appended at V[out0, ...)
__device__ __forceinline__ void rec_round(const TravArgs &a, int2 *__restrict__ V, u32 *__restrict__ R, u32 start, u32 out0,
                                          const u32 *__restrict__ idx, int round, u32 (*wtot)[3], u32 *base)
{
	u32 *cin = a.cnt + 2 + round % 3, *cout = a.cnt + 2 + (round + 1) % 3, *cnext = a.cnt + 2 + (round + 2) % 3;
	if (blockIdx.x == 0 && threadIdx.x == 0) *cnext = 0;
	const u32 nin = *(volatile u32 *)cin;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (u32 w0 = blockIdx.x * blockDim.x; w0 < nin; w0 += gridDim.x * blockDim.x)
	{
		const u32 w = w0 + threadIdx.x;
		int kind = 0, flags = 0;
		int2 np = make_int2(0, 0);
		const u32 me = w < nin ? (idx ? idx[w] : start + w) : 0u;
		if (w < nin) { np = V[me]; kind = classify_pair(a, np, flags); }
		int2 kids[3];
		const int nf = expand_pair(kind, np, kids);
		const u32 mine = (kind == 1 ? 1u : 0u) | (kind == 2 ? 1u << 10 : 0u) | ((u32)nf << 20);
		u32 incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u32 v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += v;
		}
		if (lane == 31) { wtot[warp][0] = incl & 1023u; wtot[warp][1] = (incl >> 10) & 1023u; wtot[warp][2] = incl >> 20; }
		__syncthreads();
		if (threadIdx.x < 3)
		{
			u32 tot = 0;
			for (int i = 0; i < 8; ++i) tot += wtot[i][threadIdx.x];
			u32 *c = threadIdx.x == 0 ? a.cnt + 0 : (threadIdx.x == 1 ? a.cnt + 1 : cout);
			base[threadIdx.x] = tot ? atomicAdd(c, tot) : 0u;
		}
		__syncthreads();
		const u32 excl = incl - mine;
		u32 s1 = base[0] + (excl & 1023u), s2 = base[1] + ((excl >> 10) & 1023u), s3 = base[2] + (excl >> 20);
		for (int i = 0; i < warp; ++i) { s1 += wtot[i][0]; s2 += wtot[i][1]; s3 += wtot[i][2]; }
		const int2 tagged = make_int2(np.x | (flags << kFlagShift), np.y);
		if (kind == 1 && s1 < a.cap_p2p) a.p2p[s1] = tagged;
		if (kind == 2 && s2 < a.cap_m2l) a.m2l[s2] = tagged;
		const bool fits = (unsigned long long)out0 + s3 + nf <= a.cap_front;
		if (nf && !fits) a.cnt[5] = 1u; // sticky: V did not fit (the host grows it and repeats from the root)
		if (nf && fits)
			for (int k = 0; k < nf; ++k) V[out0 + s3 + k] = kids[k];
		if (w < nin) R[me] = rec_pack(kind, flags, out0 + s3);
		__syncthreads();
	}
}

// rounds from the frontier V[cnt[11], cnt[11] + cnt[2]) until nothing is left; cnt[11] = entries of V afterwards
__global__ void __launch_bounds__(256) traverse_rec_kernel(const TravArgs a, int2 *V, u32 *R, const u32 *seeds_arg, int max_rounds)
{
	__shared__ u32 wtot[8][3];
	__shared__ u32 base[3];
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const u32 *seeds = seeds_arg;
	u32 start = *(volatile u32 *)(a.cnt + 11);
	if (seeds && *(volatile u32 *)(a.cnt + 20)) seeds = nullptr; // reuse evaluation that fell back to the root
	grid.sync(); // everybody has read the start before CTA 0 may store the end
	for (int r = 0; r < max_rounds; ++r)
	{
		const u32 nin = *(volatile u32 *)(a.cnt + 2 + r % 3);
		if (nin == 0) break;
		if ((unsigned long long)start + nin > a.cap_front) { if (blockIdx.x == 0 && threadIdx.x == 0) a.cnt[5] = 1u; break; }
		if (r == 0 && seeds)
		{
			// update: the seeds are scattered records, their children start the appended region
			rec_round(a, V, R, 0u, start, seeds, r, wtot, base);
			grid.sync();
		}
		else
		{
			rec_round(a, V, R, start, start + nin, nullptr, r, wtot, base);
			grid.sync();
			start += nin;
		}
		if (blockIdx.x == 0 && threadIdx.x == 0) a.cnt[15] = (u32)r + 1u; // rounds run (diagnostics)
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { a.cnt[11] = start; }
}

__global__ void traverse_rec_init_kernel(int2 *V, u32 *cnt)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		V[0] = make_int2(0, 0);
		cnt[0] = 0; cnt[1] = 0; cnt[2] = 1; cnt[3] = 0; cnt[4] = 0; cnt[5] = 0; cnt[11] = 0; cnt[12] = 0; cnt[20] = 0;
	}
}

// Start of a reuse evaluation.  The records only grow between two rebuilds (re-expanded subtrees are appended, retired ones stay
// as tombstones) and these evaluations are enqueued without a host round trip, so the decision whether the update still fits is
// taken HERE: while at most `limit` = half of the record capacity is in use, the change list and the retire queue (half the
// capacity each) cannot overflow and the other half is room for this evaluation's appends; beyond that the evaluation traverses
// from the root again (cnt[20] = 1: reval / retire / emit return at once, traverse_rec starts from the pair (root, root)), which
// also compacts the records.  The host guarantees capacity >= 3 x the records of a from-the-root traversal at every rebuild.
__global__ void traverse_reuse_init_kernel(u32 *cnt, int2 *V, u32 limit)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		cnt[0] = 0; cnt[1] = 0; cnt[5] = 0; cnt[12] = 0; cnt[13] = 0; cnt[14] = 0; cnt[15] = 0; cnt[16] = 0; cnt[17] = 0; cnt[18] = 0; cnt[19] = 0;
		const bool root = cnt[11] > limit;
		cnt[20] = root ? 1u : 0u;
		if (root) { V[0] = make_int2(0, 0); cnt[2] = 1; cnt[3] = 0; cnt[4] = 0; cnt[11] = 0; cnt[21] += 1u; /* diagnostics: fallbacks so far */ }
	}
}

// (1) which recorded pairs change kind with the new centres?  (kinds 0 and 3 do not depend on the centres)
__global__ void __launch_bounds__(256) reval_kernel(const TravArgs a, const int2 *__restrict__ V, const u32 *__restrict__ R,
                                                    u32 *__restrict__ clist, u32 cap_c)
{
	if (a.cnt[20]) return; // this evaluation traverses from the root (traverse_reuse_init_kernel)
	const u32 count = a.cnt[11];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
	{
		const u32 k = R[i] & 7u;
		if (k == kTomb || k == 0u || k == 3u) continue;
		int flags;
		const int nk = classify_pair(a, V[i], flags);
		if ((u32)nk != k)
		{
			const u32 pos = atomicAdd(a.cnt + 12, 1u);
			if (pos < cap_c) clist[pos] = i; else a.cnt[5] = 1u;
		}
	}
}

// (2) retire the recorded subtrees of the changed pairs (a breadth-first walk over the child links: measured 4-6
// rounds, ~30 k records per step at N = 2^24), then list the surviving changed pairs as seeds.  Cooperative launch:
// a grid barrier per round.  cnt[14] = walk queue tail, cnt[16] = seeds.
__device__ __forceinline__ u32 warp_reserve(u32 *counter, u32 count)
{
	const int lane = threadIdx.x & 31;
	u32 incl = count;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += v;
	}
	const u32 total = __shfl_sync(0xffffffffu, incl, 31);
	u32 base = 0;
	if (lane == 31 && total) base = atomicAdd(counter, total);
	base = __shfl_sync(0xffffffffu, base, 31);
	return base + incl - count;
}

__global__ void __launch_bounds__(256) retire_kernel(const TravArgs a, u32 *__restrict__ R, const u32 *__restrict__ clist,
                                                     u32 *__restrict__ bq, u32 cap_q)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	if (a.cnt[20]) return; // from-the-root evaluation: nothing to retire (uniform over the grid: no barrier is left waiting)
	const u32 nc = a.cnt[12];
	const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
	const u32 nc_pad = (nc + 31u) & ~31u; // whole warps stay in the loops (warp_reserve shuffles)
	// three rotating frontier counters (cnt[17..19]) like the traversal rounds: the size of round r is final at the
	// barrier, the pushes of round r count into the next one, the one after is cleared -- no CTA ever reads a
	// counter that a faster CTA is already adding to
	for (u32 c = gtid; c < nc_pad; c += gstride)
	{
		u32 rec = 0, nch = 0;
		if (c < nc) { rec = R[clist[c]]; const u32 k = rec & 7u; nch = (k >= 3u && k <= 5u) ? (k == 3u ? 3u : 2u) : 0u; }
		const u32 pos = warp_reserve(a.cnt + 17, nch);
		for (u32 j = 0; j < nch; ++j) if (pos + j < cap_q) bq[pos + j] = (rec >> 5) + j; else a.cnt[5] = 1u;
	}
	grid.sync();
	u32 head = 0;
	for (int r = 0;; ++r)
	{
		u32 *cin = a.cnt + 17 + r % 3, *cout = a.cnt + 17 + (r + 1) % 3, *cnext = a.cnt + 17 + (r + 2) % 3;
		const u32 span = *(volatile u32 *)cin;
		if (span == 0 || head + span > cap_q) break; // uniform
		if (gtid == 0) *cnext = 0;
		const u32 out0 = head + span, span_pad = (span + 31u) & ~31u;
		for (u32 q = gtid; q < span_pad; q += gstride)
		{
			u32 rec = 0, nch = 0;
			if (q < span)
			{
				const u32 j = bq[head + q];
				rec = R[j];
				const u32 k = rec & 7u;
				R[j] = (rec & ~7u) | kTomb;
				nch = (k >= 3u && k <= 5u) ? (k == 3u ? 3u : 2u) : 0u;
			}
			const u32 pos = out0 + warp_reserve(cout, nch);
			for (u32 t = 0; t < nch; ++t) if (pos + t < cap_q) bq[pos + t] = (rec >> 5) + t; else a.cnt[5] = 1u;
		}
		if (gtid == 0) { a.cnt[13] += 1u; a.cnt[14] = out0; }
		grid.sync();
		head = out0;
	}
	// changed pairs that were not retired as somebody's descendant start a fresh expansion (in place: see rec_round)
	for (u32 c = gtid; c < nc_pad; c += gstride)
	{
		u32 i = 0, want = 0;
		if (c < nc)
		{
			i = clist[c];
			const u32 rec = R[i];
			if ((rec & 7u) != kTomb) { R[i] = (rec & ~7u) | kTomb; want = 1u; } // not emitted; the traversal rewrites the record
		}
		const u32 pos = warp_reserve(a.cnt + 16, want);
		if (want) { if (pos < cap_q) bq[pos] = i; else a.cnt[5] = 1u; } // the walk is over: its queue now holds the seed list
	}
	grid.sync();
	if (gtid == 0) { a.cnt[2] = a.cnt[16]; a.cnt[3] = 0; a.cnt[4] = 0; }
}

// (3) the lists of the surviving records.  Four consecutive records per thread (one 16-byte load of R): a tile of 1024 records
// costs one pair of list-counter atomics and two CTA barriers -- with one record per thread the kernel was bound by the
// latency of those atomics (0.067 ms for 6 M records at N = 2^24, 80 MB of traffic).
__global__ void __launch_bounds__(256) emit_kernel(const TravArgs a, const int2 *__restrict__ V, const u32 *__restrict__ R)
{
	__shared__ u32 wtot[8][2];
	__shared__ u32 base[2];
	if (a.cnt[20]) return; // from-the-root evaluation: the traversal rounds write the lists themselves
	const u32 count = a.cnt[11];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (u32 i0 = blockIdx.x * (blockDim.x * 4u); i0 < count; i0 += gridDim.x * (blockDim.x * 4u))
	{
		const u32 i = i0 + 4u * threadIdx.x;
		u32 rec[4] = {0u, 0u, 0u, 0u};
		if (i + 3u < count) { const uint4 r4 = *reinterpret_cast<const uint4 *>(R + i); rec[0] = r4.x; rec[1] = r4.y; rec[2] = r4.z; rec[3] = r4.w; }
		else
			for (int e = 0; e < 4; ++e) if (i + e < count) rec[e] = R[i + e];
		u32 mine = 0;
#pragma unroll
		for (int e = 0; e < 4; ++e) { const u32 k = rec[e] & 7u; mine += (k == 1u ? 1u : 0u) + (k == 2u ? 1u << 16 : 0u); }
		u32 incl = mine;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u32 v = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += v;
		}
		if (lane == 31) { wtot[warp][0] = incl & 0xffffu; wtot[warp][1] = incl >> 16; }
		__syncthreads();
		if (threadIdx.x < 2)
		{
			u32 tot = 0;
			for (int w = 0; w < 8; ++w) tot += wtot[w][threadIdx.x];
			base[threadIdx.x] = tot ? atomicAdd(a.cnt + threadIdx.x, tot) : 0u;
		}
		__syncthreads();
		const u32 excl = incl - mine;
		u32 s1 = base[0] + (excl & 0xffffu), s2 = base[1] + (excl >> 16);
		for (int w = 0; w < warp; ++w) { s1 += wtot[w][0]; s2 += wtot[w][1]; }
#pragma unroll
		for (int e = 0; e < 4; ++e)
		{
			const u32 k = rec[e] & 7u;
			if (k == 1u || k == 2u)
			{
				const int2 np = V[i + e];
				const int2 tagged = make_int2(np.x | (int)(((rec[e] >> 3) & 3u) << kFlagShift), np.y);
				if (k == 1u) { if (s1 < a.cap_p2p) a.p2p[s1] = tagged; ++s1; }
				else { if (s2 < a.cap_m2l) a.m2l[s2] = tagged; ++s2; }
			}
		}
		__syncthreads();
	}
}

// end of every evaluation: fold "a list or a frontier did not fit" into a STICKY word (cnt[24]) that survives the
// per-evaluation counter resets, so that evaluations which are only enqueued can be checked later (fmm3_harvest)
__global__ void eval_check_kernel(u32 *cnt, u32 cap_list)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
		if (cnt[0] > cap_list || cnt[1] > cap_list || cnt[5]) cnt[24] = 1u;
}

__global__ void traverse_init_kernel(int2 *front, u32 *cnt)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		front[0] = make_int2(0, 0);
		cnt[0] = 0; cnt[1] = 0; cnt[2] = 1; cnt[3] = 0; cnt[4] = 0; cnt[5] = 0; cnt[6] = 0;
	}
}

// =====================================================================================
//  interaction lists bucketed by target (CsrView, fmm3_common.cuh): count, scan, fill, sort
// =====================================================================================
// target id of the directed interactions of list entry w: M2L entries come first, then the P2P entries
struct CsrArgs
{
	const int2 *m2l, *p2p;
	const u32 *cnt;          // [0] p2p pairs, [1] m2l pairs
	u32 cap_list;
	u32 *deg, *off, *bsum;
	int *src;
	u32 cap_src;
	int ntot, nrows, leaf_beg; // nrows = ntot + 2^L
	u32 *sticky;             // cnt + 5: something did not fit
};

template <bool FILL>
__global__ void __launch_bounds__(256) csr_count_fill_kernel(CsrArgs a)
{
	const u32 nm = min(a.cnt[1], a.cap_list), np = min(a.cnt[0], a.cap_list);
	for (u32 w = blockIdx.x * blockDim.x + threadIdx.x; w < nm + np; w += gridDim.x * blockDim.x)
	{
		const bool is_m = w < nm;
		int2 pr = is_m ? a.m2l[w] : a.p2p[w - nm];
		const int flags = (pr.x >> kFlagShift) & 3; // bit 0: pr.x is a target of this rank, bit 1: pr.y
		pr.x &= kNodeMask;
		const int rowx = is_m ? pr.x : a.ntot + (pr.x - a.leaf_beg), rowy = is_m ? pr.y : a.ntot + (pr.y - a.leaf_beg);
		if (flags & 1)
		{
			if (!FILL) atomicAdd(a.deg + rowx, 1u);
			else
			{
				// the counters run back to zero while the rows fill: ready for the next evaluation without a memset
				const u32 pos = a.off[rowx] + (atomicSub(a.deg + rowx, 1u) - 1u);
				if (pos < a.cap_src) a.src[pos] = pr.y;
			}
		}
		if (flags & 2)
		{
			if (!FILL) atomicAdd(a.deg + rowy, 1u);
			else
			{
				const u32 pos = a.off[rowy] + (atomicSub(a.deg + rowy, 1u) - 1u);
				if (pos < a.cap_src) a.src[pos] = pr.x;
			}
		}
	}
}

constexpr int kScanBlock = 1024, kScanPer = 4, kScanTile = kScanBlock * kScanPer;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *wsum /* 32 */, u32 &total)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	u32 incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
	__syncthreads(); // wsum may still be read from a previous call
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	u32 base = 0, tot = 0;
	for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { const u32 x = wsum[i]; if (i < w) base += x; tot += x; }
	total = tot;
	return base + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) csr_scan_sums_kernel(CsrArgs a)
{
	__shared__ u32 wsum[32];
	const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
	u32 s = 0;
#pragma unroll
	for (int k = 0; k < kScanPer; ++k) if (i0 + k < a.nrows) s += a.deg[i0 + k];
	u32 total;
	block_exclusive_scan(s, wsum, total);
	if (threadIdx.x == 0) a.bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) csr_scan_top_kernel(CsrArgs a, int nblocks)
{
	__shared__ u32 wsum[32];
	__shared__ u32 carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int b0 = 0; b0 < nblocks; b0 += kScanBlock)
	{
		const int b = b0 + threadIdx.x;
		const u32 v = b < nblocks ? a.bsum[b] : 0u;
		u32 total;
		const u32 ex = block_exclusive_scan(v, wsum, total);
		const u32 c = carry;
		if (b < nblocks) a.bsum[b] = c + ex;
		__syncthreads();
		if (threadIdx.x == 0) carry = c + total;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		a.off[a.nrows] = carry;
		if (carry > a.cap_src) *a.sticky = 1u; // the host grows the row storage and repeats the evaluation
	}
}

__global__ void __launch_bounds__(kScanBlock) csr_scan_apply_kernel(CsrArgs a)
{
	__shared__ u32 wsum[32];
	const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
	u32 d[kScanPer], s = 0;
#pragma unroll
	for (int k = 0; k < kScanPer; ++k) { d[k] = i0 + k < a.nrows ? a.deg[i0 + k] : 0u; s += d[k]; }
	u32 total;
	u32 run = a.bsum[blockIdx.x] + block_exclusive_scan(s, wsum, total);
#pragma unroll
	for (int k = 0; k < kScanPer; ++k) { if (i0 + k < a.nrows) a.off[i0 + k] = run; run += d[k]; }
}

// rows in ascending source order: the sums of the downward pass do not depend on the order in which the traversal
// emitted the pairs (run-to-run reproducible forces).  Rows are short (a few entries; at most some tens).
__global__ void __launch_bounds__(256) csr_sort_kernel(CsrArgs a)
{
	for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < a.nrows; row += gridDim.x * blockDim.x)
	{
		const u32 b = a.off[row], e = min(a.off[row + 1], a.cap_src);
		for (u32 i = b + 1; i < e; ++i)
		{
			const int v = a.src[i];
			u32 j = i;
			for (; j > b && a.src[j - 1] > v; --j) a.src[j] = a.src[j - 1];
			a.src[j] = v;
		}
	}
}

// =====================================================================================
//  near field (replaces fmm_p2p3_kdtree_coalesced / _self_, :874-959,1048-1120)
// =====================================================================================
__device__ __forceinline__ float rsqrt_approx(float x)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}

// a group of G lanes handles one leaf pair, both directions (the intra-leaf part lives in the L2P kernel)
template <int G>
__global__ void __launch_bounds__(256)
p2p_kernel(const int2 *__restrict__ list, const u32 *__restrict__ count, u32 cap, const float *__restrict__ spos,
           float *__restrict__ acc, int64_t n, int L, float eps2, PeerTab peers, u32 *__restrict__ nearbits)
{
	const int lane = threadIdx.x & (G - 1);
	const int groups = (gridDim.x * blockDim.x) / G;
	const int beg = kd_beg(L);
	const u32 npairs = min(*count, cap);
	const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
	const u32 nwork = ((npairs + (32 / G) - 1) / (32 / G)) * (32 / G); // keep whole warps in the loop
	for (u32 w = (blockIdx.x * blockDim.x + threadIdx.x) / G; w < nwork; w += groups)
	{
		if (w >= npairs) continue;
		const int2 np = list[w];
		const int flags = (np.x >> kFlagShift) & 3, l1 = (np.x & kNodeMask) - beg, l2 = np.y - beg;
		const int64_t i1 = seg_start(n, l1, L), i2 = seg_start(n, l2, L);
		const int m1 = (int)(seg_start(n, l1 + 1, L) - i1), m2 = (int)(seg_start(n, l2 + 1, L) - i2);
#pragma unroll 1
		for (int dir = 0; dir < 2; ++dir)
		{
			if (!((flags >> dir) & 1)) continue; // another rank owns these targets
			const int64_t ti = dir ? i2 : i1, si = dir ? i1 : i2;
			const int tm = dir ? m2 : m1, sm = dir ? m1 : m2;
			if (nearbits && lane == 0) { const int tl = dir ? l2 : l1; atomicOr(nearbits + (tl >> 5), 1u << (tl & 31)); } // sparse near field: rows of this leaf are live
			// multi-GPU: the particles of a remote source leaf are read from their owner's published positions
			const float *__restrict__ src = spos;
			if (peers.g > 0)
			{
				const int o = (dir ? l1 : l2) >> (L - peers.g);
				if (o != peers.me) src = peers.pos[o];
			}
			for (int h0 = 0; h0 < tm; h0 += G)
			{
				const int h = h0 + lane;
				const bool hv = h < tm;
				const float *tp = spos + 3 * (ti + (hv ? h : 0));
				const float x = tp[0], y = tp[1], z = tp[2];
				float ax = 0.f, ay = 0.f, az = 0.f;
				for (int g0 = 0; g0 < sm; g0 += G)
				{
					const int gl = g0 + lane;
					const float *sp = src + 3 * (si + (gl < sm ? gl : 0));
					const float sx = sp[0], sy = sp[1], sz = sp[2];
					const int ng = min(G, sm - g0);
					for (int g = 0; g < ng; ++g)
					{
						float dx = x - __shfl_sync(gmask, sx, g, G);
						float dy = y - __shfl_sync(gmask, sy, g, G);
						float dz = z - __shfl_sync(gmask, sz, g, G);
						float r2 = fmaf(dx, dx, eps2);
						r2 = fmaf(dy, dy, r2);
						r2 = fmaf(dz, dz, r2);
						float wv = rsqrt_approx(r2);
						wv = wv * fmaf(-0.5f * r2 * wv, wv, 1.5f); // one Newton step: error of the near field ~1e-7
						float w3 = (wv * wv) * wv;
						ax = fmaf(dx, w3, ax); ay = fmaf(dy, w3, ay); az = fmaf(dz, w3, az);
					}
				}
				if (hv)
				{
					float *o = acc + 3 * (ti + h);
					atomicAdd(o, ax); atomicAdd(o + 1, ay); atomicAdd(o + 2, az);
				}
			}
		}
	}
}

} // namespace

// =====================================================================================
//  host side: plan, phases, introspection
// =====================================================================================
enum Phase { PH_KDTOP = 0, PH_KDBOTTOM, PH_PERMUTE, PH_UPWARD, PH_TRAVERSE, PH_P2P, PH_M2L, PH_L2L, PH_L2P, PH_COUNT };
// single-kernel phases: kd_bottom, p2p, m2l, l2p (their CUDA-event times are kernel launch durations)
// by-target orders: "p2p" = bucketing the lists by target, "m2l" is empty (the M2L sums are gathered inside the "l2l" levels),
// "l2p" includes the near field; pair-list orders (7..10): the phases are what their names say
static const char *kPhaseNames[PH_COUNT] = {"kd_top", "kd_bottom", "permute", "p2m_m2m", "traverse", "p2p", "m2l", "l2l", "l2p"};

struct FmmPlan
{
	int64_t n = 0;
	int L = 0, ntot = 0, order = 0, offM = 0, offL = 0, sM = 0, sL = 0, mlt_max = 0;
	int counter = 0, rebuilt = 0, max_level = -1;
	int64_t rec_n = 0;        // records of the last evaluation's traversal (cnt[11]); rec_fallbacks: reuse evaluations that went back to the root
	u32 rec_fallbacks = 0;
	int leaf_pending = 0; // the last downward pass may have kept the leaf level of the L2L pass in registers (OrderOps::finish_leaf_locals)
	float dens = 0.f;
	int64_t p2p_n = 0, m2l_n = 0;
	u32 cap_list = 0, cap_front = 0;
	KdTree kd;
	DevBuf center, mpole, local, tmp3, accn, nearbits;
	bool accn_zero = false; // accn is known to be all zero (the sparse near-field path leaves it that way)
	DevBuf p2p, m2l, frontA, frontB, cnt, mfac;
	DevBuf csr_deg, csr_off, csr_src, csr_bsum;   // lists bucketed by target (cfg.reproducible)
	u32 cap_src = 0;
	bool csr_ready = false;
	// phase timers: a ring of event sets, one per evaluation in flight.  Evaluations between two tree rebuilds are only
	// ENQUEUED (no host synchronisation); their event sets are read at the next synchronisation point (fmm3_harvest)
	static constexpr int kEvRing = 16;
	cudaEvent_t evr[kEvRing][PH_COUNT + 1];
	int ev_head = 0, ev_npend = 0;      // sets [ev_head - ev_npend, ev_head) (mod kEvRing) await harvesting
	bool ev_rebuild[kEvRing] = {};
	cudaEvent_t *ev = nullptr;          // the set of the evaluation being enqueued
	float last_ms[PH_COUNT] = {};
	bool ev_ok = false, ev_valid = false;
	double tot_ms[PH_COUNT] = {};
	int64_t tot_evals = 0, tot_rebuilds = 0;
	int coop_blocks = -1; // grid of the cooperative traversal kernel (0 = not available)
	int queue_blocks = 0; // grid of the work-queue traversal kernel (0 = use the rounds)
	bool queue_clean = false; // frontA holds nothing but the empty-slot sentinel
	int rec_blocks = 0;       // grid of the recording traversal kernel (0 = incremental traversal unavailable)
	bool rec_valid = false;   // V / R (frontA / frontB) describe the traversal of the previous evaluation
	float rec_radius = 0.f; int rec_m2l_first = -1, rec_rank = -1, rec_world = -1;
};

static int plan_levels(int64_t n, int order, float dens, int max_level)
{
	// fmm_cart3_kdtree.cuh:1507-1516
	float s = (float)(order * order);
	int L = max_level == 0 ? (int)std::round(std::log2(dens * (float)n / s)) : max_level;
	L = std::min(std::max(L, 2), 30);
	while ((1ll << L) > n) --L;
	return L;
}

static int ensure_plan(nbco_ctx *ctx, int64_t n)
{
	const nbco_config &c = ctx->cfg;
	if (!ctx->fmm) ctx->fmm = new FmmPlan();
	FmmPlan &p = *ctx->fmm;
	if (!p.ev_ok)
	{
		for (int k = 0; k < FmmPlan::kEvRing; ++k)
			for (int i = 0; i <= PH_COUNT; ++i) NBCO_CUDA(cudaEventCreate(&p.evr[k][i]));
		p.ev = p.evr[0];
		p.ev_ok = true;
	}
	if (p.n == n && p.order == c.order && p.dens == c.dens_inhom && p.max_level == c.max_level && (p.csr_ready || !c.reproducible)) return NBCO_OK;
	if (ctx->peer.active) { set_error("n / order / levels cannot change while peers are attached (published buffers would move)"); return NBCO_ERR_INVALID; }
	if (n < 8) { set_error("fmm3_kd needs n >= 8 (got %lld)", (long long)n); return NBCO_ERR_INVALID; }
	if (n >= (1ll << 31)) { set_error("n must be < 2^31"); return NBCO_ERR_INVALID; }
	const int L = plan_levels(n, c.order, c.dens_inhom, c.max_level);
	if (L < 1) { set_error("tree depth %d", L); return NBCO_ERR_INVALID; }
	p.n = n; p.order = c.order; p.dens = c.dens_inhom; p.max_level = c.max_level;
	p.L = L; p.ntot = (1 << (L + 1)) - 1;
	p.offM = c.order * (c.order + 1) * (c.order + 2) / 6; p.offL = (c.order + 1) * (c.order + 1);
	p.sM = (p.offM + 3) & ~3; p.sL = (p.offL + 3) & ~3;
	p.mlt_max = (int)((n - 1) / (1ll << L) + 1);
	p.counter = 0;
	const size_t nt = (size_t)p.ntot;
	NBCO_TRY(kd_reserve(ctx, p.kd, n, L));
	NBCO_TRY(p.center.reserve(16 * nt));
	NBCO_TRY(p.mpole.reserve(4 * nt * p.sM)); NBCO_TRY(p.local.reserve(4 * nt * p.sL));
	NBCO_TRY(p.tmp3.reserve(12 * (size_t)n)); NBCO_TRY(p.accn.reserve(12 * (size_t)n));
	p.accn_zero = false;
	if (p.cap_list < (u32)std::min<int64_t>(8ll * p.ntot + 1024, 0x7fffffff))
	{
		p.cap_list = (u32)std::min<int64_t>(8ll * p.ntot + 1024, 0x7fffffff);
		p.cap_front = p.cap_list;
	}
	NBCO_TRY(p.p2p.reserve(8 * (size_t)p.cap_list)); NBCO_TRY(p.m2l.reserve(8 * (size_t)p.cap_list));
	NBCO_TRY(p.frontA.reserve(8 * (size_t)p.cap_front)); NBCO_TRY(p.frontB.reserve(8 * (size_t)p.cap_front));
	{
		// by-target rows (cfg.reproducible): ntot + 2^L targets
		const size_t nrows = nt + ((size_t)1 << L);
		// storage: every pair of the two lists yields at most two directed entries
		p.cap_src = (u32)std::min<int64_t>(4ll * p.cap_list, 0x7fffffff);
		if (c.reproducible)
		{
			NBCO_TRY(p.csr_deg.reserve(4 * (nrows + 1))); NBCO_TRY(p.csr_off.reserve(4 * (nrows + 1)));
			NBCO_TRY(p.csr_src.reserve(4 * (size_t)p.cap_src));
			NBCO_TRY(p.csr_bsum.reserve(4 * ((nrows + kScanTile - 1) / kScanTile + 1)));
			NBCO_CUDA(cudaMemsetAsync(p.csr_deg.p, 0, 4 * (nrows + 1), ctx->stream));
		}
		p.csr_ready = c.reproducible != 0;
	}
	p.queue_clean = false; p.rec_valid = false;
	NBCO_TRY(p.cnt.reserve(128)); NBCO_TRY(p.mfac.reserve(4 * 2 * 40));
	NBCO_CUDA(cudaMemsetAsync(p.cnt.p, 0, 128, ctx->stream)); // cnt[24] (sticky overflow) is never reset by the kernels
	// MAC factor table: M = pow(mult / N, 1/(3p+6)) evaluated with the host libm like the
	// reference CPU path (:410); a node of level l holds floor(n/2^l) or floor(n/2^l)+1 particles
	float tab[2 * 32];
	for (int l = 0; l <= L; ++l)
	{
		int lo = (int)(n >> l);
		tab[2*l]     = powf((float)lo / (float)(int)n, 1.f / (3 * c.order + 6));
		tab[2*l + 1] = powf((float)(lo + 1) / (float)(int)n, 1.f / (3 * c.order + 6));
	}
	NBCO_CUDA(cudaMemcpyAsync(p.mfac.p, tab, sizeof(float) * 2 * (L + 1), cudaMemcpyHostToDevice, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	return NBCO_OK;
}

bool fmm3_next_rebuilds(nbco_ctx *ctx, int64_t n)
{
	const nbco_config &c = ctx->cfg;
	if (!ctx->fmm) return true;
	const FmmPlan &p = *ctx->fmm;
	if (p.n != n || p.order != c.order || p.dens != c.dens_inhom || p.max_level != c.max_level) return true;
	return c.unsort || (p.counter % c.tree_steps == 0);
}

bool fmm3_rebuilds_in(nbco_ctx *ctx, int64_t n, int64_t k)
// will the evaluation that comes k evaluations after the next one (k = 0: the next one) permute pos / vel?
{
	const nbco_config &c = ctx->cfg;
	if (!ctx->fmm) return true;
	const FmmPlan &p = *ctx->fmm;
	if (p.n != n || p.order != c.order || p.dens != c.dens_inhom || p.max_level != c.max_level) return true;
	return c.unsort || ((p.counter + k) % c.tree_steps == 0);
}

int fmm3_peer_buffers(nbco_ctx *ctx, int64_t n, void **center, void **mpole)
{
	NBCO_TRY(ensure_plan(ctx, n));
	*center = ctx->fmm->center.p; *mpole = ctx->fmm->mpole.p;
	return NBCO_OK;
}

#define LAUNCHED(ctx) do { ++(ctx)->launches; } while (0)

static int run_phases(const OrderOps &ops, nbco_ctx *ctx, FmmPlan &p, float *d_pos, float *d_acc, const float *d_param, bool fuse_elastic, bool rebuild)
{
	cudaStream_t st = ctx->stream;
	const int64_t n = p.n;
	const int L = p.L;
	const nbco_config &c = ctx->cfg;
	TreeData t{p.center.as<float4>(), p.kd.size2.as<float>(), p.mpole.as<float>(), p.local.as<float>(), p.sM, p.sL, {}};
	// multi-GPU: rank r of 2^g owns the subtree of node (g, r): a contiguous range of leaves and particles
	int g = 0;
	while ((1 << g) < c.world) ++g;
	PeerState &ps = ctx->peer;
	const bool peer = ps.active && c.world > 1;
	if (peer && ps.n != n) { set_error("peer mode was set up for n = %lld", (long long)ps.n); return NBCO_ERR_INVALID; }
	// peer mode: only the own subtree is built and summarised here (pg = g); replicated mode: everything (pg = 0)
	const int pg = peer ? g : 0, pr = peer ? c.rank : 0;
	for (int q = 0; q < kMaxPeers; ++q) { t.peers.center[q] = t.center; t.peers.mpole[q] = t.mpole; t.peers.pos[q] = d_pos; }
	t.peers.g = pg; t.peers.me = pr;
	if (peer)
		for (int q = 0; q < c.world; ++q)
		{
			t.peers.center[q] = (const float4 *)ps.center[q]; t.peers.mpole[q] = (const float *)ps.mpole[q];
			t.peers.pos[q] = (const float *)((const char *)ps.pubp[q] + kPeerData);
		}
	const int64_t own_lo = seg_start(n, pr, pg), own_hi = seg_start(n, pr + 1, pg);

	NBCO_CUDA(cudaEventRecord(p.ev[PH_KDTOP], st));
	bool pulled = false;
	const float *spos = d_pos; // tree-ordered positions the passes read
	if (!rebuild)
	{
		NBCO_CUDA(cudaEventRecord(p.ev[PH_KDBOTTOM], st));
		NBCO_CUDA(cudaEventRecord(p.ev[PH_PERMUTE], st));
	}
	else
	{
		// small trees: fewer shared-memory kd blocks than ranks -> every rank builds everything (needs all positions)
		const int bg = pg <= p.kd.lt ? pg : 0;
		const bool distributed = peer && bg > 0 && !getenv("NBCO_PEER_PULL_ALL");
		if (distributed)
		{
			// Distributed build (kdtree.cu): every rank contributes the records of its own range; the first g levels are
			// selected from summed histograms, then every rank pulls the records of its subtree.  The velocities of the
			// particles that end up here are fetched afterwards from the published velocity ranges.
			NBCO_TRY(peer_publish(ctx, d_pos + 3*n, 1, n));
			NBCO_TRY(kd_build_peer(ctx, p.kd, d_pos, p.ev[PH_KDBOTTOM], pr, pg, ps.pubp));
			pulled = true;
		}
		else
		{
			if (peer && !ps.have_full)
			{
				// every rank holds only its own range: publish it, pull the others' (this path splits ALL particles on every rank)
				NBCO_TRY(peer_publish(ctx, d_pos, 0, n)); NBCO_TRY(peer_publish(ctx, d_pos + 3*n, 1, n));
				NBCO_TRY(peer_barrier(ctx));
				NBCO_TRY(peer_pull(ctx, d_pos, 0, n));
				NBCO_TRY(peer_barrier(ctx)); // the position mirrors are rewritten below
				pulled = true;
			}
			NBCO_TRY(kd_build(ctx, p.kd, d_pos, p.ev[PH_KDBOTTOM], bg ? pr : 0, bg));
		}
		NBCO_CUDA(cudaEventRecord(p.ev[PH_PERMUTE], st));
		if (c.unsort)
			spos = p.kd.spos.as<float>();
		else
		{
			// leave pos and the velocities behind it in tree order (:1359-1360,1758-1759); peer mode: own range only
			const int64_t cnt = own_hi - own_lo;
			NBCO_CUDA(cudaMemcpyAsync(d_pos + 3*own_lo, p.kd.spos.as<float>() + 3*own_lo, 12 * (size_t)cnt, cudaMemcpyDeviceToDevice, st));
			if (pulled)
			{
				VelSrc vs;
				for (int q = 0; q < kMaxPeers; ++q)
					vs.v[q] = q < c.world ? (const float *)((const char *)ps.pubp[q] + kPeerData) + 3 * (size_t)n : nullptr;
				vs.v[pr] = d_pos + 3*n; vs.g = pg; vs.me = pr;
				vs.dbg = getenv("NBCO_DEBUG_KD") ? (u32 *)((char *)ps.pub.p + 768) : nullptr;
				gather3_peer_kernel<<<grid_for(cnt, 256, ctx->sm_count, 8), 256, 0, st>>>(vs, p.kd.perm.as<int>() + own_lo, p.tmp3.as<float>() + 3*own_lo, cnt, n);
			}
			else
				gather3_kernel<<<grid_for(cnt, 256, ctx->sm_count, 8), 256, 0, st>>>(d_pos + 3*n, p.kd.perm.as<int>() + own_lo, p.tmp3.as<float>() + 3*own_lo, cnt);
			LAUNCHED(ctx);
			NBCO_CUDA(cudaMemcpyAsync(d_pos + 3*n + 3*own_lo, p.tmp3.as<float>() + 3*own_lo, 12 * (size_t)cnt, cudaMemcpyDeviceToDevice, st));
			if (ctx->ids && !peer)
			{
				// optional identity array (nbco_track_ids): the same permutation as pos / vel
				gather1_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, st>>>(ctx->ids, p.kd.perm.as<int>(), p.tmp3.as<int>(), n);
				LAUNCHED(ctx);
				NBCO_CUDA(cudaMemcpyAsync(ctx->ids, p.tmp3.p, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
			}
		}
		if (peer) ps.have_full = false;
	}
	NBCO_CUDA(cudaEventRecord(p.ev[PH_UPWARD], st));
	// by-target orders write every local exactly once (no zero-fill); pair-list orders accumulate with atomics
	// (opt-in: cfg.reproducible.  Measured on B200 at N = 2^24: the gather per target is load-imbalanced -- rows hold 0..30
	// sources -- and costs 2.3 ms per evaluation against 0.84 ms for the pair kernels: profiles/r02_notes.md)
	const bool by_target = ops.by_target && c.reproducible;
	if (!by_target)
	{
		if (peer && pg > 0)
		{
			NBCO_CUDA(cudaMemsetAsync(p.local.p, 0, sizeof(float) * (size_t)kd_beg(pg) * p.sL, st)); // replicated top levels
			const int64_t work = ((((int64_t)1 << (L - pg + 1)) - 1) * (p.sL / 4));
			zero_own_locals_kernel<<<grid_for(work, 256, ctx->sm_count, 8), 256, 0, st>>>(p.local.as<float4>(), p.sL / 4, L, pr, pg);
			LAUNCHED(ctx);
		}
		else
			NBCO_CUDA(cudaMemsetAsync(p.local.p, 0, sizeof(float) * (size_t)p.ntot * p.sL, st));
	}
	if (peer) NBCO_TRY(peer_publish(ctx, d_pos, 0, n));
	ops.upward(ctx, t, spos, n, L, pr, pg, 0);
	if (peer)
	{
		NBCO_TRY(peer_barrier(ctx)); // every subtree is summarised and every position range published
		ops.upward(ctx, t, spos, n, L, pr, pg, 1);
	}

	NBCO_CUDA(cudaEventRecord(p.ev[PH_TRAVERSE], st));
	TravArgs a;
	a.center = t.center; a.mfac = p.mfac.as<float>();
	a.p2p = p.p2p.as<int2>(); a.m2l = p.m2l.as<int2>();
	a.cnt = p.cnt.as<u32>();
	a.cap_p2p = a.cap_m2l = p.cap_list; a.cap_front = p.cap_front;
	a.n = n; a.ntot = p.ntot; a.L = L; a.m2l_first = c.m2l_first; a.radius = c.radius;
	a.sh_lo = seg_start(n, c.rank, g); a.sh_hi = seg_start(n, c.rank + 1, g);
	a.peers = t.peers;
	// breadth first: a pair is split at most once per round, 2L + 2 rounds always suffice.  (A depth-first
	// tail with private stacks was measured 2-25x slower on B200, profiles/r01_notes.md.)
	const int rounds = 2 * L + 2;
	if (p.coop_blocks < 0)
	{
		int coop = 0, per_sm = 0, per_sm_q = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->cfg.device);
		if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_all_kernel, 256, 0) == cudaSuccess && per_sm > 0)
			p.coop_blocks = ctx->sm_count * std::min(per_sm, 4);
		else
			p.coop_blocks = 0;
		if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_q, traverse_queue_kernel, 256, 0) == cudaSuccess && per_sm_q > 0)
			p.queue_blocks = ctx->sm_count * std::min(per_sm_q, 2);
		// the work queue wins while a rank visits a moderate number of pairs (measured: 0.10 vs 0.15 ms at 64 k
		// particles, 0.72 vs 0.41 ms at 16 M on one GPU); NBCO_TRAVERSE=rounds|queue overrides
		int per_sm_r = 0;
		if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_r, traverse_rec_kernel, 256, 0) == cudaSuccess && per_sm_r > 0)
			p.rec_blocks = ctx->sm_count * std::min(per_sm_r, 4);
		const char *mode = getenv("NBCO_TRAVERSE");
		if (mode && (!strcmp(mode, "rounds") || !strcmp(mode, "queue") || !strcmp(mode, "launches"))) p.rec_blocks = 0;
		if (n / c.world > (1ll << 22) && !(mode && !strcmp(mode, "queue"))) p.queue_blocks = 0;
		if (mode && !strcmp(mode, "rounds")) p.queue_blocks = 0;
		// one plain launch per round: cooperative grids do not overlap with other streams' kernels, which
		// deadlocks several ranks emulated on ONE device against each other's barrier kernels (tests only)
		if (mode && !strcmp(mode, "launches")) { p.queue_blocks = 0; p.coop_blocks = 0; }
	}
	// incremental traversal (default; NBCO_TRAVERSE=rounds|queue|launches select a from-the-root kernel instead)
	const bool incremental = p.rec_blocks > 0 && c.tree_steps > 1 && !c.unsort;
	if (incremental)
	{
		int2 *V = p.frontA.as<int2>();
		u32 *R = p.frontB.as<u32>(), *clist = R + p.cap_front;
		const u32 half = p.cap_front / 2;
		const bool reuse = p.rec_valid && !rebuild && p.rec_radius == c.radius && p.rec_m2l_first == c.m2l_first
		                   && p.rec_rank == c.rank && p.rec_world == c.world;
		p.queue_clean = false;
		if (!reuse)
		{
			traverse_rec_init_kernel<<<1, 32, 0, st>>>(V, a.cnt); LAUNCHED(ctx);
		}
		else
		{
			// NBCO_REC_LIMIT=<records>: force the from-the-root fallback earlier (tests)
			const char *lim_s = getenv("NBCO_REC_LIMIT");
			const long long lim_env = lim_s ? atoll(lim_s) : -1;
			const u32 limit = lim_env >= 0 ? (u32)std::min<long long>(lim_env, half) : half;
			traverse_reuse_init_kernel<<<1, 32, 0, st>>>(a.cnt, V, limit); LAUNCHED(ctx);
			reval_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(a, V, R, clist, half); LAUNCHED(ctx);
			{
				u32 *bq = clist + half;
				u32 cap_q = half;
				void *rargs[] = {&a, &R, &clist, &bq, &cap_q};
				NBCO_CUDA(cudaLaunchCooperativeKernel((void *)retire_kernel, dim3(ctx->sm_count * 2), dim3(256), rargs, 0, st));
				LAUNCHED(ctx);
			}
			emit_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(a, V, R); LAUNCHED(ctx);
		}
		int max_rounds = rounds;
		const u32 *seeds = reuse ? clist + half : nullptr;
		void *args[] = {&a, &V, &R, &seeds, &max_rounds};
		NBCO_CUDA(cudaLaunchCooperativeKernel((void *)traverse_rec_kernel, dim3(p.rec_blocks), dim3(256), args, 0, st));
		LAUNCHED(ctx);
		p.rec_valid = true; p.rec_radius = c.radius; p.rec_m2l_first = c.m2l_first; p.rec_rank = c.rank; p.rec_world = c.world;
	}
	else if (p.queue_blocks > 0)
	{
		p.rec_valid = false;
		// the queue lives in frontA; it is all-SENT between evaluations (consumers restore what they take)
		unsigned long long *q = p.frontA.as<unsigned long long>();
		if (!p.queue_clean)
		{
			NBCO_CUDA(cudaMemsetAsync(q, 0xff, 8 * (size_t)p.cap_front, st));
			p.queue_clean = true;
		}
		traverse_queue_init_kernel<<<1, 32, 0, st>>>(q, a.cnt); LAUNCHED(ctx);
		void *args[] = {&a, &q};
		NBCO_CUDA(cudaLaunchCooperativeKernel((void *)traverse_queue_kernel, dim3(p.queue_blocks), dim3(256), args, 0, st));
		LAUNCHED(ctx);
	}
	else if (p.coop_blocks > 0)
	{
		traverse_init_kernel<<<1, 32, 0, st>>>(p.frontA.as<int2>(), a.cnt); LAUNCHED(ctx);
		p.queue_clean = false; p.rec_valid = false;
		int2 *fa = p.frontA.as<int2>(), *fb = p.frontB.as<int2>();
		int max_rounds = rounds;
		void *args[] = {&a, &fa, &fb, &max_rounds};
		NBCO_CUDA(cudaLaunchCooperativeKernel((void *)traverse_all_kernel, dim3(p.coop_blocks), dim3(256), args, 0, st));
		LAUNCHED(ctx);
	}
	else
	{
		traverse_init_kernel<<<1, 32, 0, st>>>(p.frontA.as<int2>(), a.cnt); LAUNCHED(ctx);
		p.queue_clean = false; p.rec_valid = false;
		for (int r = 0; r < rounds; ++r)
		{
			a.front_in = (r & 1) ? p.frontB.as<int2>() : p.frontA.as<int2>();
			a.front_out = (r & 1) ? p.frontA.as<int2>() : p.frontB.as<int2>();
			traverse_round_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(a, r);
			LAUNCHED(ctx);
		}
	}

	float *accn = p.accn.as<float>();
	if (by_target)
	{
		// ---- lists bucketed by target: every local / acceleration becomes one sum in registers (no atomics) ----
		NBCO_CUDA(cudaEventRecord(p.ev[PH_P2P], st));
		CsrArgs ca;
		ca.m2l = a.m2l; ca.p2p = a.p2p; ca.cnt = a.cnt; ca.cap_list = p.cap_list;
		ca.deg = p.csr_deg.as<u32>(); ca.off = p.csr_off.as<u32>(); ca.bsum = p.csr_bsum.as<u32>(); ca.src = p.csr_src.as<int>();
		ca.cap_src = p.cap_src; ca.ntot = p.ntot; ca.nrows = p.ntot + (1 << L); ca.leaf_beg = kd_beg(L); ca.sticky = a.cnt + 5;
		const int nsb = (ca.nrows + kScanTile - 1) / kScanTile;
		csr_count_fill_kernel<false><<<ctx->sm_count * 8, 256, 0, st>>>(ca);
		csr_scan_sums_kernel<<<nsb, kScanBlock, 0, st>>>(ca);
		csr_scan_top_kernel<<<1, kScanBlock, 0, st>>>(ca, nsb);
		csr_scan_apply_kernel<<<nsb, kScanBlock, 0, st>>>(ca);
		csr_count_fill_kernel<true><<<ctx->sm_count * 8, 256, 0, st>>>(ca);
		csr_sort_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ca);
		ctx->launches += 6;
		NBCO_CUDA(cudaEventRecord(p.ev[PH_M2L], st));
		NBCO_CUDA(cudaEventRecord(p.ev[PH_L2L], st));
		const CsrView cv{ca.off, ca.src, p.ntot, p.cap_src};
		ops.downward(ctx, t, spos, nullptr, d_acc, c.unsort ? p.kd.perm.as<int>() : nullptr, d_param, fuse_elastic ? 1 : 0, n, L, c.rank, g, c.eps2, c.coll,
		             p.ev[PH_L2P], &cv);
		p.leaf_pending = 0;
	}
	else
	{
	// sparse near field (OrderOps::sparse_near): far fewer leaf pairs than leaves at the headline size, so the pair kernel
	// marks the leaves it adds to and the L2P kernel reads and re-zeroes only their rows: 2^L bits are cleared per evaluation
	// instead of 12 n bytes, and acc_near is read only where it is not zero
	u32 *nearbits = nullptr;
	if (ops.sparse_near && ops.sparse_near(n, L) && !getenv("NBCO_DENSE_NEAR"))
	{
		const size_t words = (((size_t)1 << L) + 31) / 32;
		NBCO_TRY(p.nearbits.reserve(4 * words));
		nearbits = p.nearbits.as<u32>();
		NBCO_CUDA(cudaMemsetAsync(nearbits, 0, 4 * words, st));
		if (!p.accn_zero) NBCO_CUDA(cudaMemsetAsync(accn, 0, 12 * (size_t)n, st));
	}
	else
		NBCO_CUDA(cudaMemsetAsync(accn, 0, 12 * (size_t)n, st));
	p.accn_zero = false; // until the L2P kernel of this evaluation is enqueued
	t.nearbits = nearbits;
	NBCO_CUDA(cudaEventRecord(p.ev[PH_P2P], st)); // the p2p phase is exactly one kernel
	if (c.coll)
	{
		const int blocks = ctx->sm_count * 8;
#define P2P_LAUNCH(G)                                                                                              \
		do {                                                                                                       \
			p2p_kernel<G><<<blocks, 256, 0, st>>>(a.p2p, a.cnt + 0, a.cap_p2p, spos, accn, n, L, c.eps2, t.peers, nearbits); \
		} while (0)
		if (p.mlt_max <= 4) P2P_LAUNCH(4);
		else if (p.mlt_max <= 8) P2P_LAUNCH(8);
		else if (p.mlt_max <= 16) P2P_LAUNCH(16);
		else P2P_LAUNCH(32);
#undef P2P_LAUNCH
		ctx->launches += 1;
	}

	NBCO_CUDA(cudaEventRecord(p.ev[PH_M2L], st));
	ops.m2l(ctx, t, a.m2l, a.cnt + 1, a.cap_m2l, c.eps2);

	NBCO_CUDA(cudaEventRecord(p.ev[PH_L2L], st));
	ops.downward(ctx, t, spos, accn, d_acc, c.unsort ? p.kd.perm.as<int>() : nullptr, d_param, fuse_elastic ? 1 : 0, n, L, c.rank, g, c.eps2, c.coll,
	             p.ev[PH_L2P], nullptr);
	p.leaf_pending = 1;
	p.accn_zero = nearbits != nullptr;
	}
	if (peer) NBCO_TRY(peer_barrier(ctx)); // nobody reads this rank's centres / multipoles / positions any more
	eval_check_kernel<<<1, 32, 0, st>>>(a.cnt, p.cap_list); LAUNCHED(ctx);
	NBCO_CUDA(cudaEventRecord(p.ev[PH_COUNT], st));
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int fmm3_kd_launch(nbco_ctx *ctx, float *d_pos, float *d_acc, int64_t n, const float *d_param, bool fuse_elastic)
{
	if (ctx->cfg.order > NBCO_MAX_ORDER) { set_error("fmm3_kd: order %d outside 1..%d", ctx->cfg.order, NBCO_MAX_ORDER); return NBCO_ERR_INVALID; }
	if (ctx->cfg.world != 1)
	{
		// sharded evaluation: the tree is replicated, every rank computes the accelerations of its own subtree
		if (ctx->cfg.world & (ctx->cfg.world - 1)) { set_error("fmm3_kd: world size must be a power of two"); return NBCO_ERR_INVALID; }
		if (ctx->cfg.unsort) { set_error("fmm3_kd: sharded evaluation needs unsort = 0 (shards are ranges of the tree order)"); return NBCO_ERR_INVALID; }
	}
	NBCO_TRY(ensure_plan(ctx, n));
	FmmPlan &p = *ctx->fmm;
	if ((1 << p.L) < ctx->cfg.world) { set_error("fmm3_kd: more ranks than leaves"); return NBCO_ERR_INVALID; }
	if (p.L > 26) { set_error("fmm3_kd: depth %d exceeds the list encoding", p.L); return NBCO_ERR_INVALID; }
	const bool rebuild = ctx->cfg.unsort || (p.counter % ctx->cfg.tree_steps == 0);
	const OrderOps *ops = order_ops(p.order);
	if (!ops) { set_error("order %d not instantiated", p.order); return NBCO_ERR_INVALID; }
	// Host synchronisation policy.  A rebuild evaluation is synchronised: its counters size the lists (grown with 2x
	// headroom, the evaluation is then repeated) and all pending phase timers are read.  The evaluations that reuse
	// the partition are only enqueued -- their lists differ from the rebuild evaluation's by a few per cent -- and a
	// sticky device flag reports a list that did not fit at the next synchronisation point (fmm3_harvest; every public
	// entry point ends with one).  NBCO_SYNC_EVERY_EVAL=1 restores one synchronisation per evaluation.
	static int sync_all = -1;
	if (sync_all < 0) { const char *e = getenv("NBCO_SYNC_EVERY_EVAL"); sync_all = (e && atoi(e)) ? 1 : 0; }
	const bool must_sync = rebuild || sync_all || p.ev_npend >= FmmPlan::kEvRing - 1 || getenv("NBCO_DEBUG_TRAV");
	bool do_build = rebuild;
	for (int attempt = 0; attempt < 12; ++attempt)
	{
		p.ev = p.evr[p.ev_head];
		p.ev_rebuild[p.ev_head] = do_build;
		int s = run_phases(*ops, ctx, p, d_pos, d_acc, d_param, fuse_elastic, do_build);
		NBCO_TRY(s);
		p.ev_head = (p.ev_head + 1) % FmmPlan::kEvRing; ++p.ev_npend;
		p.rebuilt = rebuild;
		if (!must_sync) { ++p.counter; return NBCO_OK; }
		int over = 0;
		NBCO_TRY(fmm3_harvest(ctx, &over));
		if (getenv("NBCO_DEBUG_TRAV"))
		{
			u32 dbg[20];
			cudaMemcpy(dbg, p.cnt.p, sizeof(dbg), cudaMemcpyDeviceToHost);
			fprintf(stderr, "trav: p2p %u m2l %u records %u changed %u seeds %u retire-rounds %u retired %u rec-rounds %u\n", dbg[0], dbg[1], dbg[11],
			        dbg[12], dbg[16], dbg[13], dbg[14], dbg[15]);
		}
		// headroom for the enqueued evaluations that follow: lists at most half full after a rebuild
		// ... and the records of this from-the-root traversal fill at most a third of their array (traverse_reuse_init_kernel)
		const bool roomy = 2 * p.p2p_n <= (int64_t)p.cap_list && 2 * p.m2l_n <= (int64_t)p.cap_list
		                   && (!p.rec_valid || 3 * p.rec_n <= (int64_t)p.cap_front);
		if (!over && (roomy || !rebuild || ctx->peer.active))
		{
			++p.counter;
			return NBCO_OK;
		}
		if (ctx->peer.active)
		{
			// the ranks pass the peer barriers in lockstep: an evaluation cannot be repeated by one rank alone
			set_error("peer mode: interaction lists exceed capacity (%lld p2p, %lld m2l of %u)", (long long)p.p2p_n, (long long)p.m2l_n, p.cap_list);
			return NBCO_ERR_OVERFLOW;
		}
		// a list or a frontier did not fit (or is more than half full): grow and redo this evaluation.  With
		// unsort == 0 the caller's arrays are already in tree order and the tree is valid: do not build again.
		if (p.cap_list >= 0x7fffffffu / 2) { if (!over) { ++p.counter; return NBCO_OK; } break; }
		{
			// at least double; jump straight to what the counters of this attempt ask for (they are lower bounds when a
			// frontier did not fit: the traversal stopped early), so that a strict MAC does not need many repeats
			int64_t want = 2 * (int64_t)p.cap_list;
			want = std::max(want, 2 * std::max(p.p2p_n, p.m2l_n) + 1024);
			if (p.rec_valid) want = std::max(want, 3 * p.rec_n + 1024);
			p.cap_list = (u32)std::min<int64_t>(want, 0x7fffffff);
			p.cap_front = p.cap_list;
		}
		NBCO_TRY(p.p2p.reserve(8 * (size_t)p.cap_list)); NBCO_TRY(p.m2l.reserve(8 * (size_t)p.cap_list));
		NBCO_TRY(p.frontA.reserve(8 * (size_t)p.cap_front)); NBCO_TRY(p.frontB.reserve(8 * (size_t)p.cap_front));
		p.cap_src = (u32)std::min<int64_t>(4ll * p.cap_list, 0x7fffffff);
		if (p.csr_ready)
		{
			NBCO_TRY(p.csr_src.reserve(4 * (size_t)p.cap_src));
			// an evaluation that overflowed may have left row counters behind: start the rows from zero again
			NBCO_CUDA(cudaMemsetAsync(p.csr_deg.p, 0, 4 * ((size_t)p.ntot + ((size_t)1 << p.L) + 1), ctx->stream));
		}
		p.queue_clean = false; p.rec_valid = false;
		if (!ctx->cfg.unsort) do_build = false;
	}
	set_error("interaction lists exceed capacity (%lld p2p, %lld m2l)", (long long)p.p2p_n, (long long)p.m2l_n);
	return NBCO_ERR_OVERFLOW;
}

// Synchronisation point: wait for the stream, read the counters of the last evaluation and the sticky flags, collect
// the phase timers of every evaluation enqueued since the previous call.  *overflow (optional) = the LAST evaluation's
// lists did not fit; an overflow of an earlier, already accepted evaluation is an error (its results were used).
int fmm3_harvest(nbco_ctx *ctx, int *overflow)
{
	if (overflow) *overflow = 0;
	if (!ctx->fmm) return NBCO_OK;
	FmmPlan &p = *ctx->fmm;
	if (p.ev_npend == 0 || !p.cnt.p) return NBCO_OK;
	u32 h[25] = {};
	NBCO_CUDA(cudaMemcpyAsync(h, p.cnt.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	unsigned perr = 0;
	if (ctx->peer.active) NBCO_TRY(peer_report_error(ctx, &perr));
	if (ctx->peer.active && getenv("NBCO_DEBUG_KD"))
	{
		u32 dw[48];
		cudaMemcpy(dw, (char *)ctx->peer.pub.p + 768, sizeof(dw), cudaMemcpyDeviceToHost);
		fprintf(stderr, "[kd sanity rank %d] select %u %u %u %u %u %u %u | exchange %u %u %u : %u %u %u %u | split", ctx->cfg.rank, dw[0], dw[1], dw[2], dw[3],
		        dw[4], dw[5], dw[6], dw[8], dw[9], dw[10], dw[11], dw[12], dw[13], dw[14]);
		for (int k = 20; k < 32; ++k) fprintf(stderr, " %u", dw[k]);
		fprintf(stderr, " | root %08x %08x %08x %08x %08x %08x | gather %u %u %u\n", dw[32], dw[33], dw[34], dw[35], dw[36], dw[37], dw[40], dw[41], dw[42]);
	}
	const int npend = p.ev_npend;
	for (int k = 0; k < npend; ++k)
	{
		const int set = (p.ev_head - npend + k + 2 * FmmPlan::kEvRing) % FmmPlan::kEvRing;
		for (int q = 0; q < PH_COUNT; ++q)
		{
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, p.evr[set][q], p.evr[set][q + 1]) == cudaSuccess) { p.tot_ms[q] += ms; p.last_ms[q] = ms; }
		}
		++p.tot_evals; p.tot_rebuilds += p.ev_rebuild[set] ? 1 : 0;
	}
	p.ev_npend = 0;
	p.ev_valid = true;
	p.p2p_n = h[0]; p.m2l_n = h[1]; p.rec_n = h[11]; p.rec_fallbacks = h[21];
	if (perr) { set_error("peer barrier timed out: a rank did not arrive (results since the last synchronisation are invalid)"); return NBCO_ERR_CUDA; }
	const bool last_over = h[0] > p.cap_list || h[1] > p.cap_list || h[5] != 0;
	if (h[24])
	{
		NBCO_CUDA(cudaMemsetAsync((u32 *)p.cnt.p + 24, 0, 4, ctx->stream));
		if (npend > 1 || !overflow)
		{
			set_error("interaction lists overflowed during an enqueued tree-reuse evaluation (%u p2p, %u m2l of %u): results since the last "
			          "synchronisation are invalid; rerun with NBCO_SYNC_EVERY_EVAL=1", h[0], h[1], p.cap_list);
			return NBCO_ERR_OVERFLOW;
		}
	}
	if (overflow) *overflow = last_over ? 1 : 0;
	return NBCO_OK;
}

extern const OrderOps kOrderOps1, kOrderOps2, kOrderOps3, kOrderOps4, kOrderOps5, kOrderOps6, kOrderOps7, kOrderOps8, kOrderOps9, kOrderOps10;

const OrderOps *order_ops(int order)
{
	switch (order)
	{
		case 1: return &kOrderOps1;
		case 2: return &kOrderOps2;
		case 3: return &kOrderOps3;
		case 4: return &kOrderOps4;
		case 5: return &kOrderOps5;
		case 6: return &kOrderOps6;
		case 7: return &kOrderOps7;
		case 8: return &kOrderOps8;
		case 9: return &kOrderOps9;
		case 10: return &kOrderOps10;
		default: return nullptr;
	}
}

void fmm3_destroy(nbco_ctx *ctx)
{
	if (!ctx->fmm) return;
	FmmPlan &p = *ctx->fmm;
	kd_release(p.kd);
	DevBuf *all[] = {&p.center, &p.mpole, &p.local, &p.tmp3, &p.accn, &p.nearbits, &p.p2p, &p.m2l, &p.frontA, &p.frontB, &p.cnt, &p.mfac,
	                 &p.csr_deg, &p.csr_off, &p.csr_src, &p.csr_bsum};
	for (DevBuf *b : all) b->release();
	if (p.ev_ok)
		for (int k = 0; k < FmmPlan::kEvRing; ++k)
			for (int i = 0; i <= PH_COUNT; ++i) cudaEventDestroy(p.evr[k][i]);
	delete ctx->fmm;
	ctx->fmm = nullptr;
}

} // namespace nbco

using namespace nbco;

extern "C" {

int nbco_fmm_get_info(nbco_ctx *ctx, nbco_fmm_info *info)
{
	if (!ctx || !info || !ctx->fmm) { set_error("no FMM evaluation yet"); return NBCO_ERR_INVALID; }
	FmmPlan &p = *ctx->fmm;
	info->levels = p.L; info->order = p.order; info->n = p.n; info->nodes = p.ntot;
	info->p2p_pairs = p.p2p_n; info->m2l_pairs = p.m2l_n; info->off_m = p.offM; info->off_l = p.offL;
	info->rebuilt = p.rebuilt; info->mlt_max = p.mlt_max; info->kernel_launches = ctx->launches;
	info->counter = p.counter;
	return NBCO_OK;
}

int nbco_fmm_get_tree(nbco_ctx *ctx, float *h_center, float *h_lbound, float *h_rbound, float *h_mpole, float *h_local,
                      int32_t *h_mult, int32_t *h_index, int32_t *h_splitdim, int32_t *h_perm)
{
	if (!ctx || !ctx->fmm) { set_error("no FMM evaluation yet"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	FmmPlan &p = *ctx->fmm;
	const size_t nt = (size_t)p.ntot;
	if (h_local && p.leaf_pending)
	{
		const OrderOps *ops = order_ops(p.order);
		int g = 0; while ((1 << g) < ctx->cfg.world) ++g;
		TreeData t{p.center.as<float4>(), p.kd.size2.as<float>(), p.mpole.as<float>(), p.local.as<float>(), p.sM, p.sL, {}};
		if (ops && ops->finish_leaf_locals) ops->finish_leaf_locals(ctx, t, p.n, p.L, ctx->cfg.world > 1 ? ctx->cfg.rank : 0, ctx->cfg.world > 1 ? g : 0);
		p.leaf_pending = 0;
	}
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (h_lbound) NBCO_CUDA(cudaMemcpy(h_lbound, p.kd.lbound.p, 12 * nt, cudaMemcpyDeviceToHost));
	if (h_rbound) NBCO_CUDA(cudaMemcpy(h_rbound, p.kd.rbound.p, 12 * nt, cudaMemcpyDeviceToHost));
	if (h_splitdim) NBCO_CUDA(cudaMemcpy(h_splitdim, p.kd.splitdim.p, 4 * nt, cudaMemcpyDeviceToHost));
	if (h_perm) NBCO_CUDA(cudaMemcpy(h_perm, p.kd.perm.p, 4 * (size_t)p.n, cudaMemcpyDeviceToHost));
	if (h_center)
	{
		std::vector<float> tmp(4 * nt);
		NBCO_CUDA(cudaMemcpy(tmp.data(), p.center.p, 16 * nt, cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < nt; ++i) { h_center[3*i] = tmp[4*i]; h_center[3*i+1] = tmp[4*i+1]; h_center[3*i+2] = tmp[4*i+2]; }
	}
	if (h_mpole)
	{
		std::vector<float> tmp(nt * p.sM);
		NBCO_CUDA(cudaMemcpy(tmp.data(), p.mpole.p, 4 * nt * p.sM, cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < nt; ++i) memcpy(h_mpole + i * p.offM, tmp.data() + i * p.sM, 4 * (size_t)p.offM);
	}
	if (h_local)
	{
		std::vector<float> tmp(nt * p.sL);
		NBCO_CUDA(cudaMemcpy(tmp.data(), p.local.p, 4 * nt * p.sL, cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < nt; ++i) memcpy(h_local + i * p.offL, tmp.data() + i * p.sL, 4 * (size_t)p.offL);
	}
	if (h_mult || h_index)
		for (int l = 0; l <= p.L; ++l)
			for (int64_t i = 0; i < (1ll << l); ++i)
			{
				int64_t s0 = seg_start(p.n, i, l), s1 = seg_start(p.n, i + 1, l);
				if (h_index) h_index[kd_beg(l) + i] = (int32_t)s0;
				if (h_mult) h_mult[kd_beg(l) + i] = (int32_t)(s1 - s0);
			}
	return NBCO_OK;
}

int nbco_fmm_get_lists(nbco_ctx *ctx, int32_t *h_p2p, int64_t p2p_cap, int32_t *h_m2l, int64_t m2l_cap)
{
	if (!ctx || !ctx->fmm) { set_error("no FMM evaluation yet"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	FmmPlan &p = *ctx->fmm;
	if (p2p_cap < p.p2p_n || m2l_cap < p.m2l_n) { set_error("list buffers too small"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	auto fetch = [&](int32_t *dst, const DevBuf &src, int64_t cnt) -> int
	{
		if (!dst || cnt == 0) return NBCO_OK;
		NBCO_CUDA(cudaMemcpy(dst, src.p, 8 * (size_t)cnt, cudaMemcpyDeviceToHost));
		int2 *q = reinterpret_cast<int2 *>(dst);
		for (int64_t i = 0; i < cnt; ++i) q[i].x &= kNodeMask;
		std::sort(q, q + cnt, [](const int2 &a, const int2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
		return NBCO_OK;
	};
	NBCO_TRY(fetch(h_p2p, p.p2p, p.p2p_n));
	NBCO_TRY(fetch(h_m2l, p.m2l, p.m2l_n));
	return NBCO_OK;
}

int nbco_fmm_get_phase_ms(nbco_ctx *ctx, const char **names, float *ms, int cap)
{
	if (!ctx || !ctx->fmm) return 0;
	cudaSetDevice(ctx->cfg.device);
	if (fmm3_harvest(ctx, nullptr) != NBCO_OK || !ctx->fmm->ev_valid) return 0;
	FmmPlan &p = *ctx->fmm;
	int k = 0;
	for (; k < PH_COUNT && k < cap; ++k) { names[k] = kPhaseNames[k]; ms[k] = p.last_ms[k]; }
	return k;
}

int nbco_fmm_phase_totals(nbco_ctx *ctx, const char **names, double *ms, int cap, int64_t *h_evals, int reset)
{
	if (!ctx || !ctx->fmm) return 0;
	cudaSetDevice(ctx->cfg.device);
	if (fmm3_harvest(ctx, nullptr) != NBCO_OK) return 0;
	FmmPlan &p = *ctx->fmm;
	int k = 0;
	for (; k < PH_COUNT && k < cap; ++k) { names[k] = kPhaseNames[k]; ms[k] = p.tot_ms[k]; }
	if (h_evals) { h_evals[0] = p.tot_evals; h_evals[1] = p.tot_rebuilds; }
	if (reset) { for (int i = 0; i < PH_COUNT; ++i) p.tot_ms[i] = 0; p.tot_evals = p.tot_rebuilds = 0; }
	return k;
}

} // extern "C"
