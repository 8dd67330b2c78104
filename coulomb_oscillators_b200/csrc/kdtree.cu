// kdtree.cu -- balanced kd-tree build for sm_100a.
//
// Replaces the tree build of fmm_cart3_kdtree (reference Simulation/fmm_cart3_kdtree.cuh:1619-1642):
// minmaxReduce2 (reductions.cuh:67-80), evalRootBox/evalBox (:89-156), evalKeys_kdtree (:158-202) and,
// per level, a full key-value sort (CUB radix at level 0, bb_segsort below, :1311-1364) followed by
// four gather/copy passes over positions and the permutation.
//
// What the tree needs: at every level each node's particles are split at a fixed
// rank along the node's widest axis, the children's boxes are cut at the two boundary particles, and
// only the last level's order is observable.  So every level is a median SELECTION, not a sort:
//   * order = the order a stable sort at every level would give; equivalently a TOTAL order per node:
//     (coordinate on the split axis, coordinates on the previously used distinct axes, most recent
//     first, input index).  With a total order any selection algorithm yields the same partition.
//   * selection by a LINEAR histogram (round 2).  A node's box bounds its particles on the split axis, so
//     bin = floor((x - lo) * nbins / (hi - lo)) is monotone in x and spreads the node's particles over all
//     bins (the radix digits of round 1 put a whole segment into a handful of bins and needed 3-4 passes).
//     ONE histogram pass finds the bin that holds the median rank; the particles of the other bins are on a
//     known side.  Only the "candidates" inside the pivot bin (~1/nbins of the segment times the local density
//     contrast) are ordered exactly, by a lexicographic radix select over the words of the total order.
//   * levels whose segments exceed kBottomCap particles ("top"): the particles travel as 16-byte records
//     (x, y, z, id) between two buffers; per level one histogram pass (16 B read / particle), one three-way
//     partition pass (16 B read + 16 B write, coalesced runs per warp) and one CTA per segment that resolves
//     the candidates and writes the children's boxes.
//   * all remaining levels run in ONE kernel, a CTA per level-lt node with the coordinates resident in
//     shared memory and a u16 slot permutation that is partitioned level by level (same histogram
//     selection; blocks of <= 32 slots and the last level are ranked by counting).
//   * particle data moves once, at the end (sorted positions + permutation).

#include "fmm3_common.cuh"

namespace nbco {

namespace {

constexpr int kBins = 2048;         // top levels: linear bins over the node's extent on its split axis
constexpr int kTopThreads = 512;
constexpr int kRows = 8;            // rows of 32 records a warp classifies per reservation
constexpr int kChunk = kTopThreads * kRows;
constexpr int kResThreads = 1024;   // candidate resolution: one CTA per segment

struct SegState { u32 pb, less, eq, curL, curE, curR, rmin, pad; }; // selection state of one segment

__device__ __forceinline__ float axis_of(const float4 &p, int axis) { return axis == 0 ? p.x : (axis == 1 ? p.y : p.z); }

// scale of the linear bins of a node: nbins / extent, 0 for a degenerate extent (everything in bin 0)
__device__ __forceinline__ float bin_scale(float lo, float hi, int nbins)
{
	return hi > lo ? __fdiv_rn((float)nbins, __fsub_rn(hi, lo)) : 0.f;
}
// monotone in x for fixed (lo, scale): subtraction, multiplication by a non-negative constant and truncation
// all preserve <=; explicit rounding intrinsics keep the histogram and the partition kernels bit-identical
__device__ __forceinline__ int bin_of(float x, float lo, float scale, int nbins)
{
	const int b = (int)__fmul_rn(__fsub_rn(x, lo), scale); // NaN -> 0, +inf -> INT_MAX
	return min(max(b, 0), nbins - 1);
}

// =====================================================================================
//  pack: AoS float3 positions -> (x, y, z, id) records, bounding box in the same pass
// =====================================================================================
__global__ void __launch_bounds__(256) pack_bbox_kernel(const float *__restrict__ pos, float4 *__restrict__ pay, int64_t n, u32 *__restrict__ out6)
{
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
	{
		const float x = pos[3*i], y = pos[3*i+1], z = pos[3*i+2];
		pay[i] = make_float4(x, y, z, __uint_as_float((u32)i));
		mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
		mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
		mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
	}
#pragma unroll
	for (int k = 0; k < 3; ++k)
		for (int o = 16; o > 0; o >>= 1)
		{
			mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
			mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
		}
	if ((threadIdx.x & 31) == 0)
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			atomicMin(out6 + k, ordered_bits(mn[k]));
			atomicMax(out6 + 3 + k, ordered_bits(mx[k]));
		}
}

__global__ void root_box_kernel(TreeGeom g, const u32 *__restrict__ bb)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		float lb[3], rb[3];
		for (int k = 0; k < 3; ++k) { lb[k] = unordered_bits(bb[k]); rb[k] = unordered_bits(bb[3+k]); }
		write_box(g, 0, lb, rb, kNoAxis | (kNoAxis << 2) | (kNoAxis << 4));
	}
}

// =====================================================================================
//  top levels: histogram select + three-way partition per segment
// =====================================================================================
struct TileRange { int64_t a, b, s0; int seg; };

__device__ __forceinline__ TileRange tile_range(int64_t n, int l, int tps, int tile, int seg0)
{
	TileRange r;
	r.seg = seg0 + blockIdx.x / tps;
	const int t = blockIdx.x % tps;
	r.s0 = seg_start(n, r.seg, l);
	const int64_t s1 = seg_start(n, r.seg + 1, l);
	r.a = r.s0 + (int64_t)t * tile;
	r.b = r.a + tile < s1 ? r.a + tile : s1;
	return r;
}

// histogram of the linear bins of every segment of level l
__global__ void __launch_bounds__(kTopThreads)
top_hist_kernel(const float4 *__restrict__ pay, TreeGeom g, u32 *__restrict__ hist, int64_t n, int l, int tps, int tile, int seg0)
{
	__shared__ u32 sh[kBins];
	for (int b = threadIdx.x; b < kBins; b += kTopThreads) sh[b] = 0;
	__syncthreads();
	const TileRange r = tile_range(n, l, tps, tile, seg0);
	const int node = kd_beg(l) + r.seg, axis = g.splitdim[node];
	const float lo = g.lbound[3*node + axis], scale = bin_scale(lo, g.rbound[3*node + axis], kBins);
	int64_t j = r.a + threadIdx.x;
	for (; j + 3 * kTopThreads < r.b; j += 4 * kTopThreads)
	{
		const float4 p0 = pay[j], p1 = pay[j + kTopThreads], p2 = pay[j + 2 * kTopThreads], p3 = pay[j + 3 * kTopThreads];
		atomicAdd(&sh[bin_of(axis_of(p0, axis), lo, scale, kBins)], 1u);
		atomicAdd(&sh[bin_of(axis_of(p1, axis), lo, scale, kBins)], 1u);
		atomicAdd(&sh[bin_of(axis_of(p2, axis), lo, scale, kBins)], 1u);
		atomicAdd(&sh[bin_of(axis_of(p3, axis), lo, scale, kBins)], 1u);
	}
	for (; j < r.b; j += kTopThreads)
		atomicAdd(&sh[bin_of(axis_of(pay[j], axis), lo, scale, kBins)], 1u);
	__syncthreads();
	if (r.a < r.b)
	{
		u32 *gh = hist + (int64_t)(r.seg - seg0) * kBins;
		for (int b = threadIdx.x; b < kBins; b += kTopThreads)
		{
			const u32 c = sh[b];
			if (c) atomicAdd(gh + b, c);
		}
	}
}

// one CTA per segment: the bin that holds the median rank; clears the histogram for the next level
__global__ void __launch_bounds__(256)
top_pick_kernel(SegState *__restrict__ st, u32 *__restrict__ hist, int64_t n, int l, int seg0)
{
	constexpr int kPer = kBins / 256;
	__shared__ u32 wsum[8];
	const int seg = seg0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	u32 *h = hist + (int64_t)blockIdx.x * kBins;
	// pivot = last particle of the left child: rank (left count - 1) in the segment
	const u32 krank = (u32)(seg_start(n, 2 * (int64_t)seg + 1, l + 1) - seg_start(n, 2 * (int64_t)seg, l + 1)) - 1u;
	u32 c[kPer], s = 0;
#pragma unroll
	for (int k = 0; k < kPer; ++k) { c[k] = h[tid * kPer + k]; h[tid * kPer + k] = 0; s += c[k]; }
	u32 incl = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	u32 wbase = 0;
	for (int i = 0; i < w; ++i) wbase += wsum[i];
	const u32 excl = wbase + incl - s;
	if (krank >= excl && krank < excl + s)
	{
		u32 run = excl;
#pragma unroll
		for (int k = 0; k < kPer; ++k)
		{
			if (krank >= run && krank < run + c[k])
			{
				SegState z;
				z.pb = (u32)(tid * kPer + k); z.less = run; z.eq = c[k];
				z.curL = z.curE = z.curR = 0; z.rmin = 0xffffffffu; z.pad = 0;
				st[blockIdx.x] = z;
			}
			run += c[k];
		}
	}
}

// three-way split of every segment by bin: [bins < pb | bin == pb (candidates) | bins > pb], unordered inside
__global__ void __launch_bounds__(kTopThreads)
top_partition_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, TreeGeom g, SegState *__restrict__ st,
                     int64_t n, int l, int tps, int tile, int seg0)
{
	constexpr int kWarps = kTopThreads / 32;
	__shared__ u32 wcnt[kWarps][3];
	__shared__ u32 base[3];
	const TileRange r = tile_range(n, l, tps, tile, seg0);
	if (r.a >= r.b) return;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const u32 lt_mask = (1u << lane) - 1u;
	SegState *sp = st + (r.seg - seg0);
	const u32 pb = sp->pb, less = sp->less, eq = sp->eq;
	const int node = kd_beg(l) + r.seg, axis = g.splitdim[node];
	const float lo = g.lbound[3*node + axis], scale = bin_scale(lo, g.rbound[3*node + axis], kBins);
	u32 rmin = 0xffffffffu;
	for (int64_t c0 = r.a; c0 < r.b; c0 += kChunk)
	{
		float4 v[kRows];
		u32 cls = 0; // 2 bits per row: 0 left, 1 candidate, 2 right, 3 no record
		u32 nl = 0, ne = 0, nr = 0;
#pragma unroll
		for (int it = 0; it < kRows; ++it)
		{
			const int64_t j = c0 + ((int64_t)w * kRows + it) * 32 + lane;
			const bool valid = j < r.b;
			v[it] = valid ? in[j] : make_float4(0.f, 0.f, 0.f, 0.f);
		}
#pragma unroll
		for (int it = 0; it < kRows; ++it)
		{
			const int64_t j = c0 + ((int64_t)w * kRows + it) * 32 + lane;
			const bool valid = j < r.b;
			const float x = axis_of(v[it], axis);
			const u32 b = (u32)bin_of(x, lo, scale, kBins);
			const u32 c = !valid ? 3u : (b < pb ? 0u : (b == pb ? 1u : 2u));
			cls |= c << (2 * it);
			nl += __popc(__ballot_sync(0xffffffffu, c == 0u));
			ne += __popc(__ballot_sync(0xffffffffu, c == 1u));
			nr += __popc(__ballot_sync(0xffffffffu, c == 2u));
			if (c == 2u) rmin = min(rmin, ordered_bits(x));
		}
		if (lane == 0) { wcnt[w][0] = nl; wcnt[w][1] = ne; wcnt[w][2] = nr; }
		__syncthreads();
		if (tid < 3)
		{
			u32 tot = 0;
			for (int i = 0; i < kWarps; ++i) tot += wcnt[i][tid];
			u32 *c = tid == 0 ? &sp->curL : (tid == 1 ? &sp->curE : &sp->curR);
			base[tid] = tot ? atomicAdd(c, tot) : 0u;
		}
		__syncthreads();
		u32 ol = base[0], oe = less + base[1], orr = less + eq + base[2];
		for (int i = 0; i < w; ++i) { ol += wcnt[i][0]; oe += wcnt[i][1]; orr += wcnt[i][2]; }
#pragma unroll
		for (int it = 0; it < kRows; ++it)
		{
			const u32 c = (cls >> (2 * it)) & 3u;
			const u32 bl = __ballot_sync(0xffffffffu, c == 0u), be = __ballot_sync(0xffffffffu, c == 1u), br = __ballot_sync(0xffffffffu, c == 2u);
			if (c == 0u) out[r.s0 + ol + __popc(bl & lt_mask)] = v[it];
			else if (c == 1u) out[r.s0 + oe + __popc(be & lt_mask)] = v[it];
			else if (c == 2u) out[r.s0 + orr + __popc(br & lt_mask)] = v[it];
			ol += __popc(bl); oe += __popc(be); orr += __popc(br);
		}
		__syncthreads(); // wcnt / base are reused by the next chunk
	}
	rmin = __reduce_min_sync(0xffffffffu, rmin);
	if (lane == 0 && rmin != 0xffffffffu) atomicMin(&sp->rmin, rmin);
}

// words of the total order of a node's particles: key on the split axis, keys on the previously used distinct
// axes (most recent first; absent axes contribute a constant), input index
struct Words { u32 w[4]; };
__device__ __forceinline__ Words words_of(const float4 &p, int chain)
{
	Words r;
	const int a1 = (chain >> 2) & 3, a2 = (chain >> 4) & 3;
	r.w[0] = ordered_bits(axis_of(p, chain & 3));
	r.w[1] = a1 == kNoAxis ? 0u : ordered_bits(axis_of(p, a1));
	r.w[2] = a2 == kNoAxis ? 0u : ordered_bits(axis_of(p, a2));
	r.w[3] = __float_as_uint(p.w);
	return r;
}

// block-wide search of the bin that holds rank r in a 2048-bin shared histogram (1024 threads, 2 bins each)
__device__ __forceinline__ void find_rank_bin(const u32 *sh, u32 r, u32 *wsum /* 32 */, u32 *found /* 3: bin, below, count */)
{
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const u32 c0 = sh[2 * tid], c1 = sh[2 * tid + 1], s = c0 + c1;
	u32 incl = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	u32 wbase = 0;
	for (int i = 0; i < w; ++i) wbase += wsum[i];
	const u32 excl = wbase + incl - s;
	if (r >= excl && r < excl + s)
	{
		if (r < excl + c0) { found[0] = 2 * tid; found[1] = excl; found[2] = c0; }
		else { found[0] = 2 * tid + 1; found[1] = excl + c0; found[2] = c1; }
	}
	__syncthreads();
}

// One CTA per segment: order the candidates of the pivot bin exactly.  Lexicographic radix select (11 + 11 + 10
// bits per word, at most four words) of the candidate of rank need - 1; the candidates are then split around it
// through the (dead) input buffer and the children's boxes are written (evalBox_krnl, :109-137).
__global__ void __launch_bounds__(kResThreads)
top_resolve_kernel(float4 *__restrict__ out, float4 *__restrict__ scratch, TreeGeom g, const SegState *__restrict__ st,
                   int64_t n, int l, int seg0)
{
	__shared__ u32 sh[kBins];
	__shared__ u32 wsum[32];
	__shared__ u32 found[3];
	__shared__ u32 run[2];
	__shared__ u32 s_minr;
	const int seg = seg0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const SegState s = st[blockIdx.x];
	const int node = kd_beg(l) + seg, axis = g.splitdim[node], chain = g.chain[node];
	const int64_t s0 = seg_start(n, seg, l);
	const u32 kleft = (u32)(seg_start(n, 2 * (int64_t)seg + 1, l + 1) - seg_start(n, 2 * (int64_t)seg, l + 1));
	const u32 eq = s.eq, need = kleft - s.less; // 1 <= need <= eq
	float4 *C = out + s0 + s.less, *T = scratch + s0 + s.less;

	// ---- select: pivot words piv[0 .. depth) ----
	u32 piv[4] = {0u, 0u, 0u, 0u};
	int depth = 0;
	u32 r = need - 1, t = eq;
	for (int d = 0; d < 4; ++d)
	{
		if ((d == 1 || d == 2) && ((chain >> (2 * d)) & 3) == kNoAxis) { depth = d + 1; continue; } // constant word
		u32 prefix = 0;
		for (int pass = 0; pass < 3; ++pass)
		{
			const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
			const u32 mask = pass == 2 ? 1023u : 2047u;
			for (int b = tid; b < kBins; b += kResThreads) sh[b] = 0;
			__syncthreads();
			// four independent loads per thread and trip: one CTA walks up to ~64 k candidates (root level), latency-bound
			for (u32 i0 = tid; i0 < eq; i0 += 4 * kResThreads)
			{
				float4 q[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) { const u32 i = i0 + k * kResThreads; q[k] = i < eq ? C[i] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
				for (int k = 0; k < 4; ++k)
				{
					if (i0 + k * kResThreads >= eq) continue;
					const Words W = words_of(q[k], chain);
					bool m = true;
					for (int e = 0; e < d; ++e) m = m && W.w[e] == piv[e];
					if (pass > 0) m = m && (W.w[d] >> (shift + (pass == 1 ? 11 : 10))) == (prefix >> (shift + (pass == 1 ? 11 : 10)));
					if (m) atomicAdd(&sh[(W.w[d] >> shift) & mask], 1u);
				}
			}
			__syncthreads();
			find_rank_bin(sh, r, wsum, found);
			prefix |= found[0] << shift;
			r -= found[1];
			t = found[2];
			__syncthreads();
		}
		piv[d] = prefix;
		depth = d + 1;
		if (t == 1u || r == t - 1u) break; // the pivot is unique, or every particle that ties with it goes left
	}

	// ---- split the candidates: lexicographic (w[0 .. depth)) <= piv goes left ----
	// four rows of kResThreads candidates per trip (independent loads, one pair of barriers per 4096 candidates)
	__shared__ u32 rowcnt[4][32];
	if (tid == 0) { run[0] = 0; run[1] = 0; s_minr = 0xffffffffu; }
	__syncthreads();
	u32 minr = 0xffffffffu;
	for (u32 base = 0; base < eq; base += 4 * kResThreads)
	{
		float4 q[4];
		bool valid[4], left[4];
		u32 bl[4], br[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			const u32 i = base + k * kResThreads + tid;
			valid[k] = i < eq;
			q[k] = valid[k] ? C[i] : make_float4(0.f, 0.f, 0.f, 0.f);
		}
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			left[k] = false;
			if (valid[k])
			{
				const Words W = words_of(q[k], chain);
				left[k] = true; // equal on every compared word: the pivot itself or a tie that goes left
				for (int e = 0; e < depth; ++e)
					if (W.w[e] != piv[e]) { left[k] = W.w[e] < piv[e]; break; }
				if (!left[k]) minr = min(minr, W.w[0]);
			}
			bl[k] = __ballot_sync(0xffffffffu, valid[k] && left[k]);
			br[k] = __ballot_sync(0xffffffffu, valid[k] && !left[k]);
			if (lane == 0) rowcnt[k][w] = (u32)__popc(bl[k]) | ((u32)__popc(br[k]) << 16);
		}
		__syncthreads();
		u32 ol = run[0], orr = run[1];
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			u32 pl = ol, pr = orr;
			for (int ww = 0; ww < 32; ++ww)
			{
				const u32 c = rowcnt[k][ww];
				if (ww < w) { pl += c & 0xffffu; pr += c >> 16; }
				ol += c & 0xffffu; orr += c >> 16;
			}
			if (valid[k]) T[left[k] ? pl + __popc(bl[k] & ((1u << lane) - 1u)) : need + pr + __popc(br[k] & ((1u << lane) - 1u))] = q[k];
		}
		__syncthreads();
		if (tid == 0) { run[0] = ol; run[1] = orr; }
		__syncthreads();
	}
	minr = __reduce_min_sync(0xffffffffu, minr);
	if (lane == 0 && minr != 0xffffffffu) atomicMin(&s_minr, minr);
	__threadfence_block();
	__syncthreads();
	for (u32 i0 = tid; i0 < eq; i0 += 4 * kResThreads)
	{
		float4 q[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) { const u32 i = i0 + k * kResThreads; if (i < eq) q[k] = T[i]; }
#pragma unroll
		for (int k = 0; k < 4; ++k) { const u32 i = i0 + k * kResThreads; if (i < eq) C[i] = q[k]; }
	}

	// ---- boxes of the children: cut at the last particle of the left and the first of the right child ----
	if (tid == 0)
	{
		float lb[3], rb[3];
		for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
		const int pch = chain;
		const float save = rb[axis];
		rb[axis] = unordered_bits(piv[0]);
		write_box(g, 2*node + 1, lb, rb, pch);
		rb[axis] = save;
		lb[axis] = unordered_bits(min(s_minr, s.rmin));
		write_box(g, 2*node + 2, lb, rb, pch);
	}
}

// =====================================================================================
//  bottom levels: one CTA per level-lt node, particles resident in shared memory
// =====================================================================================
constexpr int kMaxHistBlk = 128;     // histogram mode: blocks of >= 64 slots (at most kBottomCap / 64 of them)
constexpr u16 kNoSlot = 0xffffu;

struct BlkBox { float lb[3], rb[3]; int chain, axis; }; // box of a block's node, its tie-break chain and split axis

struct BottomSmem
{
	float *c;            // [3][kBottomCap] coordinates by slot
	u16 *ordA, *ordB;    // slot permutation, ping-pong; kNoSlot pads every block to its power-of-two size
	u32 *ckey;           // keys (ordered bits) parallel to a candidate list / to the block positions
	u32 *hist;           // [blocks][bins], at most kBottomCap / 8 counters
	BlkBox *box[2];      // histogram levels: boxes of the blocks of this level / of the next one (no global round trip)
	// per-block state of a histogram level
	int *b_kl;
	float *b_lo, *b_scale;
	u32 *b_pb, *b_less, *b_eq, *b_curL, *b_curE, *b_curR, *b_rmin, *b_cutL, *b_cutR;
};

// rest of the total order for two slots with equal split-axis keys
__device__ __forceinline__ bool slot_tie_less(const BottomSmem &s, u32 sa, u32 sb, int chain, const float4 *__restrict__ pay0)
{
	for (int c = 1; c < 3; ++c)
	{
		const int ax = (chain >> (2 * c)) & 3;
		if (ax == kNoAxis) break;
		const u32 ua = ordered_bits(s.c[ax * kBottomCap + sa]), ub = ordered_bits(s.c[ax * kBottomCap + sb]);
		if (ua != ub) return ua < ub;
	}
	return __float_as_uint(pay0[sa].w) < __float_as_uint(pay0[sb].w);
}

// write_box (fmm3_common.cuh) that also returns the node's state for the next level
__device__ __forceinline__ void write_box_keep(const TreeGeom &g, int node, const float *lb, const float *rb, int parent_chain, BlkBox *keep)
{
	write_box(g, node, lb, rb, parent_chain);
	if (keep)
	{
		const int ax = widest_axis(rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2]);
		for (int k = 0; k < 3; ++k) { keep->lb[k] = lb[k]; keep->rb[k] = rb[k]; }
		keep->axis = ax; keep->chain = chain_push(ax, parent_chain);
	}
}

__global__ void __launch_bounds__(kBottomThreads, 1)
kd_bottom_kernel(TreeGeom g, const float4 *__restrict__ pay, float *__restrict__ spos, int *__restrict__ perm,
                 int64_t n, int lt, int L, int P2, int blk0)
{
	extern __shared__ unsigned char smem_raw[];
	BottomSmem s;
	{
		unsigned char *p = smem_raw;
		s.c = reinterpret_cast<float *>(p); p += sizeof(float) * 3 * kBottomCap;
		s.ckey = reinterpret_cast<u32 *>(p); p += sizeof(u32) * kBottomCap;
		s.hist = reinterpret_cast<u32 *>(p); p += sizeof(u32) * (kBottomCap / 8);
		s.ordA = reinterpret_cast<u16 *>(p); p += sizeof(u16) * kBottomCap;
		s.ordB = reinterpret_cast<u16 *>(p); p += sizeof(u16) * kBottomCap;
		s.box[0] = reinterpret_cast<BlkBox *>(p); p += sizeof(BlkBox) * kMaxHistBlk;
		s.box[1] = reinterpret_cast<BlkBox *>(p); p += sizeof(BlkBox) * kMaxHistBlk;
		int *q = reinterpret_cast<int *>(p);
		s.b_kl = q; q += kMaxHistBlk;
		s.b_lo = reinterpret_cast<float *>(q); q += kMaxHistBlk; s.b_scale = reinterpret_cast<float *>(q); q += kMaxHistBlk;
		u32 *u = reinterpret_cast<u32 *>(q);
		s.b_pb = u; u += kMaxHistBlk; s.b_less = u; u += kMaxHistBlk; s.b_eq = u; u += kMaxHistBlk;
		s.b_curL = u; u += kMaxHistBlk; s.b_curE = u; u += kMaxHistBlk; s.b_curR = u; u += kMaxHistBlk;
		s.b_rmin = u; u += kMaxHistBlk; s.b_cutL = u; u += kMaxHistBlk; s.b_cutR = u; u += kMaxHistBlk;
	}
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int b = blk0 + blockIdx.x;
	const int64_t s0 = seg_start(n, b, lt);
	const int c0 = (int)(seg_start(n, b + 1, lt) - s0);
	const float4 *__restrict__ pay0 = pay + s0;
	constexpr int kPer = kBottomCap / kBottomThreads;

	{
		// all loads of a thread are issued before the first store (one memory round trip, not eight)
		float4 v[kPer];
#pragma unroll
		for (int e = 0; e < kPer; ++e) { const int t = tid + e * kBottomThreads; v[e] = t < c0 ? pay0[t] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
		for (int e = 0; e < kPer; ++e)
		{
			const int t = tid + e * kBottomThreads;
			if (t < c0) { s.c[t] = v[e].x; s.c[kBottomCap + t] = v[e].y; s.c[2 * kBottomCap + t] = v[e].z; }
		}
	}
	for (int p = tid; p < P2; p += kBottomThreads) s.ordA[p] = p < c0 ? (u16)p : kNoSlot;
	if (tid == 0)
	{
		const int node = kd_beg(lt) + b;
		BlkBox bx;
		for (int k = 0; k < 3; ++k) { bx.lb[k] = g.lbound[3*node+k]; bx.rb[k] = g.rbound[3*node+k]; }
		bx.axis = g.splitdim[node]; bx.chain = g.chain[node];
		s.box[0][0] = bx;
	}
	__syncthreads();

	u16 *oin = s.ordA, *oout = s.ordB;
	int cb = 0; // s.box[cb] holds the boxes of the current level's blocks while the levels run in histogram mode
	const int nlev = L - lt; // levels lt .. L-1 are split here
	for (int j = 0; j < nlev; ++j)
	{
		const int l = lt + j;
		const int B = P2 >> j, logB = 31 - __clz(B), nblk = 1 << j;
		const bool last = j + 1 == nlev;
		const int node0 = kd_beg(l) + (int)((int64_t)b << j);
		const int64_t i0 = (int64_t)b << j;
		for (int p = tid; p < P2 / 2; p += kBottomThreads) reinterpret_cast<u32 *>(oout)[p] = 0xffffffffu;

		if (!last && B >= 64)
		{
			// ================= histogram mode =================
			const int nb = min(256, B >> 3);
			const BlkBox *bx = s.box[cb];
			BlkBox *bnext = 2 * nblk <= kMaxHistBlk ? s.box[cb ^ 1] : nullptr;
			if (tid < nblk)
			{
				const int64_t i = i0 + tid;
				const int axis = bx[tid].axis;
				s.b_kl[tid] = (int)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
				const float lo = bx[tid].lb[axis];
				s.b_lo[tid] = lo; s.b_scale[tid] = bin_scale(lo, bx[tid].rb[axis], nb);
				s.b_curL[tid] = s.b_curE[tid] = s.b_curR[tid] = 0;
				s.b_rmin[tid] = s.b_cutL[tid] = s.b_cutR[tid] = 0xffffffffu;
			}
			for (int i = tid; i < nblk * nb; i += kBottomThreads) s.hist[i] = 0;
			__syncthreads();
			// (1) histogram; slot and bin stay in registers
			u32 sb[kPer];
#pragma unroll
			for (int e = 0; e < kPer; ++e)
			{
				const int p = tid + e * kBottomThreads;
				sb[e] = 0xffffffffu;
				if (p < P2)
				{
					const u32 slot = oin[p];
					if (slot != kNoSlot)
					{
						const int q = p >> logB;
						const int bin = bin_of(s.c[bx[q].axis * kBottomCap + slot], s.b_lo[q], s.b_scale[q], nb);
						atomicAdd(&s.hist[q * nb + bin], 1u);
						sb[e] = slot | ((u32)bin << 16);
					}
				}
			}
			__syncthreads();
			// (2) pivot bin of every block: a warp scans the (at most 256) bins of a block
			for (int q = warp; q < nblk; q += kBottomThreads / 32)
			{
				const u32 *h = s.hist + q * nb;
				const u32 krank = (u32)s.b_kl[q] - 1u;
				const int per = (nb + 31) >> 5; // bins per lane (1 .. 8), lane-contiguous
				u32 c[8], sum = 0;
#pragma unroll
				for (int k = 0; k < 8; ++k) { const int bi = lane * per + k; c[k] = (k < per && bi < nb) ? h[bi] : 0u; sum += c[k]; }
				u32 incl = sum;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
				u32 run = incl - sum;
				if (krank >= run && krank < run + sum)
				{
#pragma unroll
					for (int k = 0; k < 8; ++k)
					{
						if (k < per && krank >= run && krank < run + c[k]) { s.b_pb[q] = (u32)(lane * per + k); s.b_less[q] = run; s.b_eq[q] = c[k]; }
						run += c[k];
					}
				}
			}
			__syncthreads();
			// (3) three-way split: the 32 positions of a warp lie in one block (B >= 64)
#pragma unroll
			for (int e = 0; e < kPer; ++e)
			{
				const int p = tid + e * kBottomThreads;
				const int q = (p < P2 ? p : P2 - 1) >> logB;
				const bool valid = sb[e] != 0xffffffffu;
				const u32 slot = sb[e] & 0xffffu, bin = sb[e] >> 16;
				const u32 pb = s.b_pb[q], less = s.b_less[q], eq = s.b_eq[q], need = (u32)s.b_kl[q] - less;
				const bool isl = valid && bin < pb, ise = valid && bin == pb, isr = valid && bin > pb;
				const u32 bl = __ballot_sync(0xffffffffu, isl), be = __ballot_sync(0xffffffffu, ise), br = __ballot_sync(0xffffffffu, isr);
				u32 ol = 0, oe = 0, orr = 0;
				if (lane == 0)
				{
					if (bl) ol = atomicAdd(&s.b_curL[q], (u32)__popc(bl));
					if (be) oe = atomicAdd(&s.b_curE[q], (u32)__popc(be));
					if (br) orr = atomicAdd(&s.b_curR[q], (u32)__popc(br));
				}
				ol = __shfl_sync(0xffffffffu, ol, 0); oe = __shfl_sync(0xffffffffu, oe, 0); orr = __shfl_sync(0xffffffffu, orr, 0);
				const u32 lt_mask = (1u << lane) - 1u;
				u32 key = 0xffffffffu;
				if (ise || isr) key = ordered_bits(s.c[bx[q].axis * kBottomCap + slot]);
				if (br)
				{
					const u32 rmin = __reduce_min_sync(0xffffffffu, isr ? key : 0xffffffffu);
					if (lane == 0) atomicMin(&s.b_rmin[q], rmin);
				}
				const int qB = q << logB;
				if (isl) oout[qB + ol + __popc(bl & lt_mask)] = (u16)slot;
				else if (isr) oout[qB + (B >> 1) + (eq - need) + orr + __popc(br & lt_mask)] = (u16)slot;
				else if (ise)
				{
					// candidate list of the block: in the input array (every reader of it is past the barrier above)
					const int cp = qB + less + oe + __popc(be & lt_mask);
					oin[cp] = (u16)slot; s.ckey[cp] = key;
				}
			}
			__syncthreads();
			// (4) rank the candidates of every block by counting (a group of warps per block)
			{
				const int wpb = nblk >= 32 ? 1 : 32 / nblk, ngroups = 32 / wpb;
				const int gl = (warp % wpb) * 32 + lane, gstride = wpb * 32;
				for (int q = warp / wpb; q < nblk; q += ngroups)
				{
					const u32 less = s.b_less[q], eq = s.b_eq[q], need = (u32)s.b_kl[q] - less;
					const int cbase = (q << logB) + (int)less, chain = bx[q].chain;
					for (u32 c = gl; c < eq; c += gstride)
					{
						const u32 kc = s.ckey[cbase + c], sc = oin[cbase + c];
						u32 rank = 0, ties = 0;
						for (u32 f = 0; f < eq; ++f)
						{
							const u32 kf = s.ckey[cbase + f];
							rank += kf < kc ? 1u : 0u;
							ties += kf == kc ? 1u : 0u;
						}
						if (ties > 1u) // equal keys (rare): the rest of the total order decides
							for (u32 f = 0; f < eq; ++f)
								if (f != c && s.ckey[cbase + f] == kc && slot_tie_less(s, oin[cbase + f], sc, chain, pay0)) ++rank;
						if (rank < need) oout[cbase + rank] = (u16)sc;
						else oout[(q << logB) + (B >> 1) + (rank - need)] = (u16)sc;
						if (rank == need - 1) s.b_cutL[q] = kc;
						if (rank == need) s.b_cutR[q] = kc;
					}
				}
			}
			__syncthreads();
			// (5) boxes of the children (evalBox_krnl for level l+1); kept in shared memory for the next histogram level
			if (tid < nblk)
			{
				const int node = node0 + tid, axis = bx[tid].axis, pch = bx[tid].chain;
				float lb[3], rb[3];
				for (int k = 0; k < 3; ++k) { lb[k] = bx[tid].lb[k]; rb[k] = bx[tid].rb[k]; }
				const float save = rb[axis];
				rb[axis] = unordered_bits(s.b_cutL[tid]);
				write_box_keep(g, 2*node + 1, lb, rb, pch, bnext ? bnext + 2 * tid : nullptr);
				rb[axis] = save;
				lb[axis] = unordered_bits(min(s.b_cutR[tid], s.b_rmin[tid]));
				write_box_keep(g, 2*node + 2, lb, rb, pch, bnext ? bnext + 2 * tid + 1 : nullptr);
			}
			cb ^= 1;
		}
		else
		{
			// ================= ranking mode: blocks of <= 32 slots, and the last level (a true sort) =================
			// (a) keys by position; the pads of a block get the largest key, so that a block is scanned without bounds
			for (int p = tid; p < P2; p += kBottomThreads)
			{
				const u32 slot = oin[p];
				s.ckey[p] = slot != kNoSlot ? ordered_bits(s.c[g.splitdim[node0 + (p >> logB)] * kBottomCap + slot]) : 0xffffffffu;
			}
			__syncthreads();
			// (b) rank by counting smaller keys of the block; equal keys (rare) take the slow path
			for (int p = tid; p < P2; p += kBottomThreads)
			{
				const u32 slot = oin[p];
				if (slot == kNoSlot) continue;
				const int q = p >> logB, qB = q << logB;
				const u32 kc = s.ckey[p];
				u32 rank = 0, ties = 0;
				if (B >= 4)
				{
					const uint4 *kv = reinterpret_cast<const uint4 *>(s.ckey + qB);
					for (int f = 0; f < (B >> 2); ++f)
					{
						const uint4 k4 = kv[f];
						rank += (k4.x < kc ? 1u : 0u) + (k4.y < kc ? 1u : 0u) + (k4.z < kc ? 1u : 0u) + (k4.w < kc ? 1u : 0u);
						ties += (k4.x == kc ? 1u : 0u) + (k4.y == kc ? 1u : 0u) + (k4.z == kc ? 1u : 0u) + (k4.w == kc ? 1u : 0u);
					}
				}
				else
					for (int f = 0; f < B; ++f) { const u32 kf = s.ckey[qB + f]; rank += kf < kc ? 1u : 0u; ties += kf == kc ? 1u : 0u; }
				if (ties > 1u)
				{
					const int chain = g.chain[node0 + q];
					for (int f = 0; f < B; ++f)
					{
						const u32 sf = oin[qB + f];
						if (qB + f != p && sf != kNoSlot && s.ckey[qB + f] == kc && slot_tie_less(s, sf, slot, chain, pay0)) ++rank;
					}
				}
				int dst = qB + (int)rank;
				if (!last)
				{
					const int64_t i = i0 + q;
					const int kl = (int)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
					if ((int)rank >= kl) dst = qB + (B >> 1) + ((int)rank - kl);
				}
				oout[dst] = (u16)slot;
			}
			__syncthreads();
			// (c) boxes of the children: the particles of rank kl-1 and kl sit at known positions of the output
			for (int q = tid; q < nblk; q += kBottomThreads)
			{
				const int64_t i = i0 + q;
				const int node = node0 + q, qB = q << logB;
				const int kl = (int)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
				const int axis = g.splitdim[node], pch = g.chain[node];
				float lb[3], rb[3];
				for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
				const float cl = s.c[axis * kBottomCap + oout[qB + kl - 1]];
				const float cr = s.c[axis * kBottomCap + oout[last ? qB + kl : qB + (B >> 1)]];
				const float save = rb[axis];
				rb[axis] = cl;
				write_box(g, 2*node + 1, lb, rb, pch);
				rb[axis] = save; lb[axis] = cr;
				write_box(g, 2*node + 2, lb, rb, pch);
			}
		}
		__syncthreads();
		u16 *tmp = oin; oin = oout; oout = tmp;
	}
	// output: storage order = order after the level-(L-1) sort
	{
		const int j = nlev - 1, l = L - 1;
		const int B = P2 >> j, logB = 31 - __clz(B);
		for (int p = tid; p < P2; p += kBottomThreads)
		{
			const u32 slot = oin[p];
			if (slot == kNoSlot) continue;
			const int q = p >> logB, t = p & (B - 1);
			const int64_t dst = seg_start(n, ((int64_t)b << j) + q, l) + t;
			perm[dst] = (int)__float_as_uint(pay0[slot].w);
			spos[3*dst] = s.c[slot]; spos[3*dst+1] = s.c[kBottomCap + slot]; spos[3*dst+2] = s.c[2 * kBottomCap + slot];
		}
	}
}

constexpr size_t kBottomSmemBytes = sizeof(float) * 3 * kBottomCap + sizeof(u32) * kBottomCap + sizeof(u32) * (kBottomCap / 8)
                                    + 2 * sizeof(u16) * kBottomCap + 2 * sizeof(BlkBox) * kMaxHistBlk + 12 * 4 * kMaxHistBlk;

} // namespace

int kd_reserve(nbco_ctx *ctx, KdTree &t, int64_t n, int L)
{
	(void)ctx;
	t.n = n; t.L = L;
	int lt = 0;
	while (((n - 1) >> lt) + 1 > kBottomCap) ++lt; // first level whose segments fit a bottom CTA
	// the last sorting level must run in the bottom kernel (the top levels only partition)
	t.lt = std::min(lt, L - 1);
	if (((n - 1) >> t.lt) + 1 > kBottomCap)
	{
		set_error("max_level %d leaves %lld particles per level-%d node; at most %d are supported", L,
		          (long long)(((n - 1) >> t.lt) + 1), t.lt, kBottomCap);
		return NBCO_ERR_INVALID;
	}
	const size_t nt = ((size_t)1 << (L + 1)) - 1;
	NBCO_TRY(t.lbound.reserve(12 * nt)); NBCO_TRY(t.rbound.reserve(12 * nt)); NBCO_TRY(t.size2.reserve(4 * nt));
	NBCO_TRY(t.splitdim.reserve(4 * nt)); NBCO_TRY(t.chain.reserve(4 * nt));
	NBCO_TRY(t.payA.reserve(16 * (size_t)n));
	if (t.lt > 0) NBCO_TRY(t.payB.reserve(16 * (size_t)n));
	NBCO_TRY(t.spos.reserve(12 * (size_t)n)); NBCO_TRY(t.perm.reserve(4 * (size_t)n));
	const size_t nseg = (size_t)1 << std::max(t.lt - 1, 0);
	NBCO_TRY(t.hist.reserve(4 * (size_t)kBins * nseg));
	NBCO_TRY(t.seg.reserve(sizeof(SegState) * nseg));
	NBCO_TRY(t.bbox.reserve(64));
	return NBCO_OK;
}

void kd_release(KdTree &t)
{
	DevBuf *all[] = {&t.lbound, &t.rbound, &t.size2, &t.splitdim, &t.chain, &t.payA, &t.payB, &t.hist, &t.seg, &t.spos, &t.perm, &t.bbox};
	for (DevBuf *b : all) b->release();
}

int kd_build(nbco_ctx *ctx, KdTree &t, const float *pos, cudaEvent_t ev_bottom, int r, int g)
{
	cudaStream_t st = ctx->stream;
	const int64_t n = t.n;
	TreeGeom tg{t.lbound.as<float>(), t.rbound.as<float>(), t.size2.as<float>(), t.splitdim.as<int>(), t.chain.as<int>()};
	u32 *bb = t.bbox.as<u32>();
	static const u32 bb_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
	NBCO_CUDA(cudaMemcpyAsync(bb, bb_init, sizeof(bb_init), cudaMemcpyHostToDevice, st));
	float4 *pay[2] = {t.payA.as<float4>(), t.payB.as<float4>()};
	pack_bbox_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, st>>>(pos, pay[0], n, bb);
	root_box_kernel<<<1, 32, 0, st>>>(tg, bb);
	ctx->launches += 2;

	u32 *hist = t.hist.as<u32>();
	SegState *seg = t.seg.as<SegState>();
	const int ltop = t.lt; // levels [0, ltop) are partitioned globally
	if (ltop > 0) NBCO_CUDA(cudaMemsetAsync(hist, 0, 4 * (size_t)kBins * ((size_t)1 << (ltop - 1)), st));
	int cur = 0;
	for (int l = 0; l < ltop; ++l)
	{
		// below level g only the segments of rank r's subtree (multi-GPU: the other subtrees are built by their owners)
		const int nseg = l >= g ? 1 << (l - g) : 1 << l, seg0 = l >= g ? r << (l - g) : 0;
		const int64_t maxseg = ((n - 1) >> l) + 1;
		// tiles: about four CTAs per SM over the level, a whole number of reservation chunks each
		int64_t want = (maxseg * nseg + (int64_t)ctx->sm_count * 4 - 1) / ((int64_t)ctx->sm_count * 4);
		int64_t tile = ((std::max<int64_t>(want, kChunk) + kChunk - 1) / kChunk) * kChunk;
		const int tps = (int)((maxseg + tile - 1) / tile);
		const int tiles = nseg * tps;
		top_hist_kernel<<<tiles, kTopThreads, 0, st>>>(pay[cur], tg, hist, n, l, tps, (int)tile, seg0);
		top_pick_kernel<<<nseg, 256, 0, st>>>(seg, hist, n, l, seg0);
		top_partition_kernel<<<tiles, kTopThreads, 0, st>>>(pay[cur], pay[cur ^ 1], tg, seg, n, l, tps, (int)tile, seg0);
		top_resolve_kernel<<<nseg, kResThreads, 0, st>>>(pay[cur ^ 1], pay[cur], tg, seg, n, l, seg0);
		ctx->launches += 4;
		cur ^= 1;
	}
	if (ev_bottom) NBCO_CUDA(cudaEventRecord(ev_bottom, st));
	if (!t.bottom_attr)
	{
		NBCO_CUDA(cudaFuncSetAttribute(kd_bottom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBottomSmemBytes));
		t.bottom_attr = true;
	}
	int64_t maxseg = ((n - 1) >> ltop) + 1;
	int P2 = 2; while (P2 < maxseg) P2 <<= 1;
	while ((P2 >> (t.L - 1 - ltop)) < 2) P2 <<= 1; // the last level sorts blocks of at least 2 slots
	if (P2 > kBottomCap) { set_error("internal: bottom block %d", P2); return NBCO_ERR_INVALID; }
	if (g > ltop) { set_error("more ranks than shared-memory kd blocks (2^%d > 2^%d)", g, ltop); return NBCO_ERR_INVALID; }
	kd_bottom_kernel<<<1 << (ltop - g), kBottomThreads, kBottomSmemBytes, st>>>(tg, pay[cur], t.spos.as<float>(), t.perm.as<int>(),
	                                                                            n, ltop, t.L, P2, r << (ltop - g));
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

} // namespace nbco
