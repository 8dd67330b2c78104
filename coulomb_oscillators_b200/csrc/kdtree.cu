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
//   * shallow trees (max_level so small that a level-(L-1) node exceeds kBottomCap): the order inside a leaf is the
//     order of its PARENT's sort, so the build simply continues below the leaves with virtual levels that inherit the
//     parent's axis and tie-break chain (TreeGeom::baxis / bchain) until a segment fits a bottom CTA; the tree's own
//     arrays are a prefix of the build's.

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
		write_box(g, 0, lb, rb, kNoAxis | (kNoAxis << 2) | (kNoAxis << 4), kNoAxis);
	}
}

// =====================================================================================
//  top levels: histogram select + three-way partition per segment
// =====================================================================================
struct TileRange { int64_t a, b, s0; int seg; };

// lb != nullptr (distributed top levels, peer mode): segment s of this rank's records is [base + lb[s], base + lb[s+1])
// (a table on the device: the local sizes are data dependent); else the global rule [seg_start(s), seg_start(s + 1))
__device__ __forceinline__ TileRange tile_range(int64_t n, int l, int tps, int tile, int seg0, const int64_t *lb = nullptr, int64_t base = 0)
{
	TileRange r;
	r.seg = seg0 + blockIdx.x / tps;
	const int t = blockIdx.x % tps;
	int64_t s1;
	if (lb) { r.s0 = base + lb[r.seg]; s1 = base + lb[r.seg + 1]; }
	else { r.s0 = seg_start(n, r.seg, l); s1 = seg_start(n, r.seg + 1, l); }
	r.a = r.s0 + (int64_t)t * tile;
	r.b = r.a + tile < s1 ? r.a + tile : s1;
	return r;
}

// histogram of the linear bins of every segment of level l
__global__ void __launch_bounds__(kTopThreads)
top_hist_kernel(const float4 *__restrict__ pay, TreeGeom g, u32 *__restrict__ hist, int64_t n, int l, int tps, int tile, int seg0,
                const int64_t *__restrict__ lb = nullptr, int64_t base = 0)
{
	__shared__ u32 sh[kBins];
	for (int b = threadIdx.x; b < kBins; b += kTopThreads) sh[b] = 0;
	__syncthreads();
	const TileRange r = tile_range(n, l, tps, tile, seg0, lb, base);
	const int node = kd_beg(l) + r.seg, axis = g.baxis[node];
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
                     int64_t n, int l, int tps, int tile, int seg0, const int64_t *__restrict__ lb = nullptr, int64_t rbase = 0)
{
	constexpr int kWarps = kTopThreads / 32;
	__shared__ u32 wcnt[kWarps][3];
	__shared__ u32 base[3];
	const TileRange r = tile_range(n, l, tps, tile, seg0, lb, rbase);
	if (r.a >= r.b) return;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const u32 lt_mask = (1u << lane) - 1u;
	SegState *sp = st + (r.seg - seg0);
	const u32 pb = sp->pb, less = sp->less, eq = sp->eq;
	const int node = kd_beg(l) + r.seg, axis = g.baxis[node];
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
	const int node = kd_beg(l) + seg, axis = g.baxis[node], chain = g.bchain[node];
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
			for (u32 i = tid; i < eq; i += kResThreads)
			{
				const Words W = words_of(C[i], chain);
				bool m = true;
				for (int e = 0; e < d; ++e) m = m && W.w[e] == piv[e];
				if (pass > 0) m = m && (W.w[d] >> (shift + (pass == 1 ? 11 : 10))) == (prefix >> (shift + (pass == 1 ? 11 : 10)));
				if (m) atomicAdd(&sh[(W.w[d] >> shift) & mask], 1u);
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
	if (tid == 0) { run[0] = 0; run[1] = 0; s_minr = 0xffffffffu; }
	__syncthreads();
	u32 minr = 0xffffffffu;
	for (u32 base = 0; base < eq; base += kResThreads)
	{
		const u32 i = base + tid;
		const bool valid = i < eq;
		float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
		bool left = false;
		if (valid)
		{
			p = C[i];
			const Words W = words_of(p, chain);
			left = true; // equal on every compared word: the pivot itself or a tie that goes left
			for (int e = 0; e < depth; ++e)
				if (W.w[e] != piv[e]) { left = W.w[e] < piv[e]; break; }
			if (!left) minr = min(minr, W.w[0]);
		}
		const u32 bl = __ballot_sync(0xffffffffu, valid && left), br = __ballot_sync(0xffffffffu, valid && !left);
		if (lane == 0) { wsum[w] = (u32)__popc(bl) | ((u32)__popc(br) << 16); }
		__syncthreads();
		u32 ol = run[0], orr = run[1];
		for (int k = 0; k < w; ++k) { ol += wsum[k] & 0xffffu; orr += wsum[k] >> 16; }
		if (valid) T[left ? ol + __popc(bl & ((1u << lane) - 1u)) : need + orr + __popc(br & ((1u << lane) - 1u))] = p;
		__syncthreads();
		if (tid == 0)
		{
			u32 tl = 0, tr = 0;
			for (int k = 0; k < 32; ++k) { tl += wsum[k] & 0xffffu; tr += wsum[k] >> 16; }
			run[0] += tl; run[1] += tr;
		}
		__syncthreads();
	}
	minr = __reduce_min_sync(0xffffffffu, minr);
	if (lane == 0 && minr != 0xffffffffu) atomicMin(&s_minr, minr);
	__threadfence_block();
	__syncthreads();
	for (u32 i = tid; i < eq; i += kResThreads) C[i] = T[i];

	// ---- boxes of the children: cut at the last particle of the left and the first of the right child ----
	if (tid == 0)
	{
		float lb[3], rb[3];
		for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
		const int pch = chain;
		const float save = rb[axis];
		rb[axis] = unordered_bits(piv[0]);
		write_box(g, 2*node + 1, lb, rb, pch, axis);
		rb[axis] = save;
		lb[axis] = unordered_bits(min(s_minr, s.rmin));
		write_box(g, 2*node + 2, lb, rb, pch, axis);
	}
}

// =====================================================================================
//  distributed top levels (peer mode): rank r holds N/G records; the first g = log2 G levels
//  select their medians from per-rank histograms summed through the published scratch blocks, every rank splits its own
//  records, the candidates of the pivot bin are ordered from the union of all ranks' candidates (read over NVLink), and
//  after level g - 1 every rank pulls the records of its own subtree from wherever they are.  No rank ever holds, packs
//  or partitions all N records (round 1 pulled all positions and built the top g levels over all particles everywhere).
// =====================================================================================
// scratch block of a rank (kPeerScratch bytes, published): byte offsets
constexpr size_t kScrBBox = 0;                 // u32[6]: ordered bits of the local bounding box
constexpr size_t kScrRange = 64;               // int64 lb[level 0 .. g][9]: local segment boundaries (relative to the rank's range)
constexpr size_t kScrSeg = 1024;               // SegState[level][8]: LOCAL selection state (less = local count below the pivot bin,
                                               // eq = local candidates, rmin = local minimum key right of the pivot bin)
constexpr size_t kScrGlob = 2048;              // u32[level][8][2]: GLOBAL {less, eq} of the segment
constexpr size_t kScrPivot = 2560;             // u32[level][8][8]: pivot words [0..3], depth, smallest right key (select -> split)
constexpr size_t kScrHist = 4096;              // u32[level][segment][kBins]: local histograms, level l starts at (2^l - 1) rows
static_assert(kScrHist + 7 * kBins * 4 <= kPeerScratch, "scratch block too small for 8 ranks");

// sanity words of the distributed build (published header, byte 768; printed by fmm3_harvest when NBCO_DEBUG_KD is set)
__device__ __forceinline__ u32 *kd_dbg(unsigned char *scr) { return reinterpret_cast<u32 *>(scr - 256); }

struct PeerKd
{
	int world, me, g;
	unsigned char *scr[kMaxPeers];   // scratch blocks
	float4 *pay[kMaxPeers][2];       // record buffers
	int64_t lo[kMaxPeers + 1];       // global range of rank q: [lo[q], lo[q+1])
};
__device__ __forceinline__ int64_t *scr_range(unsigned char *s, int l) { return reinterpret_cast<int64_t *>(s + kScrRange) + 9 * l; }
__device__ __forceinline__ SegState *scr_seg(unsigned char *s, int l) { return reinterpret_cast<SegState *>(s + kScrSeg) + 8 * l; }
__device__ __forceinline__ u32 *scr_glob(unsigned char *s, int l) { return reinterpret_cast<u32 *>(s + kScrGlob) + 16 * l; }
__device__ __forceinline__ u32 *scr_pivot(unsigned char *s, int l) { return reinterpret_cast<u32 *>(s + kScrPivot) + 64 * l; }
__device__ __forceinline__ u32 *scr_hist(unsigned char *s, int l) { return reinterpret_cast<u32 *>(s + kScrHist) + (size_t)((1 << l) - 1) * kBins; }

__global__ void __launch_bounds__(256) pack_own_kernel(const float *__restrict__ pos, float4 *__restrict__ pay, int64_t lo, int64_t hi, u32 *__restrict__ out6)
{
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
	{
		const float x = pos[3*i], y = pos[3*i+1], z = pos[3*i+2];
		pay[i] = make_float4(x, y, z, __uint_as_float((u32)i)); // id = global index in the previous tree order
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

__global__ void bbox_init_kernel(u32 *bb)
{
	if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
	else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

// root box = union of the ranks' boxes (identical on every rank); local range table of level 0
__global__ void root_box_peer_kernel(TreeGeom g, PeerKd pk)
{
	if (threadIdx.x == 0 && blockIdx.x == 0)
	{
		u32 mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
		for (int q = 0; q < pk.world; ++q)
		{
			const volatile u32 *bb = reinterpret_cast<const volatile u32 *>(pk.scr[q] + kScrBBox);
			for (int k = 0; k < 3; ++k)
			{
				const u32 lo_k = bb[k], hi_k = bb[3 + k]; // plain unsigned compares (no overload resolution on volatile operands)
				if (lo_k < mn[k]) mn[k] = lo_k;
				if (hi_k > mx[k]) mx[k] = hi_k;
			}
		}
		float lb[3], rb[3];
		for (int k = 0; k < 3; ++k) { lb[k] = unordered_bits(mn[k]); rb[k] = unordered_bits(mx[k]); }
		write_box(g, 0, lb, rb, kNoAxis | (kNoAxis << 2) | (kNoAxis << 4), kNoAxis);
		int64_t *r0 = scr_range(pk.scr[pk.me], 0);
		r0[0] = 0; r0[1] = pk.lo[pk.me + 1] - pk.lo[pk.me];
		u32 *dw = kd_dbg(pk.scr[pk.me]);
		for (int k = 0; k < 3; ++k) { dw[32 + k] = mn[k]; dw[35 + k] = mx[k]; }
	}
}

// one CTA per segment: sum the ranks' histograms, find the pivot bin (global rank), derive the LOCAL split counts
__global__ void __launch_bounds__(256) top_pick_peer_kernel(PeerKd pk, int64_t n, int l)
{
	constexpr int kPer = kBins / 256;
	__shared__ u32 wsum[8];
	__shared__ u32 s_pb, s_less, s_eq, s_nl;
	const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const u32 krank = (u32)(seg_start(n, 2 * (int64_t)seg + 1, l + 1) - seg_start(n, 2 * (int64_t)seg, l + 1)) - 1u;
	u32 c[kPer], own[kPer], s = 0;
#pragma unroll
	for (int k = 0; k < kPer; ++k) c[k] = 0;
	for (int q = 0; q < pk.world; ++q)
	{
		const volatile u32 *h = scr_hist(pk.scr[q], l) + (size_t)seg * kBins;
#pragma unroll
		for (int k = 0; k < kPer; ++k) { const u32 v = h[tid * kPer + k]; c[k] += v; if (q == pk.me) own[k] = v; }
	}
#pragma unroll
	for (int k = 0; k < kPer; ++k) s += c[k];
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
			if (krank >= run && krank < run + c[k]) { s_pb = (u32)(tid * kPer + k); s_less = run; s_eq = c[k]; }
			run += c[k];
		}
	}
	if (tid == 0) s_nl = 0;
	__syncthreads();
	// local records below the pivot bin
	const u32 pb = s_pb;
	u32 below = 0, at = 0;
#pragma unroll
	for (int k = 0; k < kPer; ++k)
	{
		const u32 bin = (u32)(tid * kPer + k);
		if (bin < pb) below += own[k];
		if (bin == pb) at = own[k];
	}
	below = __reduce_add_sync(0xffffffffu, below);
	if (lane == 0 && below) atomicAdd(&s_nl, below);
	__syncthreads();
	if (tid * kPer <= (int)pb && (int)pb < (tid + 1) * kPer)
	{
		SegState z;
		z.pb = pb; z.less = s_nl; z.eq = at; z.curL = z.curE = z.curR = 0; z.rmin = 0xffffffffu; z.pad = 0;
		scr_seg(pk.scr[pk.me], l)[seg] = z;
		u32 *gl = scr_glob(pk.scr[pk.me], l) + 2 * seg;
		gl[0] = s_less; gl[1] = s_eq;
	}
}

// union of the ranks' candidate lists of a segment, addressed by one index
struct CandUnion
{
	const float4 *base[kMaxPeers];
	u32 pre[kMaxPeers + 1];
	int world;
	__device__ __forceinline__ float4 at(u32 i) const
	{
		int q = 0;
		while (q + 1 < world && i >= pre[q + 1]) ++q;
		return base[q][i - pre[q]];
	}
};

// One CTA per segment: the pivot of the UNION of all ranks' candidates (lexicographic radix select, identical on every
// rank).  The result goes to this rank's scratch; the candidates are only READ here -- every rank reorders its own ones in
// top_split_peer_kernel, after a barrier (a peer may still be scanning them).
__global__ void __launch_bounds__(kResThreads)
top_select_peer_kernel(PeerKd pk, TreeGeom g, int cur /* buffer that holds the partitioned records */, int64_t n, int l)
{
	__shared__ u32 sh[kBins];
	__shared__ u32 wsum[32];
	__shared__ u32 found[3];
	__shared__ u32 run[2];
	__shared__ u32 s_minr;
	__shared__ CandUnion cu;
	const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	unsigned char *myscr = pk.scr[pk.me];
	const int node = kd_beg(l) + seg, axis = g.baxis[node], chain = g.bchain[node];
	const u32 kleft = (u32)(seg_start(n, 2 * (int64_t)seg + 1, l + 1) - seg_start(n, 2 * (int64_t)seg, l + 1));
	if (tid == 0)
	{
		u32 acc = 0, rmin = 0xffffffffu;
		for (int q = 0; q < pk.world; ++q)
		{
			const volatile SegState *sq = scr_seg(pk.scr[q], l) + seg;
			const volatile int64_t *lbq = scr_range(pk.scr[q], l);
			cu.base[q] = pk.pay[q][cur] + pk.lo[q] + lbq[seg] + sq->less;
			cu.pre[q] = acc;
			acc += sq->eq;
			const u32 rq = sq->rmin;
			if (rq < rmin) rmin = rq;
		}
		cu.pre[pk.world] = acc;
		cu.world = pk.world;
		s_minr = rmin; // global minimum key right of the pivot bin
		run[0] = 0; run[1] = 0;
	}
	__syncthreads();
	const u32 *gl = scr_glob(myscr, l) + 2 * seg;
	const u32 eq = cu.pre[pk.world], need = kleft - gl[0]; // 1 <= need <= eq
	if (tid == 0 && (eq != gl[1] || need < 1u || need > eq))
	{
		u32 *dw = kd_dbg(myscr);
		dw[0] = 100u + (u32)l; dw[1] = (u32)seg; dw[2] = eq; dw[3] = gl[1]; dw[4] = need; dw[5] = kleft; dw[6] = gl[0];
	}

	// ---- select over the union ----
	u32 piv[4] = {0u, 0u, 0u, 0u};
	int depth = 0;
	u32 r = need - 1, t = eq;
	for (int d = 0; d < 4; ++d)
	{
		if ((d == 1 || d == 2) && ((chain >> (2 * d)) & 3) == kNoAxis) { depth = d + 1; continue; }
		u32 prefix = 0;
		for (int pass = 0; pass < 3; ++pass)
		{
			const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
			const u32 mask = pass == 2 ? 1023u : 2047u;
			for (int b = tid; b < kBins; b += kResThreads) sh[b] = 0;
			__syncthreads();
			for (u32 i0 = tid; i0 < eq; i0 += 4 * kResThreads)
			{
				float4 q4[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) { const u32 i = i0 + k * kResThreads; q4[k] = i < eq ? cu.at(i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
				for (int k = 0; k < 4; ++k)
				{
					if (i0 + k * kResThreads >= eq) continue;
					const Words W = words_of(q4[k], chain);
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
		if (t == 1u || r == t - 1u) break;
	}
	// ---- smallest key among the union's candidates that go right ----
	u32 minr = 0xffffffffu;
	for (u32 i = tid; i < eq; i += kResThreads)
	{
		const Words W = words_of(cu.at(i), chain);
		bool left = true;
		for (int e = 0; e < depth; ++e)
			if (W.w[e] != piv[e]) { left = W.w[e] < piv[e]; break; }
		if (!left) minr = min(minr, W.w[0]);
	}
	minr = __reduce_min_sync(0xffffffffu, minr);
	if (lane == 0 && minr != 0xffffffffu) atomicMin(&s_minr, minr);
	__syncthreads();
	if (tid == 0)
	{
		u32 *pv = scr_pivot(myscr, l) + 8 * seg;
		pv[0] = piv[0]; pv[1] = piv[1]; pv[2] = piv[2]; pv[3] = piv[3]; pv[4] = (u32)depth; pv[5] = s_minr;
	}
}

// One CTA per segment: split this rank's own candidates around the pivot (through the other buffer), write the children's
// boxes and the local range table of the next level
__global__ void __launch_bounds__(kResThreads)
top_split_peer_kernel(PeerKd pk, TreeGeom g, int cur, int64_t n, int l)
{
	__shared__ u32 wsum[32];
	__shared__ u32 run[2];
	const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	unsigned char *myscr = pk.scr[pk.me];
	const int node = kd_beg(l) + seg, axis = g.baxis[node], chain = g.bchain[node];
	const u32 *pv = scr_pivot(myscr, l) + 8 * seg;
	const u32 piv[4] = {pv[0], pv[1], pv[2], pv[3]};
	const int depth = (int)pv[4];
	const u32 minr_all = pv[5];
	if (tid == 0) { run[0] = 0; run[1] = 0; }
	__syncthreads();
	// ---- split this rank's own candidates through the other buffer ----
	const SegState mine = scr_seg(myscr, l)[seg];
	const int64_t *lbm = scr_range(myscr, l);
	float4 *C = pk.pay[pk.me][cur] + pk.lo[pk.me] + lbm[seg] + mine.less;
	float4 *T = pk.pay[pk.me][cur ^ 1] + pk.lo[pk.me] + lbm[seg] + mine.less;
	const u32 ne = mine.eq;
	// first pass: how many of the own candidates go left (the right ones are placed behind them)
	u32 nel = 0;
	for (u32 i = tid; i < ne; i += kResThreads)
	{
		const Words W = words_of(C[i], chain);
		bool left = true;
		for (int e = 0; e < depth; ++e)
			if (W.w[e] != piv[e]) { left = W.w[e] < piv[e]; break; }
		nel += left ? 1u : 0u;
	}
	nel = __reduce_add_sync(0xffffffffu, nel);
	__syncthreads();
	if (lane == 0) wsum[w] = nel;
	__syncthreads();
	u32 nleft = 0;
	for (int k = 0; k < 32; ++k) nleft += wsum[k];
	__syncthreads();
	for (u32 base = 0; base < ne; base += kResThreads)
	{
		const u32 i = base + tid;
		const bool valid = i < ne;
		float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
		bool left = false;
		if (valid)
		{
			p = C[i];
			const Words W = words_of(p, chain);
			left = true;
			for (int e = 0; e < depth; ++e)
				if (W.w[e] != piv[e]) { left = W.w[e] < piv[e]; break; }
		}
		const u32 bl = __ballot_sync(0xffffffffu, valid && left), br = __ballot_sync(0xffffffffu, valid && !left);
		if (lane == 0) wsum[w] = (u32)__popc(bl) | ((u32)__popc(br) << 16);
		__syncthreads();
		u32 ol = run[0], orr = run[1];
		for (int k = 0; k < w; ++k) { ol += wsum[k] & 0xffffu; orr += wsum[k] >> 16; }
		if (valid) T[left ? ol + __popc(bl & ((1u << lane) - 1u)) : nleft + orr + __popc(br & ((1u << lane) - 1u))] = p;
		__syncthreads();
		if (tid == 0)
		{
			u32 tl = 0, tr = 0;
			for (int k = 0; k < 32; ++k) { tl += wsum[k] & 0xffffu; tr += wsum[k] >> 16; }
			run[0] += tl; run[1] += tr;
		}
		__syncthreads();
	}
	__threadfence_block();
	__syncthreads();
	for (u32 i = tid; i < ne; i += kResThreads) C[i] = T[i];

	if (tid == 0)
	{
		float lb[3], rb[3];
		for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
		const float save = rb[axis];
		rb[axis] = unordered_bits(piv[0]);
		write_box(g, 2*node + 1, lb, rb, chain, axis);
		rb[axis] = save;
		lb[axis] = unordered_bits(minr_all);
		write_box(g, 2*node + 2, lb, rb, chain, axis);
		// local ranges of the children
		u32 *dw = kd_dbg(myscr);
		dw[20 + 4 * l] = mine.less; dw[21 + 4 * l] = ne; dw[22 + 4 * l] = nleft; dw[23 + 4 * l] = (u32)(lbm[seg + 1] - lbm[seg]);
		int64_t *nx = scr_range(myscr, l + 1);
		nx[2 * seg] = lbm[seg];
		nx[2 * seg + 1] = lbm[seg] + mine.less + nleft;
		nx[2 * seg + 2] = lbm[seg + 1];
	}
}

// after level g - 1 the local records are grouped by destination rank: every rank pulls the pieces of its own subtree
__global__ void __launch_bounds__(256) exchange_pull_kernel(PeerKd pk, int cur, float4 *__restrict__ dst /* own range of the other buffer */)
{
	__shared__ int64_t pre[kMaxPeers + 1];
	__shared__ const float4 *src[kMaxPeers];
	if (threadIdx.x == 0)
	{
		int64_t acc = 0;
		for (int q = 0; q < pk.world; ++q)
		{
			const volatile int64_t *lbq = scr_range(pk.scr[q], pk.g);
			src[q] = pk.pay[q][cur] + pk.lo[q] + lbq[pk.me];
			pre[q] = acc;
			acc += lbq[pk.me + 1] - lbq[pk.me];
		}
		pre[pk.world] = acc;
	}
	__syncthreads();
	const int64_t total = pre[pk.world], stride = (int64_t)gridDim.x * blockDim.x;
	if (blockIdx.x == 0 && threadIdx.x == 0 && total != pk.lo[pk.me + 1] - pk.lo[pk.me])
	{
		u32 *dw = kd_dbg(pk.scr[pk.me]);
		dw[8] = 200u; dw[9] = (u32)total; dw[10] = (u32)(pk.lo[pk.me + 1] - pk.lo[pk.me]);
		for (int q = 0; q < pk.world; ++q) dw[11 + q] = (u32)(pre[q + 1] - pre[q]);
	}
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
	{
		int q = 0;
		while (q + 1 < pk.world && t >= pre[q + 1]) ++q;
		dst[t] = src[q][t - pre[q]];
	}
}

// =====================================================================================
//  bottom levels: one CTA per level-lt node, particles resident in shared memory
// =====================================================================================
constexpr int kMaxHistBlk = kBottomCap / 64; // histogram mode: blocks of >= 64 slots
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
__device__ __forceinline__ void write_box_keep(const TreeGeom &g, int node, const float *lb, const float *rb, int parent_chain, int parent_axis, BlkBox *keep)
{
	write_box(g, node, lb, rb, parent_chain, parent_axis);
	if (keep)
	{
		const int ax = widest_axis(rb[0] - lb[0], rb[1] - lb[1], rb[2] - lb[2]);
		const bool inh = kd_inherits(g, node);
		for (int k = 0; k < 3; ++k) { keep->lb[k] = lb[k]; keep->rb[k] = rb[k]; }
		keep->axis = inh ? parent_axis : ax; keep->chain = inh ? parent_chain : chain_push(ax, parent_chain);
	}
}

__global__ void __launch_bounds__(kBottomThreads, kBottomCtasPerSm)
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
		bx.axis = g.baxis[node]; bx.chain = g.bchain[node];
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
				constexpr int kWarps = kBottomThreads / 32;
				const int wpb = nblk >= kWarps ? 1 : kWarps / nblk, ngroups = kWarps / wpb;
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
				write_box_keep(g, 2*node + 1, lb, rb, pch, axis, bnext ? bnext + 2 * tid : nullptr);
				rb[axis] = save;
				lb[axis] = unordered_bits(min(s.b_cutR[tid], s.b_rmin[tid]));
				write_box_keep(g, 2*node + 2, lb, rb, pch, axis, bnext ? bnext + 2 * tid + 1 : nullptr);
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
				s.ckey[p] = slot != kNoSlot ? ordered_bits(s.c[g.baxis[node0 + (p >> logB)] * kBottomCap + slot]) : 0xffffffffu;
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
					const int chain = g.bchain[node0 + q];
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
				const int axis = g.baxis[node], pch = g.bchain[node];
				float lb[3], rb[3];
				for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
				const float cl = s.c[axis * kBottomCap + oout[qB + kl - 1]];
				const float cr = s.c[axis * kBottomCap + oout[last ? qB + kl : qB + (B >> 1)]];
				const float save = rb[axis];
				rb[axis] = cl;
				write_box(g, 2*node + 1, lb, rb, pch, axis);
				rb[axis] = save; lb[axis] = cr;
				write_box(g, 2*node + 2, lb, rb, pch, axis);
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
	// The last sorting level must run in the bottom kernel (the top levels only partition).  Ordinary trees: lt <= L - 1.
	// Shallow trees (fmm_cart3_kdtree.cuh:1508-1512 accepts any max_level): the build continues below the leaves with
	// virtual levels that split along the parent's axis (TreeGeom::baxis) until a segment fits a bottom CTA.
	t.Lb = std::max(L, lt + 1);
	t.lt = std::min(lt, t.Lb - 1);
	if (t.Lb > 30) { set_error("kd build: depth %d", t.Lb); return NBCO_ERR_INVALID; }
	const size_t nt = ((size_t)1 << (t.Lb + 1)) - 1;
	NBCO_TRY(t.lbound.reserve(12 * nt)); NBCO_TRY(t.rbound.reserve(12 * nt)); NBCO_TRY(t.size2.reserve(4 * nt));
	NBCO_TRY(t.splitdim.reserve(4 * nt)); NBCO_TRY(t.chain.reserve(4 * nt));
	if (t.Lb > L) { NBCO_TRY(t.baxis.reserve(4 * nt)); NBCO_TRY(t.bchain.reserve(4 * nt)); }
	NBCO_TRY(t.payA.reserve(16 * (size_t)n));
	if (t.lt > 0) NBCO_TRY(t.payB.reserve(16 * (size_t)n));
	NBCO_TRY(t.spos.reserve(12 * (size_t)n)); NBCO_TRY(t.perm.reserve(4 * (size_t)n));
	const size_t nseg = (size_t)1 << std::max(t.lt - 1, 0);
	NBCO_TRY(t.hist.reserve(4 * (size_t)kBins * nseg));
	NBCO_TRY(t.seg.reserve(sizeof(SegState) * nseg));
	NBCO_TRY(t.bbox.reserve(64));
	return NBCO_OK;
}

static TreeGeom geom_of(KdTree &t)
{
	const bool deep = t.Lb > t.L;
	return TreeGeom{t.lbound.as<float>(), t.rbound.as<float>(), t.size2.as<float>(), t.splitdim.as<int>(), t.chain.as<int>(),
	                deep ? t.baxis.as<int>() : t.splitdim.as<int>(), deep ? t.bchain.as<int>() : t.chain.as<int>(), kd_beg(t.L)};
}

void kd_release(KdTree &t)
{
	DevBuf *all[] = {&t.lbound, &t.rbound, &t.size2, &t.splitdim, &t.chain, &t.baxis, &t.bchain, &t.payA, &t.payB, &t.hist, &t.seg, &t.spos, &t.perm, &t.bbox};
	for (DevBuf *b : all) b->release();
}

// Load the modules of every kernel of the build NOW (CUDA loads a kernel lazily at its first launch, and that load
// synchronises the device: if it happens while another rank's flag-barrier kernel spins on the SAME device -- several ranks
// emulated on one GPU, tests/test_peer_gpu.py -- the ranks wait for each other until the barrier times out).
int kd_preload_kernels()
{
	cudaFuncAttributes fa;
	NBCO_CUDA(cudaFuncGetAttributes(&fa, bbox_init_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, pack_own_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, pack_bbox_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, root_box_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, root_box_peer_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_hist_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_pick_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_pick_peer_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_partition_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_resolve_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_select_peer_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, top_split_peer_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, exchange_pull_kernel));
	NBCO_CUDA(cudaFuncGetAttributes(&fa, kd_bottom_kernel));
	return NBCO_OK;
}

// levels [l0, lt) of the subtree(s) this rank builds alone, then the shared-memory kernel for the rest
static int kd_local_levels(nbco_ctx *ctx, KdTree &t, float4 *pay[2], int cur, int l0, cudaEvent_t ev_bottom, int r, int g)
{
	cudaStream_t st = ctx->stream;
	const int64_t n = t.n;
	TreeGeom tg = geom_of(t);
	u32 *hist = t.hist.as<u32>();
	SegState *seg = t.seg.as<SegState>();
	const int ltop = t.lt; // levels [0, ltop) are partitioned globally
	for (int l = l0; l < ltop; ++l)
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
	while ((P2 >> (t.Lb - 1 - ltop)) < 2) P2 <<= 1; // the last level sorts blocks of at least 2 slots
	if (P2 > kBottomCap) { set_error("internal: bottom block %d", P2); return NBCO_ERR_INVALID; }
	if (g > ltop) { set_error("more ranks than shared-memory kd blocks (2^%d > 2^%d)", g, ltop); return NBCO_ERR_INVALID; }
	kd_bottom_kernel<<<1 << (ltop - g), kBottomThreads, kBottomSmemBytes, st>>>(tg, pay[cur], t.spos.as<float>(), t.perm.as<int>(),
	                                                                            n, ltop, t.Lb, P2, r << (ltop - g));
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

int kd_build(nbco_ctx *ctx, KdTree &t, const float *pos, cudaEvent_t ev_bottom, int r, int g)
{
	cudaStream_t st = ctx->stream;
	const int64_t n = t.n;
	TreeGeom tg = geom_of(t);
	u32 *bb = t.bbox.as<u32>();
	static const u32 bb_init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
	NBCO_CUDA(cudaMemcpyAsync(bb, bb_init, sizeof(bb_init), cudaMemcpyHostToDevice, st));
	float4 *pay[2] = {t.payA.as<float4>(), t.payB.as<float4>()};
	pack_bbox_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, st>>>(pos, pay[0], n, bb);
	root_box_kernel<<<1, 32, 0, st>>>(tg, bb);
	ctx->launches += 2;
	if (t.lt > 0) NBCO_CUDA(cudaMemsetAsync(t.hist.p, 0, 4 * (size_t)kBins * ((size_t)1 << (t.lt - 1)), st));
	return kd_local_levels(ctx, t, pay, 0, 0, ev_bottom, r, g);
}

// Distributed build (peer mode, g = log2(world) <= lt): see the section "distributed top levels" above.  pub[q] = the
// published buffer of rank q as mapped on this device (common.cuh: layout); pos = this rank's array, of which only the own
// range [lo, hi) is read.  The ranks meet at 3 g + 3 flag barriers (peer_barrier) on their context streams.
int kd_build_peer(nbco_ctx *ctx, KdTree &t, const float *pos, cudaEvent_t ev_bottom, int r, int g, void *const *pub)
{
	cudaStream_t st = ctx->stream;
	const int64_t n = t.n;
	const int world = 1 << g;
	static const bool dbg = getenv("NBCO_DEBUG_KD_SYNC") != nullptr; // synchronise and report after every stage
#define KD_STAGE(name) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(st); fprintf(stderr, "[kd_peer rank %d] %-28s %s\n", r, name, cudaGetErrorString(e_)); } } while (0)
	if (g < 1 || g > 3 || g > t.lt) { set_error("kd_build_peer: %d ranks for %d global levels", world, t.lt); return NBCO_ERR_INVALID; }
	if (t.Lb > t.L) { set_error("kd_build_peer: max_level %d leaves more than %d particles per leaf pair; shallow trees are built on one GPU only", t.L, kBottomCap); return NBCO_ERR_INVALID; }
	TreeGeom tg = geom_of(t);
	PeerKd pk;
	pk.world = world; pk.me = r; pk.g = g;
	for (int q = 0; q < kMaxPeers; ++q)
	{
		unsigned char *base = q < world ? (unsigned char *)pub[q] : nullptr;
		pk.scr[q] = base ? base + kPeerHeader : nullptr;
		pk.pay[q][0] = base ? (float4 *)(base + peer_pay_offset(n, 0)) : nullptr;
		pk.pay[q][1] = base ? (float4 *)(base + peer_pay_offset(n, 1)) : nullptr;
	}
	for (int q = 0; q <= world; ++q) pk.lo[q] = seg_start(n, q, g);
	for (int q = world + 1; q <= kMaxPeers; ++q) pk.lo[q] = n;
	const int64_t lo = pk.lo[r], cnt = pk.lo[r + 1] - lo;
	unsigned char *scr = pk.scr[r];
	float4 *pay[2] = {pk.pay[r][0], pk.pay[r][1]};

	// local histograms of all distributed levels start from zero; local box
	NBCO_CUDA(cudaMemsetAsync(scr + kScrHist, 0, 4 * (size_t)kBins * ((size_t)world - 1), st));
	bbox_init_kernel<<<1, 32, 0, st>>>(reinterpret_cast<u32 *>(scr + kScrBBox)); // (device-side: strictly stream ordered)
	pack_own_kernel<<<grid_for(cnt, 256, ctx->sm_count, 8), 256, 0, st>>>(pos, pay[0], lo, lo + cnt, reinterpret_cast<u32 *>(scr + kScrBBox));
	++ctx->launches;
	KD_STAGE("pack_own");
	NBCO_TRY(peer_barrier(ctx));                       // every local box is published
	root_box_peer_kernel<<<1, 32, 0, st>>>(tg, pk);
	++ctx->launches;
	KD_STAGE("root_box");
	if (dbg)
	{
		for (int q = 0; q < world; ++q)
		{
			u32 bbq[6]; long long rg[2];
			cudaMemcpy(bbq, pk.scr[q] + kScrBBox, sizeof(bbq), cudaMemcpyDeviceToHost);
			cudaMemcpy(rg, pk.scr[q] + kScrRange, sizeof(rg), cudaMemcpyDeviceToHost);
			fprintf(stderr, "[kd_peer rank %d] sees rank %d (scr %p): bbox %08x %08x %08x | %08x %08x %08x  range0 %lld %lld\n", r, q, (void *)pk.scr[q],
			        bbq[0], bbq[1], bbq[2], bbq[3], bbq[4], bbq[5], rg[0], rg[1]);
		}
		float rb[6];
		cudaMemcpy(rb, tg.lbound, 12, cudaMemcpyDeviceToHost); cudaMemcpy(rb + 3, tg.rbound, 12, cudaMemcpyDeviceToHost);
		fprintf(stderr, "[kd_peer rank %d] root box %g %g %g | %g %g %g\n", r, rb[0], rb[1], rb[2], rb[3], rb[4], rb[5]);
	}
	int cur = 0;
	for (int l = 0; l < g; ++l)
	{
		const int nseg = 1 << l;
		// a local segment holds at most cnt records: tiles sized for that (most tiles of a level find an empty range)
		int64_t want = (cnt + (int64_t)ctx->sm_count * 4 - 1) / ((int64_t)ctx->sm_count * 4);
		int64_t tile = ((std::max<int64_t>(want, kChunk) + kChunk - 1) / kChunk) * kChunk;
		const int tps = (int)((cnt + tile - 1) / tile);
		const int64_t *lb = reinterpret_cast<const int64_t *>(scr + kScrRange) + 9 * l;
		u32 *hist = reinterpret_cast<u32 *>(scr + kScrHist) + (size_t)(nseg - 1) * kBins;
		SegState *seg = reinterpret_cast<SegState *>(scr + kScrSeg) + 8 * l;
		top_hist_kernel<<<nseg * tps, kTopThreads, 0, st>>>(pay[cur], tg, hist, n, l, tps, (int)tile, 0, lb, lo);
		++ctx->launches;
		KD_STAGE("hist");
		NBCO_TRY(peer_barrier(ctx));                   // every local histogram of this level is published
		top_pick_peer_kernel<<<nseg, 256, 0, st>>>(pk, n, l);
		KD_STAGE("pick");
		top_partition_kernel<<<nseg * tps, kTopThreads, 0, st>>>(pay[cur], pay[cur ^ 1], tg, seg, n, l, tps, (int)tile, 0, lb, lo);
		ctx->launches += 2;
		KD_STAGE("partition");
		NBCO_TRY(peer_barrier(ctx));                   // every rank's candidates and local minima are in place
		top_select_peer_kernel<<<nseg, kResThreads, 0, st>>>(pk, tg, cur ^ 1, n, l);
		++ctx->launches;
		KD_STAGE("select");
		NBCO_TRY(peer_barrier(ctx));                   // nobody scans this rank's candidates any more: they may be reordered
		top_split_peer_kernel<<<nseg, kResThreads, 0, st>>>(pk, tg, cur ^ 1, n, l);
		++ctx->launches;
		KD_STAGE("split");
		cur ^= 1;
	}
	NBCO_TRY(peer_barrier(ctx));                       // every rank's records are grouped by destination, tables published
	exchange_pull_kernel<<<grid_for(cnt, 256, ctx->sm_count, 8), 256, 0, st>>>(pk, cur, pay[cur ^ 1] + lo);
	++ctx->launches;
	KD_STAGE("exchange");
	cur ^= 1;
	NBCO_TRY(peer_barrier(ctx));                       // nobody reads this rank's records any more: the local levels may overwrite them
	if (t.lt > g) NBCO_CUDA(cudaMemsetAsync(t.hist.p, 0, 4 * (size_t)kBins * ((size_t)1 << (t.lt - 1 - g)), st));
	const int rc = kd_local_levels(ctx, t, pay, cur, g, ev_bottom, r, g);
	KD_STAGE("local levels + bottom");
#undef KD_STAGE
	return rc;
}

} // namespace nbco
