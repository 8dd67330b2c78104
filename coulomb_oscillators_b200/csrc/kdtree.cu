// kdtree.cu -- balanced kd-tree build for sm_100a.
//
// Replaces the tree build of fmm_cart3_kdtree (reference Simulation/fmm_cart3_kdtree.cuh:1619-1642):
// minmaxReduce2 (reductions.cuh:67-80), evalRootBox/evalBox (:89-156), evalKeys_kdtree (:158-202) and,
// per level, a full key-value sort (CUB radix at level 0, bb_segsort below, :1311-1364) followed by
// four gather/copy passes over positions and the permutation.
//
// What the tree needs: at every level each node's particles are split at a fixed
// rank along the node's widest axis, the children's boxes are cut at the two boundary particles, and
// only the last level's order is observable.  So:
//   * order = the order a stable sort at every level would give; equivalently a TOTAL order per node:
//     (coordinate on the split axis, coordinates on the previously used distinct axes, most recent
//     first, input index).  With a total order any selection algorithm yields the same partition;
//   * levels whose segments exceed kBottomCap particles ("top"): per segment a 3-pass MSD radix SELECT
//     of the pivot key (11+11+10 bits, shared-memory histograms), then ONE unordered two-way partition
//     of the (u32) ids with block-aggregated cursors -- ~50 B/particle/level, coalesced, instead of a
//     4-pass sort with scattered writes.  Ties on the pivot key (rare) are ranked by the rest of the
//     total order in a separate small kernel;
//   * all remaining levels run in ONE kernel, a CTA per subtree with coordinates resident in shared
//     memory: blocks of >= 512 words use the same select + partition, smaller ones a bitonic sort of
//     64-bit (key, slot) words; children occupy the two halves of their parent's power-of-two block;
//   * particle data moves once, at the end (sorted positions + permutation).

#include "fmm3_common.cuh"

namespace nbco {

namespace {

constexpr int kSelTile = 8192;      // elements per block of the select / partition kernels
constexpr int kSelThreads = 256;
constexpr int kSelPer = kSelTile / kSelThreads;
constexpr int kBins0 = 2048;        // key bits 31..21, then 20..10 (2048 bins), then 9..0 (1024 bins)
constexpr int kSortMax = 256;       // bottom kernel: blocks up to this many words are bitonic-sorted

struct SegSel { u32 prefix, krem, less, eq; };          // radix-select state of one segment
struct SegCur { u32 curL, curR, curE, ties, rmin, pad0, pad1, pad2; }; // partition cursors of one segment

// =====================================================================================
//  bounding box (one pass; min/max are exact in any order)
// =====================================================================================
__global__ void __launch_bounds__(256) bbox_kernel(const float *__restrict__ pos, int64_t n, u32 *__restrict__ out6)
{
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
	{
#pragma unroll
		for (int k = 0; k < 3; ++k)
		{
			float v = pos[3*i+k];
			mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v);
		}
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
//  top levels: radix select + partition per segment
// =====================================================================================
struct TileRange { int64_t a, b, s0; int seg; };

__device__ __forceinline__ TileRange tile_range(int64_t n, int l, int tps, int seg0)
{
	TileRange r;
	r.seg = seg0 + blockIdx.x / tps;
	int t = blockIdx.x % tps;
	r.s0 = seg_start(n, r.seg, l);
	int64_t s1 = seg_start(n, r.seg + 1, l);
	r.a = r.s0 + (int64_t)t * kSelTile;
	r.b = r.a + kSelTile < s1 ? r.a + kSelTile : s1;
	return r;
}

// Shared-memory histogram update of one warp.  Keys of a segment share their high bits, so a whole warp
// often hits ONE bin: that case costs a vote and a single atomic; otherwise plain atomics (few conflicts).
__device__ __forceinline__ void hist_add(u32 *sh, u32 bin, bool valid)
{
	const u32 vmask = __ballot_sync(0xffffffffu, valid);
	if (vmask == 0) return;
	const int leader = __ffs(vmask) - 1;
	const u32 b0 = __shfl_sync(0xffffffffu, bin, leader);
	if (__all_sync(0xffffffffu, !valid || bin == b0))
	{
		if ((int)(threadIdx.x & 31) == leader) atomicAdd(&sh[b0], (u32)__popc(vmask));
	}
	else if (valid) atomicAdd(&sh[bin], 1u);
}

__device__ __forceinline__ void hist_flush(const u32 *sh, u32 *__restrict__ gh, int bins)
{
	for (int b = threadIdx.x; b < bins; b += kSelThreads)
	{
		u32 c = sh[b];
		if (c) atomicAdd(gh + b, c);
	}
}

// x[n] | y[n] | z[n]: a level's keys are gathered from ONE coordinate array per segment (n*4 B, L2-sized)
// instead of 12-byte-strided AoS rows
__global__ void __launch_bounds__(256) to_soa_kernel(const float *__restrict__ pos, float *__restrict__ soa, int64_t n)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
	{
		soa[i] = pos[3*i]; soa[n + i] = pos[3*i+1]; soa[2*n + i] = pos[3*i+2];
	}
}

// keys of level l (evalKeys_kdtree, :158-192) + histogram of key bits 31..21
__global__ void __launch_bounds__(kSelThreads)
keygen_hist_kernel(const float *__restrict__ soa, const int *__restrict__ splitdim, const u32 *__restrict__ idx,
                   u32 *__restrict__ keys, u32 *__restrict__ hist, int64_t n, int l, int tps, int seg0)
{
	__shared__ u32 sh[kBins0];
	for (int b = threadIdx.x; b < kBins0; b += kSelThreads) sh[b] = 0;
	__syncthreads();
	const TileRange r = tile_range(n, l, tps, seg0);
	const int axis = splitdim[kd_beg(l) + r.seg];
	for (int64_t j0 = r.a; j0 < r.b; j0 += kSelThreads)
	{
		const int64_t j = j0 + threadIdx.x;
		const bool valid = j < r.b;
		u32 key = 0;
		if (valid)
		{
			const int64_t id = idx ? (int64_t)idx[j] : j;
			key = ordered_bits(soa[(int64_t)axis * n + id]);
			keys[j] = key;
		}
		hist_add(sh, key >> 21, valid);
	}
	__syncthreads();
	if (r.a < r.b) hist_flush(sh, hist + (int64_t)r.seg * kBins0, kBins0);
}

// histogram of the next digit over the keys that share the prefix selected so far
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
sel_hist_kernel(const u32 *__restrict__ keys, const SegSel *__restrict__ sel, u32 *__restrict__ hist, int64_t n, int l, int tps, int seg0)
{
	constexpr int kHi = PASS == 1 ? 21 : 10, kLo = PASS == 1 ? 10 : 0, kBins = PASS == 1 ? 2048 : 1024;
	__shared__ u32 sh[kBins];
	for (int b = threadIdx.x; b < kBins; b += kSelThreads) sh[b] = 0;
	__syncthreads();
	const TileRange r = tile_range(n, l, tps, seg0);
	const u32 want = sel[r.seg].prefix >> kHi;
	for (int64_t j0 = r.a; j0 < r.b; j0 += kSelThreads)
	{
		const int64_t j = j0 + threadIdx.x;
		u32 key = j < r.b ? keys[j] : 0;
		const bool valid = j < r.b && (key >> kHi) == want;
		hist_add(sh, (key >> kLo) & (kBins - 1), valid);
	}
	__syncthreads();
	if (r.a < r.b) hist_flush(sh, hist + (int64_t)r.seg * kBins0, kBins);
}

// one block per segment: find the bin that holds rank krem, descend into it, clear the histogram
template <int PASS>
__global__ void __launch_bounds__(256)
sel_pick_kernel(SegSel *__restrict__ sel, SegCur *__restrict__ cur, u32 *__restrict__ hist, int64_t n, int l, int seg0)
{
	constexpr int kLo = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0), kBins = PASS == 2 ? 1024 : 2048, kPer = kBins / 256;
	__shared__ u32 wsum[8];
	__shared__ u32 s_found[3];
	const int seg = seg0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	u32 *h = hist + (int64_t)seg * kBins0;
	SegSel st;
	if (PASS == 0)
	{
		// pivot = last particle of the left child: rank (left count - 1) in the segment
		const int64_t k = seg_start(n, 2 * (int64_t)seg + 1, l + 1) - seg_start(n, 2 * (int64_t)seg, l + 1);
		st.prefix = 0; st.krem = (u32)(k - 1); st.less = 0; st.eq = 0;
	}
	else st = sel[seg];
	u32 c[kPer], s = 0;
#pragma unroll
	for (int k = 0; k < kPer; ++k) { c[k] = h[tid * kPer + k]; h[tid * kPer + k] = 0; s += c[k]; }
	u32 incl = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	u32 wbase = 0;
	for (int i = 0; i < w; ++i) wbase += wsum[i];
	const u32 excl = wbase + incl - s;
	if (st.krem >= excl && st.krem < excl + s)
	{
		u32 run = excl;
#pragma unroll
		for (int k = 0; k < kPer; ++k)
		{
			if (st.krem >= run && st.krem < run + c[k]) { s_found[0] = (u32)(tid * kPer + k); s_found[1] = run; s_found[2] = c[k]; }
			run += c[k];
		}
	}
	__syncthreads();
	if (tid == 0)
	{
		st.prefix |= s_found[0] << kLo;
		st.less += s_found[1];
		st.krem -= s_found[1];
		st.eq = s_found[2];
		sel[seg] = st;
		if (PASS == 2)
		{
			SegCur z; z.curL = z.curR = z.curE = z.ties = 0; z.rmin = 0xffffffffu; z.pad0 = z.pad1 = z.pad2 = 0;
			cur[seg] = z;
		}
	}
}

// unordered two-way partition of the ids of every segment around its pivot key
__global__ void __launch_bounds__(kSelThreads)
partition_kernel(const u32 *__restrict__ keys, const u32 *__restrict__ idx_in, u32 *__restrict__ idx_out, u32 *__restrict__ tie,
                 const SegSel *__restrict__ sel, SegCur *__restrict__ cur, int64_t n, int l, int tps, int seg0)
{
	__shared__ u32 wcnt[8][3];
	__shared__ u32 base[3];
	const TileRange r = tile_range(n, l, tps, seg0);
	if (r.a >= r.b) return;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const u32 lt_mask = (1u << lane) - 1u;
	const SegSel st = sel[r.seg];
	const u32 kp = st.prefix, need = st.krem + 1;            // how many of the keys == kp go left
	const bool split_ties = st.eq != need;
	const int64_t kleft = (int64_t)st.less + need;             // size of the left child
	u32 key[kSelPer];
	u32 nl = 0, nr = 0, ne = 0, rmin = 0xffffffffu;
	// a warp owns kSelPer consecutive rows of 32 elements
#pragma unroll
	for (int it = 0; it < kSelPer; ++it)
	{
		const int64_t j = r.a + ((int64_t)w * kSelPer + it) * 32 + lane;
		key[it] = j < r.b ? keys[j] : 0;
		const bool v = j < r.b;
		nl += __popc(__ballot_sync(0xffffffffu, v && key[it] < kp));
		nr += __popc(__ballot_sync(0xffffffffu, v && key[it] > kp));
		ne += __popc(__ballot_sync(0xffffffffu, v && key[it] == kp));
		if (v && key[it] > kp) rmin = min(rmin, key[it]);
	}
	if (lane == 0) { wcnt[w][0] = nl; wcnt[w][1] = nr; wcnt[w][2] = ne; }
	rmin = __reduce_min_sync(0xffffffffu, rmin);
	if (lane == 0 && rmin != 0xffffffffu) atomicMin(&cur[r.seg].rmin, rmin);
	__syncthreads();
	if (tid < 3)
	{
		u32 tot = 0;
		for (int i = 0; i < 8; ++i) tot += wcnt[i][tid];
		u32 *c = tid == 0 ? &cur[r.seg].curL : (tid == 1 ? &cur[r.seg].curR : (split_ties ? &cur[r.seg].ties : &cur[r.seg].curE));
		base[tid] = tot ? atomicAdd(c, tot) : 0;
	}
	__syncthreads();
	u32 ol = base[0], orr = base[1], oe = base[2];
	for (int i = 0; i < w; ++i) { ol += wcnt[i][0]; orr += wcnt[i][1]; oe += wcnt[i][2]; }
#pragma unroll
	for (int it = 0; it < kSelPer; ++it)
	{
		const int64_t j = r.a + ((int64_t)w * kSelPer + it) * 32 + lane;
		const bool v = j < r.b;
		const bool isl = v && key[it] < kp, isr = v && key[it] > kp, ise = v && key[it] == kp;
		const u32 bl = __ballot_sync(0xffffffffu, isl), br = __ballot_sync(0xffffffffu, isr), be = __ballot_sync(0xffffffffu, ise);
		if (v)
		{
			const u32 id = idx_in ? idx_in[j] : (u32)j;
			if (isl) idx_out[r.s0 + ol + __popc(bl & lt_mask)] = id;
			else if (isr) idx_out[r.s0 + kleft + orr + __popc(br & lt_mask)] = id;
			else if (!split_ties) idx_out[r.s0 + st.less + oe + __popc(be & lt_mask)] = id;
			else tie[r.s0 + oe + __popc(be & lt_mask)] = id;
		}
		ol += __popc(bl); orr += __popc(br); oe += __popc(be);
	}
}

// total order among particles whose key on the split axis is equal: previous axes, then input index
__device__ __forceinline__ bool tie_less(const float *__restrict__ pos, u32 a, u32 b, int chain)
{
	for (int c = 1; c < 3; ++c)
	{
		int ax = (chain >> (2 * c)) & 3;
		if (ax == kNoAxis) break;
		u32 ua = ordered_bits(pos[3 * (int64_t)a + ax]), ub = ordered_bits(pos[3 * (int64_t)b + ax]);
		if (ua != ub) return ua < ub;
	}
	return a < b;
}

// segments whose pivot key is shared by particles on both sides: rank the tied ids
__global__ void __launch_bounds__(256)
ties_kernel(const float *__restrict__ pos, const int *__restrict__ chain, const u32 *__restrict__ tie, u32 *__restrict__ idx_out,
            const SegSel *__restrict__ sel, SegCur *__restrict__ cur, int64_t n, int l, int seg0)
{
	const int seg = seg0 + blockIdx.x;
	const SegSel st = sel[seg];
	const u32 need = st.krem + 1;
	if (st.eq == need) return;
	const int64_t s0 = seg_start(n, seg, l), cnt = seg_start(n, seg + 1, l) - s0;
	const int ch = chain[kd_beg(l) + seg];
	const u32 t = st.eq, greater = (u32)cnt - st.less - st.eq;
	const u32 *ties = tie + s0;
	for (u32 e = threadIdx.x; e < t; e += blockDim.x)
	{
		const u32 id = ties[e];
		u32 rank = 0;
		for (u32 f = 0; f < t; ++f) rank += tie_less(pos, ties[f], id, ch) ? 1u : 0u;
		if (rank < need) idx_out[s0 + st.less + rank] = id;
		else idx_out[s0 + st.less + need + greater + (rank - need)] = id;
	}
	if (threadIdx.x == 0) atomicMin(&cur[seg].rmin, st.prefix); // a tied key also starts the right child
}

// boxes of level l+1 from the pivots of level l (evalBox_krnl, :109-137)
__global__ void __launch_bounds__(256)
evalbox_top_kernel(TreeGeom g, const SegSel *__restrict__ sel, const SegCur *__restrict__ cur, int l, int seg0, int nseg)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nseg) return;
	const int seg = seg0 + k;
	const int node = kd_beg(l) + seg, axis = g.splitdim[node], pch = g.chain[node];
	float lb[3], rb[3];
	for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
	const float save = rb[axis];
	rb[axis] = unordered_bits(sel[seg].prefix);   // coordinate of the left child's last particle
	write_box(g, 2*node + 1, lb, rb, pch);
	rb[axis] = save;
	lb[axis] = unordered_bits(cur[seg].rmin);     // coordinate of the right child's first particle
	write_box(g, 2*node + 2, lb, rb, pch);
}

// =====================================================================================
//  bottom levels: one CTA per level-lt node, particles resident in shared memory
// =====================================================================================
struct BottomSmem
{
	float *sx, *sy, *sz;
	u64 *comp;        // (key << 13 | slot) words, ~0 = empty
	u32 *hist;        // [<=16 blocks][256] select histograms; reused as u16 tie lists
	SegSel *sel;      // per block
	u32 *cur;         // per block: curL, curR, curE, ties, rmin (5 words)
};

__device__ __forceinline__ float slot_coord(const BottomSmem &s, int axis, u32 slot)
{
	return axis == 0 ? s.sx[slot] : (axis == 1 ? s.sy[slot] : s.sz[slot]);
}

// rest of the total order for two slots with equal split-axis keys
__device__ __forceinline__ bool slot_tie_less(const BottomSmem &s, u32 sa, u32 sb, int chain, const u32 *__restrict__ idx_in, int64_t s0)
{
	for (int c = 1; c < 3; ++c)
	{
		int ax = (chain >> (2 * c)) & 3;
		if (ax == kNoAxis) break;
		u32 ua = ordered_bits(slot_coord(s, ax, sa)), ub = ordered_bits(slot_coord(s, ax, sb));
		if (ua != ub) return ua < ub;
	}
	u32 ia = idx_in ? idx_in[s0 + sa] : sa, ib = idx_in ? idx_in[s0 + sb] : sb;
	return ia < ib;
}

__device__ __forceinline__ bool comp_less(u64 a, u64 b, const BottomSmem &s, int chain, const u32 *__restrict__ idx_in, int64_t s0)
{
	u64 ka = a >> 13, kb = b >> 13;
	if (ka != kb) return ka < kb;
	if (a == ~0ull) return false;
	return slot_tie_less(s, (u32)a & kSlotMask, (u32)b & kSlotMask, chain, idx_in, s0);
}


// ---- warp-local bitonic sort of the 256 words a warp owns (blocks of B <= 256 words) ----
// lane holds words e = i*32 + lane (i = 0..7) in registers: partners at distance < 32 come by shuffle,
// larger distances are other registers of the same lane; no shared memory, no block barrier.
struct WarpSortCtx
{
	const BottomSmem *s; const int *chain_arr; const u32 *idx_in; int64_t s0; int node0; int base; int logB;
};

// Words of different keys order like plain u64 (the key sits above the slot).  Equal keys with different
// slots are rare: the slow path (rest of the total order) runs only when some lane of the warp sees one.
__device__ __forceinline__ bool ws_first_less(u64 a, u64 b, int e, const WarpSortCtx &c)
{
	bool less = a < b;
	const bool tie = ((a ^ b) >> 13) == 0 && a != b;
	if (__any_sync(0xffffffffu, tie))
	{
		if (tie)
		{
			const int chain = c.chain_arr[c.node0 + ((c.base + e) >> c.logB)];
			less = slot_tie_less(*c.s, (u32)a & kSlotMask, (u32)b & kSlotMask, chain, c.idx_in, c.s0);
		}
	}
	return less;
}

template <int M>
__device__ __forceinline__ void ws_reg_stage(u64 (&v)[8], int lane, int k, int B, const WarpSortCtx &c)
{
#pragma unroll
	for (int i = 0; i < 8; ++i)
		if ((i & M) == 0)
		{
			const int e = i * 32 + lane;
			const bool asc = (e & k) == 0 || k == B;
			const u64 a = v[i], b = v[i | M];
			const bool b_first = ws_first_less(b, a, e, c);
			const bool sw = (asc == b_first) && a != b;
			v[i] = sw ? b : a;
			v[i | M] = sw ? a : b;
		}
}

__device__ __forceinline__ void warp_sort_blocks(u64 (&v)[8], int lane, int B, const WarpSortCtx &c)
{
	for (int k = 2; k <= B; k <<= 1)
		for (int jj = k >> 1; jj > 0; jj >>= 1)
		{
			if (jj >= 32)
			{
				if (jj == 32) ws_reg_stage<1>(v, lane, k, B, c);
				else if (jj == 64) ws_reg_stage<2>(v, lane, k, B, c);
				else ws_reg_stage<4>(v, lane, k, B, c);
			}
			else
			{
				const bool lower = (lane & jj) == 0;
#pragma unroll
				for (int i = 0; i < 8; ++i)
				{
					const int e = i * 32 + lane;
					const u64 a = v[i];
					const u64 o = __shfl_xor_sync(0xffffffffu, a, jj);
					const bool asc = (e & k) == 0 || k == B;
					const bool o_first = ws_first_less(o, a, e, c);
					// the lower index keeps the smaller word when ascending
					const bool take = ((lower == asc) == o_first) && o != a;
					v[i] = take ? o : a;
				}
			}
		}
}

__global__ void __launch_bounds__(kBottomThreads, 1)
kd_bottom_kernel(TreeGeom g, const float *__restrict__ pos, const u32 *__restrict__ idx_in,
                 float *__restrict__ spos, int *__restrict__ perm, int64_t n, int lt, int L, int P2, int blk0)
{
	extern __shared__ unsigned char smem_raw[];
	BottomSmem s;
	s.comp = reinterpret_cast<u64 *>(smem_raw);
	s.sx = reinterpret_cast<float *>(smem_raw + sizeof(u64) * kBottomCap);
	s.sy = s.sx + kBottomCap;
	s.sz = s.sy + kBottomCap;
	s.hist = reinterpret_cast<u32 *>(s.sz + kBottomCap);
	s.sel = reinterpret_cast<SegSel *>(s.hist + 16 * 256);
	s.cur = reinterpret_cast<u32 *>(s.sel + 16);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int b = blk0 + blockIdx.x;
	const int64_t s0 = seg_start(n, b, lt);
	const int c0 = (int)(seg_start(n, b + 1, lt) - s0);
	constexpr int kPer = kBottomCap / kBottomThreads;

	for (int t = tid; t < c0; t += kBottomThreads)
	{
		int64_t id = idx_in ? (int64_t)idx_in[s0 + t] : s0 + t;
		s.sx[t] = pos[3*id]; s.sy[t] = pos[3*id+1]; s.sz[t] = pos[3*id+2];
	}
	__syncthreads();

	const int nlev = L - lt; // levels lt .. L-1 are split here
	for (int j = 0; j < nlev; ++j)
	{
		const int l = lt + j;
		const int B = P2 >> j, logB = 31 - __clz(B), nblk = 1 << j;
		// (a) words (key << 13 | slot) of every block, padded with ~0
		for (int p = tid; p < P2; p += kBottomThreads)
		{
			int q = p >> logB, t = p & (B - 1);
			int64_t i = ((int64_t)b << j) + q;
			int cnt = (int)(seg_start(n, i + 1, l) - seg_start(n, i, l));
			u64 c = ~0ull;
			if (t < cnt)
			{
				u32 slot = (j == 0) ? (u32)t : ((u32)s.comp[p] & kSlotMask);
				int axis = g.splitdim[kd_beg(l) + (int)i];
				c = ((u64)ordered_bits(slot_coord(s, axis, slot)) << 13) | slot;
			}
			s.comp[p] = c;
		}
		__syncthreads();

		const bool last = j + 1 == nlev;
		if (B > kSortMax && !last)
		{
			// ---- radix select of the pivot key of every block (4 x 8 bits, MSD first) ----
			if (tid < nblk)
			{
				int64_t i = ((int64_t)b << j) + tid;
				int kl = (int)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
				SegSel z; z.prefix = 0; z.krem = (u32)(kl - 1); z.less = 0; z.eq = 0;
				s.sel[tid] = z;
			}
			for (int pass = 0; pass < 4; ++pass)
			{
				const int lo = 24 - 8 * pass;
				for (int i = tid; i < nblk * 256; i += kBottomThreads) s.hist[i] = 0;
				__syncthreads();
#pragma unroll
				for (int e = 0; e < kPer; ++e)
				{
					const int p = tid + e * kBottomThreads;
					bool valid = false;
					u32 bin = 0;
					if (p < P2)
					{
						const u64 c = s.comp[p];
						const int q = p >> logB;
						const u32 key = (u32)(c >> 13);
						valid = c != ~0ull && (pass == 0 || (key >> (lo + 8)) == (s.sel[q].prefix >> (lo + 8)));
						bin = (u32)q * 256u + ((key >> lo) & 255u);
					}
					hist_add(s.hist, bin, valid);
				}
				__syncthreads();
				if (warp < nblk)
				{
					// a warp picks the bin of its block that holds rank krem
					const u32 *h = s.hist + warp * 256;
					SegSel st = s.sel[warp];
					u32 c[8], sum = 0;
#pragma unroll
					for (int k = 0; k < 8; ++k) { c[k] = h[lane * 8 + k]; sum += c[k]; }
					u32 incl = sum;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
					const u32 excl = incl - sum;
					const bool mine = st.krem >= excl && st.krem < excl + sum;
					u32 fb = 0, fr = 0, fc = 0;
					if (mine)
					{
						u32 run = excl;
#pragma unroll
						for (int k = 0; k < 8; ++k)
						{
							if (st.krem >= run && st.krem < run + c[k]) { fb = lane * 8 + k; fr = run; fc = c[k]; }
							run += c[k];
						}
					}
					const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
					fb = __shfl_sync(0xffffffffu, fb, src); fr = __shfl_sync(0xffffffffu, fr, src); fc = __shfl_sync(0xffffffffu, fc, src);
					if (lane == 0)
					{
						st.prefix |= fb << lo; st.less += fr; st.krem -= fr; st.eq = fc;
						s.sel[warp] = st;
					}
				}
				__syncthreads();
			}
			if (tid < nblk)
			{
				u32 *c = s.cur + 5 * tid;
				c[0] = c[1] = c[2] = c[3] = 0; c[4] = 0xffffffffu;
			}
			// ---- partition straight into the children's half blocks ----
			u64 v[kPer];
			u32 dst[kPer];
			u16 *tl = reinterpret_cast<u16 *>(s.hist); // tie lists, one region of B entries per block
			__syncthreads();
#pragma unroll
			for (int e = 0; e < kPer; ++e)
			{
				// the 32 positions of a warp lie in one block (B >= 512): one cursor atomic per side per warp
				const int p = tid + e * kBottomThreads;
				v[e] = p < P2 ? s.comp[p] : ~0ull;
				dst[e] = 0xffffffffu;
				const bool valid = v[e] != ~0ull;
				const int q = (p < P2 ? p : P2 - 1) >> logB;
				const SegSel st = s.sel[q];
				const u32 key = (u32)(v[e] >> 13), need = st.krem + 1;
				const bool isl = valid && key < st.prefix, isr = valid && key > st.prefix, ise = valid && key == st.prefix;
				const u32 bl = __ballot_sync(0xffffffffu, isl), br = __ballot_sync(0xffffffffu, isr), be = __ballot_sync(0xffffffffu, ise);
				u32 *c = s.cur + 5 * q;
				const bool split_ties = st.eq != need;
				u32 ol = 0, orr = 0, oe = 0;
				if (lane == 0)
				{
					if (bl) ol = atomicAdd(&c[0], (u32)__popc(bl));
					if (br) orr = atomicAdd(&c[1], (u32)__popc(br));
					if (be) oe = atomicAdd(split_ties ? &c[3] : &c[2], (u32)__popc(be));
				}
				ol = __shfl_sync(0xffffffffu, ol, 0); orr = __shfl_sync(0xffffffffu, orr, 0); oe = __shfl_sync(0xffffffffu, oe, 0);
				const u32 lt_mask = (1u << lane) - 1u;
				const u32 rmin = __reduce_min_sync(0xffffffffu, isr ? key : 0xffffffffu);
				if (lane == 0 && br) atomicMin(&c[4], rmin);
				if (isl) dst[e] = q * B + ol + __popc(bl & lt_mask);
				else if (isr) dst[e] = q * B + (B >> 1) + orr + __popc(br & lt_mask);
				else if (ise && !split_ties) dst[e] = q * B + st.less + oe + __popc(be & lt_mask);
				else if (ise) tl[q * B + oe + __popc(be & lt_mask)] = (u16)((u32)v[e] & kSlotMask); // ranked below
			}
			__syncthreads();
			for (int p = tid; p < P2; p += kBottomThreads) s.comp[p] = ~0ull;
			__syncthreads();
#pragma unroll
			for (int e = 0; e < kPer; ++e)
				if (dst[e] != 0xffffffffu) s.comp[dst[e]] = v[e];
			// tied pivots: rank the tied slots of every block by the rest of the total order
			for (int q = 0; q < nblk; ++q)
			{
				const SegSel st = s.sel[q];
				const u32 need = st.krem + 1;
				if (st.eq == need) continue;
				int64_t i = ((int64_t)b << j) + q;
				const int cnt = (int)(seg_start(n, i + 1, l) - seg_start(n, i, l));
				const int ch = g.chain[kd_beg(l) + (int)i];
				const u32 greater = (u32)cnt - st.less - st.eq;
				for (u32 e = tid; e < st.eq; e += kBottomThreads)
				{
					const u32 slot = tl[q * B + e];
					u32 rank = 0;
					for (u32 f = 0; f < st.eq; ++f) rank += slot_tie_less(s, tl[q * B + f], slot, ch, idx_in, s0) ? 1u : 0u;
					const u64 w = ((u64)st.prefix << 13) | slot;
					if (rank < need) s.comp[q * B + st.less + rank] = w;
					else s.comp[q * B + (B >> 1) + greater + (rank - need)] = w;
				}
				if (tid == 0) atomicMin(&s.cur[5 * q + 4], st.prefix);
			}
			__syncthreads();
			// boxes of the children
			if (tid < nblk)
			{
				const int q = tid;
				int64_t i = ((int64_t)b << j) + q;
				int node = kd_beg(l) + (int)i;
				int axis = g.splitdim[node], pch = g.chain[node];
				float lb[3], rb[3];
				for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
				const float save = rb[axis];
				rb[axis] = unordered_bits(s.sel[q].prefix);
				write_box(g, 2*node + 1, lb, rb, pch);
				rb[axis] = save; lb[axis] = unordered_bits(s.cur[5 * q + 4]);
				write_box(g, 2*node + 2, lb, rb, pch);
			}
			__syncthreads();
			continue;
		}

		// ---- small blocks (and the last level): bitonic sort inside every block of B words ----
		if (B <= 256)
		{
			if (warp * 256 < P2)
			{
				WarpSortCtx wc{&s, g.chain, idx_in, s0, kd_beg(l) + (int)((int64_t)b << j), warp * 256, logB};
				u64 v[8];
#pragma unroll
				for (int i = 0; i < 8; ++i) v[i] = s.comp[warp * 256 + i * 32 + lane];
				warp_sort_blocks(v, lane, B, wc);
#pragma unroll
				for (int i = 0; i < 8; ++i) s.comp[warp * 256 + i * 32 + lane] = v[i];
			}
			__syncthreads();
		}
		else
		for (int k = 2; k <= B; k <<= 1)
			for (int jj = k >> 1; jj > 0; jj >>= 1)
			{
				for (int t = tid; t < (P2 >> 1); t += kBottomThreads)
				{
					int lo = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
					int hi = lo | jj;
					bool asc = (lo & k) == 0 || k == B;
					u64 a = s.comp[lo], c = s.comp[hi];
					int chain = 0;
					bool tie = (a >> 13) == (c >> 13) && a != ~0ull;
					if (tie) chain = g.chain[kd_beg(l) + (int)(((int64_t)b << j) + (lo >> logB))];
					bool sw = asc ? comp_less(c, a, s, chain, idx_in, s0) : comp_less(a, c, s, chain, idx_in, s0);
					if (sw) { s.comp[lo] = c; s.comp[hi] = a; }
				}
				__syncthreads();
			}
		// boxes of the children (evalBox_krnl for level l+1)
		for (int q = tid; q < nblk; q += kBottomThreads)
		{
			int64_t i = ((int64_t)b << j) + q;
			int node = kd_beg(l) + (int)i;
			int kl = (int)(seg_start(n, 2*i + 1, l + 1) - seg_start(n, 2*i, l + 1));
			int axis = g.splitdim[node], pch = g.chain[node];
			float lb[3], rb[3];
			for (int k = 0; k < 3; ++k) { lb[k] = g.lbound[3*node+k]; rb[k] = g.rbound[3*node+k]; }
			float cl = slot_coord(s, axis, (u32)s.comp[q * B + kl - 1] & kSlotMask);
			float cr = slot_coord(s, axis, (u32)s.comp[q * B + kl] & kSlotMask);
			float save = rb[axis];
			rb[axis] = cl;
			write_box(g, 2*node + 1, lb, rb, pch);
			rb[axis] = save; lb[axis] = cr;
			write_box(g, 2*node + 2, lb, rb, pch);
		}
		if (!last)
		{
			// move every right child to the start of the second half of its parent's block
			u64 v[kPer];
			const int Bh = B >> 1, logBh = logB - 1;
#pragma unroll
			for (int e = 0; e < kPer; ++e)
			{
				int p = tid + e * kBottomThreads;
				v[e] = ~0ull;
				if (p < P2)
				{
					int q2 = p >> logBh, t = p & (Bh - 1), q = q2 >> 1;
					int64_t i2 = ((int64_t)b << (j + 1)) + q2;
					int cnt = (int)(seg_start(n, i2 + 1, l + 1) - seg_start(n, i2, l + 1));
					int kl = (int)(seg_start(n, (i2 | 1), l + 1) - seg_start(n, (i2 & ~1ll), l + 1));
					if (t < cnt) v[e] = s.comp[q * B + ((q2 & 1) ? kl + t : t)];
				}
			}
			__syncthreads();
#pragma unroll
			for (int e = 0; e < kPer; ++e)
			{
				int p = tid + e * kBottomThreads;
				if (p < P2) s.comp[p] = v[e];
			}
		}
		__syncthreads();
	}
	// output: storage order = order after the level-(L-1) sort
	{
		const int j = nlev - 1, l = L - 1;
		const int B = P2 >> j, logB = 31 - __clz(B);
		for (int p = tid; p < P2; p += kBottomThreads)
		{
			int q = p >> logB, t = p & (B - 1);
			int64_t i = ((int64_t)b << j) + q;
			int64_t st = seg_start(n, i, l);
			int cnt = (int)(seg_start(n, i + 1, l) - st);
			if (t < cnt)
			{
				u32 slot = (u32)s.comp[p] & kSlotMask;
				int64_t dst = st + t;
				perm[dst] = idx_in ? (int)idx_in[s0 + slot] : (int)(s0 + slot);
				spos[3*dst] = s.sx[slot]; spos[3*dst+1] = s.sy[slot]; spos[3*dst+2] = s.sz[slot];
			}
		}
	}
}

constexpr size_t kBottomSmemBytes = (sizeof(u64) + 3 * sizeof(float)) * kBottomCap + 16 * 256 * sizeof(u32)
                                    + 16 * sizeof(SegSel) + 16 * 5 * sizeof(u32);

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
	NBCO_TRY(t.keys.reserve(4 * (size_t)n)); NBCO_TRY(t.tie.reserve(4 * (size_t)n));
	NBCO_TRY(t.idxA.reserve(4 * (size_t)n)); NBCO_TRY(t.idxB.reserve(4 * (size_t)n));
	NBCO_TRY(t.spos.reserve(12 * (size_t)n)); NBCO_TRY(t.perm.reserve(4 * (size_t)n));
	const size_t nseg = (size_t)1 << std::max(t.lt - 1, 0);
	NBCO_TRY(t.hist.reserve(4 * (size_t)kBins0 * nseg));
	NBCO_TRY(t.sel.reserve(sizeof(SegSel) * nseg)); NBCO_TRY(t.cur.reserve(sizeof(SegCur) * nseg));
	NBCO_TRY(t.bbox.reserve(64));
	if (t.lt > 0) NBCO_TRY(t.soa.reserve(12 * (size_t)n));
	return NBCO_OK;
}

void kd_release(KdTree &t)
{
	DevBuf *all[] = {&t.lbound, &t.rbound, &t.size2, &t.splitdim, &t.chain, &t.keys, &t.idxA, &t.idxB, &t.tie,
	                 &t.hist, &t.sel, &t.cur, &t.spos, &t.perm, &t.bbox, &t.soa};
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
	bbox_kernel<<<grid_for(n, 256, ctx->sm_count, 4), 256, 0, st>>>(pos, n, bb);
	root_box_kernel<<<1, 32, 0, st>>>(tg, bb);
	ctx->launches += 2;

	u32 *keys = t.keys.as<u32>(), *tie = t.tie.as<u32>(), *hist = t.hist.as<u32>();
	SegSel *sel = t.sel.as<SegSel>();
	SegCur *cur = t.cur.as<SegCur>();
	u32 *ibuf[2] = {t.idxA.as<u32>(), t.idxB.as<u32>()};
	const int ltop = t.lt; // levels [0, ltop) are partitioned globally
	float *soa = t.soa.as<float>();
	if (ltop > 0)
	{
		NBCO_CUDA(cudaMemsetAsync(hist, 0, 4 * (size_t)kBins0 * ((size_t)1 << (ltop - 1)), st));
		to_soa_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, st>>>(pos, soa, n);
		++ctx->launches;
	}
	const u32 *iin = nullptr; // level 0 reads the identity
	for (int l = 0; l < ltop; ++l)
	{
		// below level g only the segments of rank r's subtree (multi-GPU: the other subtrees are built by their owners)
		const int nseg = l >= g ? 1 << (l - g) : 1 << l, seg0 = l >= g ? r << (l - g) : 0;
		const int64_t maxseg = ((n - 1) >> l) + 1;
		const int tps = (int)((maxseg + kSelTile - 1) / kSelTile);
		const int tiles = nseg * tps;
		u32 *iout = ibuf[l & 1];
		keygen_hist_kernel<<<tiles, kSelThreads, 0, st>>>(soa, tg.splitdim, iin, keys, hist, n, l, tps, seg0);
		sel_pick_kernel<0><<<nseg, 256, 0, st>>>(sel, cur, hist, n, l, seg0);
		sel_hist_kernel<1><<<tiles, kSelThreads, 0, st>>>(keys, sel, hist, n, l, tps, seg0);
		sel_pick_kernel<1><<<nseg, 256, 0, st>>>(sel, cur, hist, n, l, seg0);
		sel_hist_kernel<2><<<tiles, kSelThreads, 0, st>>>(keys, sel, hist, n, l, tps, seg0);
		sel_pick_kernel<2><<<nseg, 256, 0, st>>>(sel, cur, hist, n, l, seg0);
		partition_kernel<<<tiles, kSelThreads, 0, st>>>(keys, iin, iout, tie, sel, cur, n, l, tps, seg0);
		ties_kernel<<<nseg, 256, 0, st>>>(pos, tg.chain, tie, iout, sel, cur, n, l, seg0);
		evalbox_top_kernel<<<(nseg + 255) / 256, 256, 0, st>>>(tg, sel, cur, l, seg0, nseg);
		ctx->launches += 9;
		iin = iout;
	}
	if (ev_bottom) NBCO_CUDA(cudaEventRecord(ev_bottom, st));
	if (!t.bottom_attr)
	{
		NBCO_CUDA(cudaFuncSetAttribute(kd_bottom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBottomSmemBytes));
		t.bottom_attr = true;
	}
	int64_t maxseg = ((n - 1) >> ltop) + 1;
	int P2 = 2; while (P2 < maxseg) P2 <<= 1;
	while ((P2 >> (t.L - 1 - ltop)) < 2) P2 <<= 1; // the last level sorts blocks of at least 2 words
	if (P2 > kBottomCap) { set_error("internal: bottom block %d", P2); return NBCO_ERR_INVALID; }
	if (g > ltop) { set_error("more ranks than shared-memory kd blocks (2^%d > 2^%d)", g, ltop); return NBCO_ERR_INVALID; }
	kd_bottom_kernel<<<1 << (ltop - g), kBottomThreads, kBottomSmemBytes, st>>>(tg, pos, iin, t.spos.as<float>(), t.perm.as<int>(),
	                                                                            n, ltop, t.L, P2, r << (ltop - g));
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

} // namespace nbco
