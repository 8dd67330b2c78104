// peer.cu -- multi-GPU plumbing over NVLink peer memory (one process per GPU, CUDA IPC).
//
// rank r of 2^g owns the subtree of kd node (g, r) = the contiguous tree-order range
// [ceil(n r/w), ceil(n (r+1)/w)) of particles (fmm_cart3_kdtree.cuh:117-118).  Instead of replicating the
// tree and all-gathering the positions every step, every rank PUBLISHES three device buffers
//     centres (float4 / node), multipoles (sM floats / node), [flags | positions | velocities]
// through cudaIpcGetMemHandle; the others map them (cudaIpcOpenMemHandle) and the traversal, M2L, P2P and
// top-level M2M kernels read a remote node or leaf straight from its owner (fmm3_common.cuh: PeerTab).  Only the
// locally essential part of the remote tree ever crosses NVLink, tile by tile, inside the kernels that use
// it.  Ranks synchronise with a flag barrier in peer memory (one small kernel on the context stream, no host
// round trip, no NCCL): twice per evaluation, four times on a tree-rebuild evaluation.
// The 64-byte IPC handles travel between the processes through torch.distributed (parallel.py).

#include "common.cuh"
#include "fmm3_common.cuh"

namespace nbco {

namespace {

struct BarrierArgs { unsigned long long *flags[kPeerMax]; int world, me; };

// every rank writes `epoch` into slot [me] of every rank's flag array, then waits until all slots of its own
// array reached `epoch`.  Bounded spin: a rank that never arrives raises err instead of hanging the GPU.
__global__ void peer_barrier_kernel(BarrierArgs a, unsigned long long epoch, unsigned *err, long long timeout_cycles, unsigned long long *dbg)
{
	const int q = threadIdx.x;
	__threadfence_system();
	if (q < a.world)
	{
		volatile unsigned long long *dst = a.flags[q] + a.me;
		*dst = epoch;
	}
	__threadfence_system();
	if (q < a.world)
	{
		volatile unsigned long long *src = a.flags[a.me] + q;
		const long long t0 = clock64();
		unsigned long long spins = 0;
		while (*src < epoch)
		{
			if (clock64() - t0 > timeout_cycles) { *err = 1u; break; }
			__nanosleep(100);
			++spins;
		}
		if (dbg) { dbg[4 * q] = epoch; dbg[4 * q + 1] = spins; dbg[4 * q + 2] = *src; dbg[4 * q + 3] = (unsigned long long)(clock64() - t0); }
	}
	__threadfence_system();
}

} // namespace

int peer_barrier(nbco_ctx *ctx)
{
	PeerState &ps = ctx->peer;
	if (!ps.active || ps.world == 1) return NBCO_OK;
	BarrierArgs a;
	for (int q = 0; q < kPeerMax; ++q) a.flags[q] = q < ps.world ? (unsigned long long *)ps.pubp[q] : nullptr;
	a.world = ps.world; a.me = ps.me;
	unsigned *err = (unsigned *)((char *)ps.pub.p + 512);
	// Bounded spin.  Ranks may arrive with host-side skew (a snapshot written by one rank, first-call allocations, a GC
	// pause in a Python driver): the default bound is generous (30 s of SM clock at ~2 GHz); NBCO_PEER_TIMEOUT_S overrides.
	// Callers that do rank-local host work between two library calls should meet at a host barrier before re-entering.
	static long long timeout_cycles = 0;
	if (!timeout_cycles)
	{
		const char *e = getenv("NBCO_PEER_TIMEOUT_S");
		double sec = e ? atof(e) : 30.0;
		if (!(sec > 0.0)) sec = 30.0;
		timeout_cycles = (long long)(sec * 2.0e9);
	}
	static const bool debug = getenv("NBCO_DEBUG_KD_SYNC") != nullptr;
	unsigned long long *dbg = debug ? (unsigned long long *)((char *)ps.pub.p + 640) : nullptr; // 4 words per peer, inside the header
	peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(a, ++ps.epoch, err, timeout_cycles, dbg);
	if (debug)
	{
		unsigned long long h[4 * kPeerMax]; unsigned e = 0;
		cudaStreamSynchronize(ctx->stream);
		cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&e, err, 4, cudaMemcpyDeviceToHost);
		for (int q = 0; q < ps.world; ++q)
			fprintf(stderr, "[barrier rank %d] epoch %llu waited on rank %d: spins %llu, flag %llu, cycles %llu, err %u, timeout %lld\n", ps.me,
			        h[4*q], q, h[4*q+1], h[4*q+2], h[4*q+3], e, timeout_cycles);
	}
	++ctx->launches;
	NBCO_CUDA(cudaGetLastError());
	return NBCO_OK;
}

static float *mirror(void *pub, int which, int64_t n) { return (float *)((char *)pub + kPeerData) + (size_t)which * 3 * (size_t)n; }

int peer_publish(nbco_ctx *ctx, const float *d_full, int which, int64_t n)
{
	PeerState &ps = ctx->peer;
	int64_t lo, hi;
	nbco_shard_range(n, ps.me, ps.world, &lo, &hi);
	NBCO_CUDA(cudaMemcpyAsync(mirror(ps.pub.p, which, n) + 3 * lo, d_full + 3 * lo, 12 * (size_t)(hi - lo), cudaMemcpyDeviceToDevice, ctx->stream));
	return NBCO_OK;
}

int peer_pull(nbco_ctx *ctx, float *d_full, int which, int64_t n)
{
	PeerState &ps = ctx->peer;
	for (int k = 1; k < ps.world; ++k)
	{
		const int q = (ps.me + k) % ps.world; // staggered: no two ranks start on the same source
		int64_t lo, hi;
		nbco_shard_range(n, q, ps.world, &lo, &hi);
		NBCO_CUDA(cudaMemcpyAsync(d_full + 3 * lo, mirror(ps.pubp[q], which, n) + 3 * lo, 12 * (size_t)(hi - lo), cudaMemcpyDefault, ctx->stream));
	}
	return NBCO_OK;
}

// CUDA IPC: every importer must close its mapping before the exporter frees the allocation.  Detaching therefore only
// CLOSES this rank's imports; the exported buffers (pub here, centres / multipoles in the FMM plan) stay allocated until
// nbco_destroy, and callers put a host barrier between nbco_peer_detach and nbco_destroy.
void peer_release(nbco_ctx *ctx, bool free_exports)
{
	PeerState &ps = ctx->peer;
	for (int q = 0; q < kPeerMax; ++q)
		if (ps.opened[q])
		{
			cudaIpcCloseMemHandle(ps.center[q]); cudaIpcCloseMemHandle(ps.mpole[q]); cudaIpcCloseMemHandle(ps.pubp[q]);
			ps.opened[q] = false;
		}
	if (free_exports) ps.pub.release();
	ps.active = false;
}

int peer_report_error(nbco_ctx *ctx, unsigned *h_err)
// reads and CLEARS the sticky time-out word of this rank's flag block (the context stays usable after a detach/export cycle)
{
	*h_err = 0;
	PeerState &ps = ctx->peer;
	if (!ps.pub.p) return NBCO_OK;
	unsigned *d = (unsigned *)((char *)ps.pub.p + 512);
	NBCO_CUDA(cudaMemcpyAsync(h_err, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	if (*h_err) NBCO_CUDA(cudaMemsetAsync(d, 0, 4, ctx->stream));
	return NBCO_OK;
}

} // namespace nbco

using namespace nbco;

extern "C" {

int nbco_peer_export(nbco_ctx *ctx, int64_t n, void *h_handles)
{
	if (!ctx || !h_handles) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	const int w = ctx->cfg.world;
	if (w < 1 || w > kPeerMax || (w & (w - 1))) { set_error("peer mode: world size %d is not a power of two <= %d", w, kPeerMax); return NBCO_ERR_INVALID; }
	if (ctx->cfg.unsort) { set_error("peer mode needs unsort = 0 (shards are ranges of the tree order)"); return NBCO_ERR_INVALID; }
	PeerState &ps = ctx->peer;
	if (ps.active) { set_error("peers already attached"); return NBCO_ERR_INVALID; }
	void *center = nullptr, *mpole = nullptr;
	NBCO_TRY(fmm3_peer_buffers(ctx, n, &center, &mpole));
	{
		cudaFuncAttributes fa; // see kd_preload_kernels(): no lazy module load while a barrier kernel spins
		NBCO_CUDA(cudaFuncGetAttributes(&fa, peer_barrier_kernel));
		NBCO_TRY(kd_preload_kernels());
	}
	NBCO_TRY(ps.pub.reserve(peer_pub_bytes(n)));
	NBCO_CUDA(cudaMemsetAsync(ps.pub.p, 0, kPeerData, ctx->stream));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	ps.world = w; ps.me = ctx->cfg.rank; ps.n = n; ps.epoch = 0; ps.have_full = true;
	ps.center[ps.me] = center; ps.mpole[ps.me] = mpole; ps.pubp[ps.me] = ps.pub.p;
	cudaIpcMemHandle_t *h = (cudaIpcMemHandle_t *)h_handles;
	NBCO_CUDA(cudaIpcGetMemHandle(h + 0, center));
	NBCO_CUDA(cudaIpcGetMemHandle(h + 1, mpole));
	NBCO_CUDA(cudaIpcGetMemHandle(h + 2, ps.pub.p));
	return NBCO_OK;
}

int nbco_peer_attach(nbco_ctx *ctx, int32_t peer_rank, const void *h_handles)
{
	if (!ctx || !h_handles) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	PeerState &ps = ctx->peer;
	if (peer_rank < 0 || peer_rank >= ps.world) { set_error("peer rank %d of %d", peer_rank, ps.world); return NBCO_ERR_INVALID; }
	if (peer_rank == ps.me) return NBCO_OK;
	if (!ps.pubp[ps.me]) { set_error("call nbco_peer_export first"); return NBCO_ERR_INVALID; }
	const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)h_handles;
	NBCO_CUDA(cudaIpcOpenMemHandle(&ps.center[peer_rank], h[0], cudaIpcMemLazyEnablePeerAccess));
	NBCO_CUDA(cudaIpcOpenMemHandle(&ps.mpole[peer_rank], h[1], cudaIpcMemLazyEnablePeerAccess));
	NBCO_CUDA(cudaIpcOpenMemHandle(&ps.pubp[peer_rank], h[2], cudaIpcMemLazyEnablePeerAccess));
	ps.opened[peer_rank] = true;
	return NBCO_OK;
}

int nbco_peer_attach_local(nbco_ctx *ctx, int32_t peer_rank, nbco_ctx *other)
// same process: no IPC, the published buffers of `other` are used directly (one process driving several GPUs,
// or several ranks emulated on one device from different host threads -- tests/test_peer_gpu.py)
{
	if (!ctx || !other) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	PeerState &ps = ctx->peer;
	const PeerState &po = other->peer;
	if (peer_rank < 0 || peer_rank >= ps.world || peer_rank == ps.me || po.me != peer_rank || !po.pubp[po.me] || po.n != ps.n || !ps.pubp[ps.me])
	{
		set_error("attach_local: export both contexts first (rank %d of %d, other is rank %d)", peer_rank, ps.world, po.me);
		return NBCO_ERR_INVALID;
	}
	if (other->cfg.device != ctx->cfg.device)
	{
		cudaError_t e = cudaDeviceEnablePeerAccess(other->cfg.device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
		(void)cudaGetLastError();
	}
	ps.center[peer_rank] = po.center[po.me]; ps.mpole[peer_rank] = po.mpole[po.me]; ps.pubp[peer_rank] = po.pubp[po.me];
	ps.opened[peer_rank] = false;
	return NBCO_OK;
}

int nbco_peer_commit(nbco_ctx *ctx)
{
	if (!ctx) { set_error("null context"); return NBCO_ERR_INVALID; }
	PeerState &ps = ctx->peer;
	for (int q = 0; q < ps.world; ++q)
		if (!ps.pubp[q]) { set_error("peer %d not attached", q); return NBCO_ERR_INVALID; }
	ps.active = true;
	return NBCO_OK;
}

int nbco_peer_detach(nbco_ctx *ctx)
{
	if (!ctx) { set_error("null context"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	peer_release(ctx, false);
	DevBuf keep = ctx->peer.pub; // still mapped by the other ranks until they detach too: freed at nbco_destroy
	ctx->peer = PeerState();
	ctx->peer.pub = keep;
	return NBCO_OK;
}

int nbco_peer_barrier(nbco_ctx *ctx)
{
	if (!ctx) { set_error("null context"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	NBCO_TRY(peer_barrier(ctx));
	unsigned err = 0;
	NBCO_TRY(peer_report_error(ctx, &err));
	if (err) { set_error("peer barrier timed out (a rank did not arrive)"); return NBCO_ERR_CUDA; }
	return NBCO_OK;
}

int nbco_peer_gather(nbco_ctx *ctx, void *d_buf, int64_t n)
// leave the full [pos | vel | acc] on every rank (each rank holds its own range): three publish / pull rounds
{
	if (!ctx || !d_buf) { set_error("null argument"); return NBCO_ERR_INVALID; }
	NBCO_CUDA(cudaSetDevice(ctx->cfg.device));
	PeerState &ps = ctx->peer;
	if (!ps.active || ps.n != n) { set_error("peer mode is not active for n = %lld", (long long)n); return NBCO_ERR_INVALID; }
	float *buf = (float *)d_buf;
	NBCO_TRY(peer_barrier(ctx)); // nobody still reads the mirrors
	NBCO_TRY(peer_publish(ctx, buf, 0, n));
	NBCO_TRY(peer_publish(ctx, buf + 3 * n, 1, n));
	NBCO_TRY(peer_barrier(ctx));
	NBCO_TRY(peer_pull(ctx, buf, 0, n));
	NBCO_TRY(peer_pull(ctx, buf + 3 * n, 1, n));
	NBCO_TRY(peer_barrier(ctx));
	NBCO_TRY(peer_publish(ctx, buf + 6 * n, 1, n));
	NBCO_TRY(peer_barrier(ctx));
	NBCO_TRY(peer_pull(ctx, buf + 6 * n, 1, n));
	NBCO_TRY(peer_barrier(ctx));
	// the position mirror keeps serving the P2P kernels of the next evaluation
	ps.have_full = true;
	NBCO_CUDA(cudaStreamSynchronize(ctx->stream));
	return NBCO_OK;
}

} // extern "C"
