// common.cuh -- shared declarations of the sm_100a implementation behind include/nbco.h
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nbco.h"

namespace nbco {

constexpr int kSMs = 148; // B200: 2 dies x 74 SMs; the real count is queried at context creation

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define NBCO_CUDA(expr)                                                              \
	do {                                                                             \
		cudaError_t e__ = (expr);                                                    \
		if (e__ != cudaSuccess) return ::nbco::cuda_fail(e__, #expr, __FILE__, __LINE__); \
	} while (0)

#define NBCO_TRY(expr)                      \
	do {                                    \
		int s__ = (expr);                   \
		if (s__ != NBCO_OK) return s__;     \
	} while (0)

// Growable device buffer owned by a context (replaces the function-local statics of
// fmm_cart3_kdtree.cuh:1480-1498; never shrinks, freed with the context).
struct DevBuf
{
	void *p = nullptr;
	size_t bytes = 0;
	int reserve(size_t need);
	void release();
	template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct FmmPlan; // fmm3.cu

// Multi-GPU over peer memory (peer.cu): what this rank publishes and what it has mapped from the others.
constexpr int kPeerMax = 8;
// published buffer of a rank: [flag words (kPeerHeader) | scratch of the distributed kd build (kPeerScratch) |
//   tree-ordered positions 12 n | velocities 12 n | (16-byte aligned) record buffers A and B, 16 n each]
constexpr size_t kPeerHeader = 1024;
constexpr size_t kPeerScratch = 256 * 1024;
constexpr size_t kPeerData = kPeerHeader + kPeerScratch;
inline size_t peer_pay_offset(int64_t n, int which) { return ((kPeerData + 24 * (size_t)n + 15) & ~(size_t)15) + (size_t)which * 16 * (size_t)n; }
inline size_t peer_pub_bytes(int64_t n) { return peer_pay_offset(n, 2); }
struct PeerState
{
	bool active = false;
	int world = 1, me = 0;
	int64_t n = 0;
	DevBuf pub;                       // [flags | tree-ordered positions 12 n | velocities 12 n]
	void *center[kPeerMax] = {}, *mpole[kPeerMax] = {}, *pubp[kPeerMax] = {}; // [me] = local
	bool opened[kPeerMax] = {};
	unsigned long long epoch = 0;     // barriers passed so far (identical on every rank)
	bool have_full = true;            // the caller's arrays hold ALL positions / velocities (first evaluation)
	float barrier_ms = 0.f;
};
struct Fmm2Plan; // fmm2.cu

} // namespace nbco

struct nbco_ctx
{
	nbco_config cfg;
	int sm_count = nbco::kSMs;
	cudaStream_t stream = nullptr;
	int64_t launches = 0; // kernels launched through this context

	// direct sum scratch
	nbco::DevBuf pos4;      // float4-padded copy of the sources
	nbco::DevBuf dpart;     // partial sums of the source runs (direct sum, split over the sources)
	// diagnostics scratch
	nbco::DevBuf red;       // reduction partials
	// host-call staging
	nbco::DevBuf h_state;   // [pos|vel|acc] for nbco_eval_host / nbco_run_host
	nbco::DevBuf h_param;
	void *pinned = nullptr; size_t pinned_bytes = 0;
	cudaStream_t copy_stream = nullptr;   // nbco_step_host: read-back of the positions overlaps the force evaluation
	cudaEvent_t ev_drift = nullptr, ev_copied = nullptr;

	int32_t *ids = nullptr;   // nbco_track_ids: permuted with pos / vel at every rebuild (unsort = 0)

	nbco::FmmPlan *fmm = nullptr;
	nbco::PeerState peer;
	nbco::Fmm2Plan *fmm2 = nullptr;
};

namespace nbco {

// direct.cu
int direct3_launch(nbco_ctx *ctx, const float *d_pos, float *d_acc, int64_t n, const float *d_param);
int pair_energy_launch(nbco_ctx *ctx, const float *d_pos, int64_t n, double *d_out);
// integrate.cu
int add_elastic_launch(nbco_ctx *ctx, const float *d_pos, float *d_acc, int64_t n, const float *d_k3);
int step_launch(nbco_ctx *ctx, float *d_b, const float *d_a, float ds, int64_t n);
// the same update fused with the kinetic / elastic energy sums of the state it produces (d_out2: two doubles, accumulated)
int step_energy_launch(nbco_ctx *ctx, float *d_b, const float *d_a, float ds, const float *d_other, bool b_is_vel,
                       const float *d_param, int64_t n, double *d_out2);
int kick_drift_launch(nbco_ctx *ctx, float *d_pos, float *d_vel, const float *d_acc, float k1, float k2, bool two, float dt, int64_t n);
int rel_err_launch(nbco_ctx *ctx, const float *d_a, const float *d_ref, int64_t n, double *h_mean, double *h_max);
int kinetic_elastic_launch(nbco_ctx *ctx, const float *d_buf, int64_t n, const float *d_param, double *h_out2);
// fmm3.cu
int fmm3_kd_launch(nbco_ctx *ctx, float *d_pos, float *d_acc, int64_t n, const float *d_param, bool fuse_elastic);
void fmm3_destroy(nbco_ctx *ctx);
int fmm3_harvest(nbco_ctx *ctx, int *overflow); // synchronisation point of the enqueued evaluations (counters, sticky flags, timers)
bool fmm3_next_rebuilds(nbco_ctx *ctx, int64_t n); // will the next FMM evaluation of n particles permute pos / vel?
bool fmm3_rebuilds_in(nbco_ctx *ctx, int64_t n, int64_t k); // ... and the one k evaluations after it?
// peer.cu
int peer_barrier(nbco_ctx *ctx);                                           // all ranks, on the context streams
int peer_publish(nbco_ctx *ctx, const float *d_full, int which, int64_t n);  // own range of pos (0) / vel (1) -> published mirror
int peer_pull(nbco_ctx *ctx, float *d_full, int which, int64_t n);           // the other ranks' ranges <- their mirrors
void peer_release(nbco_ctx *ctx, bool free_exports);
int peer_report_error(nbco_ctx *ctx, unsigned *h_err); // reads and clears the sticky barrier time-out word
int fmm3_peer_buffers(nbco_ctx *ctx, int64_t n, void **center, void **mpole); // fmm3.cu: plans, returns the node arrays
// fmm2.cu (2D fp64 path)
int fmm2_launch(nbco_ctx *ctx, double *d_pos, double *d_acc, int64_t n, const double *d_param, bool fuse_elastic);
int direct2_launch(nbco_ctx *ctx, const double *d_pos, double *d_acc, int64_t n, const double *d_param);
int step2_launch(nbco_ctx *ctx, double *d_b, const double *d_a, double ds, int64_t n);
int kick_drift2_launch(nbco_ctx *ctx, double *d_pos, double *d_vel, const double *d_acc, double kc, double dc, int64_t n);
int add_elastic2_launch(nbco_ctx *ctx, const double *d_pos, double *d_acc, int64_t n, const double *d_k2);
void fmm2_destroy(nbco_ctx *ctx);

inline int grid_for(int64_t work, int block, int sm_count, int per_sm)
{
	int64_t g = (work + block - 1) / block;
	int64_t cap = (int64_t)sm_count * per_sm;
	if (g > cap) g = cap;
	if (g < 1) g = 1;
	return (int)g;
}

} // namespace nbco
