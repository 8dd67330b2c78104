// Host instantiation of coulomb_oscillators_b200/csrc/fmm_ops.cuh (the same templates the CUDA
// kernels use) behind a runtime-order switch, so that tests can compare them with oracle/ on the CPU.
#include "../coulomb_oscillators_b200/csrc/fmm_ops.cuh"
using namespace nbco::ops;
#define SWITCH_P(CALL) switch (p) { case 1: { constexpr int P = 1; CALL; } break; case 2: { constexpr int P = 2; CALL; } break; \
	case 3: { constexpr int P = 3; CALL; } break; case 4: { constexpr int P = 4; CALL; } break; case 5: { constexpr int P = 5; CALL; } break; \
	case 6: { constexpr int P = 6; CALL; } break; case 7: { constexpr int P = 7; CALL; } break; case 8: { constexpr int P = 8; CALL; } break; default: return -1; }
extern "C" {
int ops_p2m(float *M, int p, const float *d) { SWITCH_P(p2m_acc<P>(M, d[0], d[1], d[2])); return 0; }
int ops_m2m(float *Mo, const float *Mi, int p, const float *d) { SWITCH_P(m2m_acc<P>(Mo, Mi, d[0], d[1], d[2])); return 0; }
int ops_m2l(float *L, const float *M, int p, const float *u, float r) { SWITCH_P(m2l_acc<P>(L, M, u[0], u[1], u[2], 1.f / r)); return 0; }
int ops_l2l(float *Lc, const float *Lp, int p, const float *d)
{ SWITCH_P(float S[sym_off(P + 1)]; S[0] = 0; local_expand<P>(S, Lp); l2l_acc<P>(Lc, S, d[0], d[1], d[2])); return 0; }
int ops_l2p(float *f, const float *L, int p, const float *d)
{ SWITCH_P(float S[sym_off(P + 1)]; S[0] = 0; local_expand<P>(S, L); l2p_field<P>(f, S, d[0], d[1], d[2])); return 0; }
}
