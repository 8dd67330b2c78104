// oracle/ref_harness2d.cu -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" shell around the UNMODIFIED 2D fp64 code path of the reference (SCAL = double, DIM = 2:
// fmm_cart.cuh, direct.cuh, integrator.cuh, appel.cuh), compiled where it lies into
// oracle/_ref/libnbco_ref2d.so.  The reference's own 2D driver main.cu does not compile at HEAD
// (it redefines the globals of constants.cuh), so only the headers are used;
// every number comes out of a reference function.
#define SCAL double
#define DIM 2
#include "kernel.cuh"
#include "direct.cuh"
#include "integrator.cuh"
#include "appel.cuh"
#include "fmm_cart.cuh"
#include "reductions.cuh"

namespace {
typedef void (*eval_fn)(VEC*, VEC*, int, const SCAL*);

// coulombOscillator evaluators of main.cu:52-83 restated with the 2D parameter block {xi/N, 0, kx, ky}
void osc_direct_cpu(VEC *p, VEC *a, int n, const SCAL *param) { direct2_cpu(p, a, n, param); add_elastic_cpu(p, a, n, param + 2); }
void osc_fmm_cpu(VEC *p, VEC *a, int n, const SCAL *param) { fmm_cart_cpu(p, a, n, param); add_elastic_cpu(p, a, n, param + 2); }

eval_fn pick(int which)
{
	switch (which)
	{
		case 0: return direct2_cpu;     // direct.cuh:181
		case 1: return fmm_cart_cpu;    // fmm_cart.cuh:546
		case 2: return osc_direct_cpu;
		case 3: return osc_fmm_cpu;
		default: return nullptr;
	}
}
}

extern "C" {

void ref2_config(int order, double radius, double eps2, double dens, int threads, int coll_)
{
	::fmm_order = order; ::tree_radius = radius; ::EPS2 = eps2; ::dens_inhom = dens; ::CPU_THREADS = threads; ::coll = coll_ != 0;
}

int ref2_eval(int which, double *buf, int n, const double *param)
// buf = [pos | vel | acc], 3*n double2; pos AND vel are permuted into cell order by the FMM (fmm_cart.cuh:644-650)
{
	eval_fn f = pick(which);
	if (!f) return -1;
	compute_force(f, buf, n, param);
	return 0;
}

int ref2_integrate(int scheme, int which, double *buf, int n, const double *param, double dt, int nsteps)
{
	eval_fn f = pick(which);
	if (!f) return -1;
	for (int s = 0; s < nsteps; ++s)
		switch (scheme)
		{
			case 0: symplectic_euler(f, buf, n, param, dt, step_cpu, 1); break;
			case 1: leapfrog(f, buf, n, param, dt, step_cpu, 1); break;
			case 2: forestruth(f, buf, n, param, dt, step_cpu, 1); break;
			case 3: pefrl(f, buf, n, param, dt, step_cpu, 1); break;
			default: return -1;
		}
	return 0;
}

int ref2_levels(int n)
// fmm_cart.cuh:562-564
{
	int order = ::fmm_order;
	SCAL s = order*sqrt((SCAL)order);
	int L = (int)std::round(std::log2(::dens_inhom*(SCAL)n/s)/DIM);
	return std::max(L, 2);
}

} // extern "C"
