// nbco_shim.cuh -- the reference-side binding of libnbco.so.
//
// A maintainer of coulomb_oscillators adds this file to Simulation/ and includes it from main3.cu after
// "fmm_cart3_kdtree.cuh"; the evaluators below have the reference's plugin signature
//     void f(VEC *p, VEC *a, int n, const SCAL *param)          (Simulation/integrator.cuh:22)
// and can be passed to compute_force / leapfrog / forestruth / pefrl / test_accuracy / test_time unchanged.
// It is compiled for real by oracle/ref_dropin.cu (the unmodified main3.cu translation unit + this shim, linked
// against libnbco.so) and exercised by tests/test_dropin_gpu.py.
// Requires SCAL = float, DIM = 3 (the nbco3 program).  Link with -lnbco.
#pragma once
#include "nbco.h"
#include <cstdlib>
#include <iostream>

static nbco_ctx *g_nbco = nullptr;

static void nbco_sync_config()    // mirror the mutable globals of constants.cuh:36-52 into the context
{
	// nbco_config is passed by pointer across the ABI: a shim compiled against another header version must not run
	if (nbco_abi_version() != NBCO_ABI_VERSION)
	{
		std::cerr << "nbco_shim: libnbco.so has ABI version " << nbco_abi_version() << ", this shim was compiled for " << NBCO_ABI_VERSION << std::endl;
		exit(-1);
	}
	nbco_config c;
	nbco_default_config(&c);
	c.order = ::fmm_order;        c.radius = ::tree_radius;   c.eps2 = ::EPS2;
	c.dens_inhom = ::dens_inhom;  c.max_level = ::tree_L;     c.tree_steps = ::tree_steps;
	c.coll = ::coll;              c.unsort = ::b_unsort;      c.m2l_first = 1;   // GPU traversal order (fmm_cart3_kdtree.cuh:1668)
	if (!g_nbco) { if (nbco_create(&c, &g_nbco)) { std::cerr << nbco_last_error() << std::endl; exit(-1); } }
	else if (nbco_set_config(g_nbco, &c)) { std::cerr << nbco_last_error() << std::endl; exit(-1); }
}

static void nbco_check(int status)
{
	if (status) { std::cerr << nbco_last_error() << std::endl; exit(-1); }   // the reference's gpuErrchk policy (kernel.cuh:52-65)
}

// drop-in replacements with the reference's plugin signature
void fmm_cart3_kdtree_b200(VEC *p, VEC *a, int n, const SCAL *param)          // replaces fmm_cart3_kdtree (fmm_cart3_kdtree.cuh:1478)
{
	nbco_sync_config();
	nbco_check(nbco_force_fmm3_kd(g_nbco, p, a, n, param));
}
void direct3_b200(VEC *p, VEC *a, int n, const SCAL *param)                   // replaces direct3 (direct.cuh:233)
{
	nbco_sync_config();
	nbco_check(nbco_force_direct3(g_nbco, p, a, n, param));
}
void coulombOscillatorFMMKD3_b200(VEC *p, VEC *a, int n, const SCAL *param)   // replaces coulombOscillatorFMMKD3 (main3.cu:59-63)
{
	nbco_sync_config();
	nbco_check(nbco_coulomb_fmm3_kd(g_nbco, p, a, n, param));
}
void coulombOscillatorDirect_b200(VEC *p, VEC *a, int n, const SCAL *param)   // replaces coulombOscillatorDirect (main3.cu:47-51)
{
	nbco_sync_config();
	nbco_check(nbco_coulomb_direct3(g_nbco, p, a, n, param));
}
void step_b200(VEC *b, const VEC *a, SCAL ds, int n)                           // replaces step (kernel.cuh:100)
{
	if (!g_nbco) nbco_sync_config();
	nbco_check(nbco_step(g_nbco, b, a, ds, n));
}
