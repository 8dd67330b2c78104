// oracle/ref_harness_f64.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The UNMODIFIED reference 3D translation unit (main3.cu) compiled with SCAL = double (constants.cuh:22-24 honours a
// predefined SCAL): the fp64 ground truth of the SAME algorithm (same kd-tree rule -- the sort
// keys stay fp32, fmm_cart3_kdtree.cuh:158-165 --, same MAC, same operators).  tests/test_fmm_gpu.py uses it to settle the
// max-norm tolerance: max rel_diff1(ours fp32, fp64) must not exceed max rel_diff1(reference fp32, fp64).
// Built by oracle/Makefile into oracle/_ref/libnbco_ref_f64.so (about five minutes).
#define SCAL double
#define main nbco_ref64_cli_main
#include "main3.cu"
#undef main

extern "C" {

void ref64_config(int order, double radius, double eps2, double dens, int threads, int coll_)
{
	::fmm_order = order;
	::tree_radius = radius;
	::EPS2 = eps2;
	::dens_inhom = dens;
	::CPU_THREADS = threads;
	::coll = coll_ != 0;
	::b_unsort = true;   // accelerations come back in input order
	::tree_steps = 1;
}

// which: 0 direct3_cpu (direct.cuh:247), 1 fmm_cart3_kdtree_cpu (fmm_cart3_kdtree.cuh:1773); pos, acc: n x 3 doubles
int ref64_eval(int which, const double *pos, double *acc, int n, const double *param)
{
	std::vector<VEC> buf(3 * (size_t)n); // [pos | vel | acc], the layout the evaluators expect (integrator.cuh:24)
	for (int i = 0; i < n; ++i) buf[i] = VEC{pos[3*i], pos[3*i+1], pos[3*i+2]};
	if (which == 0) direct3_cpu(buf.data(), buf.data() + 2 * (size_t)n, n, param);
	else if (which == 1) fmm_cart3_kdtree_cpu(buf.data(), buf.data() + 2 * (size_t)n, n, param);
	else return -1;
	for (int i = 0; i < n; ++i)
	{
		const VEC a = buf[2 * (size_t)n + i];
		acc[3*i] = a.x; acc[3*i+1] = a.y; acc[3*i+2] = a.z;
	}
	return 0;
}

} // extern "C"
