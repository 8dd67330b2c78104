// oracle/ref_dropin.cu -- TEST INFRASTRUCTURE ONLY: proves the drop-in by compiling it.
//
// The UNMODIFIED reference translation unit (main3.cu, its main() renamed) + include/nbco_shim.cuh (the binding a
// maintainer adds), linked against coulomb_oscillators_b200/libnbco.so.  The entry points
// below run the REFERENCE's own host code -- compute_force / leapfrog (integrator.cuh:22-28,68-96) and test_accuracy
// (main3.cu:139-182) -- over this repo's evaluators.  Built by oracle/Makefile into oracle/_ref/libnbco_dropin.so.
#define main nbco_ref_cli_main
#include "main3.cu"
#undef main
#include "nbco_shim.cuh"

extern "C" {

void dropin_config(int order, float radius, float eps2, float dens, int coll_, int unsort, int tsteps)
{
	::fmm_order = order; ::tree_radius = radius; ::EPS2 = eps2; ::dens_inhom = dens;
	::coll = coll_ != 0; ::b_unsort = unsort != 0; ::tree_steps = tsteps;
}

// main3.cu:835-846 with the shim's evaluators: a = f(x), then `steps` calls of the reference's leapfrog.
// buf = [pos | vel | acc] on the host (9 n floats), read back at the end.
int dropin_leapfrog(float *buf, int n, const float *param6, double dt, int steps)
{
	SCAL *d_buf, *d_par;
	if (cudaMalloc(&d_buf, sizeof(VEC) * 3 * (size_t)n) != cudaSuccess) return -1;
	if (cudaMalloc(&d_par, sizeof(SCAL) * 6) != cudaSuccess) return -1;
	cudaMemcpy(d_buf, buf, sizeof(VEC) * 2 * (size_t)n, cudaMemcpyHostToDevice);
	cudaMemcpy(d_par, param6, sizeof(SCAL) * 6, cudaMemcpyHostToDevice);
	SCAL dts = (SCAL)dt; // main3.cu:231
	compute_force(coulombOscillatorFMMKD3_b200, d_buf, n, d_par);
	for (int s = 0; s < steps; ++s)
		leapfrog(coulombOscillatorFMMKD3_b200, d_buf, n, d_par, dts, step_b200, 1);
	if (cudaDeviceSynchronize() != cudaSuccess) return -2;
	cudaMemcpy(buf, d_buf, sizeof(VEC) * 3 * (size_t)n, cudaMemcpyDeviceToHost);
	cudaFree(d_buf); cudaFree(d_par);
	return 0;
}

// the reference's own test_accuracy (main3.cu:139-182, including its relerrReduce2 kernel) comparing the shim's FMM
// with the shim's direct sum; buf = [pos | vel | acc] on the host
float dropin_test_accuracy(float *buf, int n, const float *param6)
{
	SCAL *d_buf, *d_par;
	if (cudaMalloc(&d_buf, sizeof(VEC) * 3 * (size_t)n) != cudaSuccess) return -1.f;
	if (cudaMalloc(&d_par, sizeof(SCAL) * 6) != cudaSuccess) return -1.f;
	cudaMemcpy(d_buf, buf, sizeof(VEC) * 2 * (size_t)n, cudaMemcpyHostToDevice);
	cudaMemcpy(d_par, param6, sizeof(SCAL) * 6, cudaMemcpyHostToDevice);
	SCAL err = test_accuracy(fmm_cart3_kdtree_b200, direct3_b200, d_buf, n, d_par, true);
	cudaFree(d_buf); cudaFree(d_par);
	return (float)err;
}

} // extern "C"
