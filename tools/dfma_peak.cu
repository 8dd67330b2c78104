// tools/dfma_peak.cu -- measures the FP64 FMA issue peak of the GPU (roofline denominator of the 2D fp64 near
// field): 8 independent DFMA chains per thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dfma_peak tools/dfma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
	double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; ++i)
	{
#pragma unroll 8
		for (int u = 0; u < 8; ++u)
		{
			x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
			x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
		}
	}
	out[blockIdx.x * 256 + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main()
{
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	double *d; cudaMalloc(&d, sizeof(double) * sms * 8 * 256);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 20000;
	for (int rep = 0; rep < 3; ++rep)
	{
		cudaEventRecord(e0);
		k<<<sms * 8, 256>>>(d, iters, 0.999999, 1e-9);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		double dfma = (double)sms * 8 * 256 * iters * 64.0;
		printf("sms %d: %.3f ms, %.3f T DFMA/s = %.2f TFLOP/s fp64, %.2f DFMA/clk/SM at 1.965 GHz\n", sms, ms, dfma / ms / 1e9,
		       2 * dfma / ms / 1e9, dfma / (ms * 1e-3) / sms / 1.965e9);
	}
	return 0;
}
