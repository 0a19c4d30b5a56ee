// Latency probes for the interpreter's building blocks on sm_100a: nvcc -arch=sm_100a -fmad=false lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cyc, int iters, int mode)
{
	__shared__ double sm[2048];
	__shared__ int idx[2048];
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; idx[i] = (i * 37 + 11) & 2047; }
	__syncthreads();
	double a = 1.0 + threadIdx.x * 1e-6, b = 1.0000001, c = 0.9999999;
	int j = threadIdx.x;
	long long t0 = clock64();
	if (mode == 0) for (int i = 0; i < iters; ++i) { a = a + b; }                       // dependent DADD
	else if (mode == 1) for (int i = 0; i < iters; ++i) { a = a * b; }                  // dependent DMUL
	else if (mode == 2) for (int i = 0; i < iters; ++i) { a = a - b * c; b = b + 1e-9; } // DMUL independent, DADD dependent
	else if (mode == 3) for (int i = 0; i < iters; ++i) { a = a / b; }                  // dependent division
	else if (mode == 4) for (int i = 0; i < iters; ++i) { j = idx[j]; }                 // dependent LDS.32
	else if (mode == 5) for (int i = 0; i < iters; ++i) { a = a - sm[j] * sm[(j + 1) & 2047]; j = (j + 33) & 2047; } // 2 LDS + mul + sub
	else if (mode == 6) for (int i = 0; i < iters; ++i) { a = a - sm[j] * sm[(j + 1) & 2047]; j = (j + 33) & 2047; sm[(j + 700) & 2047] = a; __syncwarp(); }
	else if (mode == 7) for (int i = 0; i < iters; ++i) { int k = j; k = k * 3 + 1; k = k ^ (k >> 3); k = k + 7; k = k * 5; j = k & 2047; } // dependent int ALU x5
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = a + j;
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = t1 - t0;
}
int main()
{
	double *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 8); cudaMallocManaged(&cyc, 64 * 8);
	const char *names[] = {"dependent DADD", "dependent DMUL", "DADD dep + DMUL indep", "dependent DDIV", "dependent LDS.32", "2 LDS + DMUL + DADD", "same + STS + syncwarp", "5 dependent int ops"};
	for (int warps : {1, 2, 7, 16}) {
		printf("== %d warps per SM (148 CTAs)\n", warps);
		for (int mode = 0; mode < 8; ++mode) {
			const int iters = 4096;
			probe<<<148, warps * 32>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
			probe<<<148, warps * 32>>>(out, cyc, iters, mode); cudaDeviceSynchronize();
			printf("  %-26s %.1f cycles/iter\n", names[mode], (double)cyc[mode] / iters);
		}
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
