// Microbenchmark: HBM bandwidth of the throughput kernel's access pattern -- every warp reads (and
// optionally writes) whole 256-byte lines at pseudo-random positions inside its own ~1 MB tile, with
// NLOADS independent loads in flight per warp.  usage: randline <warps_per_sm> <tile_kb> <iters>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int NL>
__global__ void k(const double *base, size_t tile_doubles, int iters, double *out, int write_every)
{
	const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const double *tile = base + (size_t)warp * tile_doubles;
	const unsigned lines = (unsigned)(tile_doubles / 32);
	unsigned s = warp * 2654435761u + 12345u;
	double acc = 0;
	for (int it = 0; it < iters; ++it) {
		double v[NL];
#pragma unroll
		for (int u = 0; u < NL; ++u) {
			s = s * 1664525u + 1013904223u;
			const unsigned line = (s >> 8) % lines;
			asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[u]) : "l"(tile + (size_t)line * 32 + lane) : "memory");
		}
#pragma unroll
		for (int u = 0; u < NL; ++u) acc += v[u];
		if (write_every && it % write_every == 0) const_cast<double *>(tile)[(size_t)((s >> 8) % lines) * 32 + lane] = acc;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main(int argc, char **argv)
{
	const int wps = argc > 1 ? atoi(argv[1]) : 8, tile_kb = argc > 2 ? atoi(argv[2]) : 820, iters = argc > 3 ? atoi(argv[3]) : 2000;
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount, warps = sms * wps;
	const size_t tile_doubles = (size_t)tile_kb * 1024 / 8;
	double *base, *out;
	cudaMalloc(&base, warps * tile_doubles * 8); cudaMemset(base, 0, warps * tile_doubles * 8);
	cudaMalloc(&out, (size_t)warps * 32 * 8);
	for (int nl : {8, 16, 32}) {
		for (int rep = 0; rep < 2; ++rep) {
			cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
			cudaEventRecord(a);
			if (nl == 8) k<8><<<sms, wps * 32>>>(base, tile_doubles, iters, out, 8);
			else if (nl == 16) k<16><<<sms, wps * 32>>>(base, tile_doubles, iters, out, 8);
			else k<32><<<sms, wps * 32>>>(base, tile_doubles, iters, out, 8);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep) printf("warps/SM %d tile %d KB loads in flight/warp %d: %.1f GB/s (%s)\n", wps, tile_kb, nl, (double)warps * iters * nl * 256 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
		}
	}
	return 0;
}
