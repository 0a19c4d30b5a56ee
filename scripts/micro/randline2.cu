// As randline.cu, but each lane reads 16 bytes (a warp request = one 512-byte line): is the per-warp
// limit on outstanding requests or on outstanding bytes?  usage: randline2 <warps_per_sm> <tile_kb> <iters>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int NL>
__global__ void k(const double2 *base, size_t tile_elems, int iters, double *out)
{
	const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const double2 *tile = base + (size_t)warp * tile_elems;
	const unsigned lines = (unsigned)(tile_elems / 32);
	unsigned s = warp * 2654435761u + 12345u;
	double acc = 0;
	for (int it = 0; it < iters; ++it) {
		double2 v[NL];
#pragma unroll
		for (int u = 0; u < NL; ++u) {
			s = s * 1664525u + 1013904223u;
			const unsigned line = (s >> 8) % lines;
			asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(tile + (size_t)line * 32 + lane) : "memory");
		}
#pragma unroll
		for (int u = 0; u < NL; ++u) acc += v[u].x + v[u].y;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main(int argc, char **argv)
{
	const int wps = argc > 1 ? atoi(argv[1]) : 8, tile_kb = argc > 2 ? atoi(argv[2]) : 820, iters = argc > 3 ? atoi(argv[3]) : 2000;
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount, warps = sms * wps;
	const size_t tile_elems = (size_t)tile_kb * 1024 / 16;
	double2 *base; double *out;
	cudaMalloc(&base, warps * tile_elems * 16); cudaMemset(base, 0, warps * tile_elems * 16);
	cudaMalloc(&out, (size_t)warps * 32 * 8);
	for (int nl : {4, 8, 16}) {
		for (int rep = 0; rep < 2; ++rep) {
			cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
			cudaEventRecord(a);
			if (nl == 4) k<4><<<sms, wps * 32>>>(base, tile_elems, iters, out);
			else if (nl == 8) k<8><<<sms, wps * 32>>>(base, tile_elems, iters, out);
			else k<16><<<sms, wps * 32>>>(base, tile_elems, iters, out);
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep) printf("512B lines: warps/SM %d tile %d KB loads in flight/warp %d: %.1f GB/s (%s)\n", wps, tile_kb, nl, (double)warps * iters * nl * 512 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
		}
	}
	return 0;
}
