// Microbenchmark: HBM bandwidth of the throughput kernel's access pattern with NO address arithmetic in the
// way: every warp reads whole lines (256 bytes: one double per lane, or 512: a double2 per lane) at
// pseudo-random positions inside its own 1 MB tile; NL independent loads are issued back to back, then
// consumed.  usage: randline3 <warps_per_sm> <iters>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int NL, int W>
__global__ void k(const double *base, int iters, double *out)
{
	const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const char *tile = reinterpret_cast<const char *>(base) + ((size_t)warp << 20) + lane * 8 * W;
	unsigned s = warp * 2654435761u + 12345u;
	double acc = 0;
	for (int it = 0; it < iters; ++it) {
		double v[NL * W];
#pragma unroll
		for (int u = 0; u < NL; ++u) {
			s = s * 1664525u + 1013904223u;
			const unsigned off = (s >> 8) & (W == 1 ? 0xfff00u : 0xffe00u); // line-aligned offset inside 1 MB
			if (W == 1) asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[u]) : "l"(tile + off) : "memory");
			else asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[2 * u]), "=d"(v[2 * u + 1]) : "l"(tile + off) : "memory");
		}
#pragma unroll
		for (int u = 0; u < NL * W; ++u) acc += v[u];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NL, int W>
void run(const double *base, double *out, int sms, int wps, int iters)
{
	for (int rep = 0; rep < 2; ++rep) {
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		cudaEventRecord(a);
		k<NL, W><<<sms, wps * 32>>>(base, iters, out);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (rep) printf("warps/SM %2d  line %3d B  loads in flight/warp %2d: %7.1f GB/s (%s)\n", wps, 256 * W, NL, (double)sms * wps * iters * NL * 256.0 * W / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
	}
}
int main(int argc, char **argv)
{
	const int iters = argc > 1 ? atoi(argv[1]) : 2000;
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	double *base, *out;
	cudaMalloc(&base, (size_t)sms * 32 << 20); cudaMemset(base, 0, (size_t)sms * 32 << 20);
	cudaMalloc(&out, (size_t)sms * 32 * 32 * 8);
	for (int wps : {8, 16, 32}) {
		run<8, 1>(base, out, sms, wps, iters); run<16, 1>(base, out, sms, wps, iters); run<32, 1>(base, out, sms, wps, iters); run<64, 1>(base, out, sms, wps, iters);
		run<8, 2>(base, out, sms, wps, iters); run<16, 2>(base, out, sms, wps, iters); run<32, 2>(base, out, sms, wps, iters);
	}
	return 0;
}
