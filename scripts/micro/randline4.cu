// Which path lifts the per-warp cap of randline3 (8 warps/SM pull ~3.7 TB/s of random 256-byte lines whatever
// the number of loads in flight)?  Variants, all 8 warps per SM, 32 requests per batch:
//   0 ld.global          1 ld.global.L1::no_allocate     2 ld.global.nc     3 ld.global.cg
//   4 cp.async (LDGSTS, 8 bytes per lane) into shared, wait_group 0 per batch
//   5 cp.async.bulk (TMA, one 256-byte copy per request issued by lane 0) into shared, mbarrier per batch
//   6 as 0 but 8 CTAs of one warp per SM instead of one CTA of 8 warps
//   7 as 0 with 200 KB of dynamic shared memory per CTA (small L1)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define NL 32
__device__ __forceinline__ unsigned next_off(unsigned &s) { s = s * 1664525u + 1013904223u; return (s >> 8) & 0xfff00u; }
template <int MODE>
__global__ void k(const double *base, int iters, double *out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const char *tile = reinterpret_cast<const char *>(base) + ((size_t)warp << 20);
	const char *mine = tile + lane * 8;
	unsigned s = warp * 2654435761u + 12345u;
	double acc = 0;
	if (MODE == 4 || MODE == 5) {
		unsigned char *ring = smem + wib * (NL * 256 + 64);
		const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + 64;
		const unsigned bar = ring_s - 64;
		if (MODE == 5) {
			if (lane == 0) asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(bar));
			__syncwarp();
		}
		unsigned phase = 0;
		for (int it = 0; it < iters; ++it) {
			if (MODE == 4) {
#pragma unroll
				for (int u = 0; u < NL; ++u) {
					const unsigned off = next_off(s);
					asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ring_s + u * 256 + lane * 8), "l"(mine + off) : "memory");
				}
				asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
			} else {
				if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(bar), "r"(NL * 256) : "memory");
#pragma unroll
				for (int u = 0; u < NL; ++u) {
					const unsigned off = next_off(s);
					if (lane == 0)
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];" ::"r"(ring_s + u * 256), "l"(tile + off), "r"(bar) : "memory");
				}
				unsigned done = 0;
				while (!done)
					asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
				phase ^= 1;
			}
			__syncwarp();
#pragma unroll
			for (int u = 0; u < NL; ++u) acc += *reinterpret_cast<volatile double *>(ring + 64 + u * 256 + lane * 8);
			__syncwarp();
		}
	} else {
		for (int it = 0; it < iters; ++it) {
			double v[NL];
#pragma unroll
			for (int u = 0; u < NL; ++u) {
				const unsigned off = next_off(s);
				if (MODE == 1) asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v[u]) : "l"(mine + off) : "memory");
				else if (MODE == 2) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v[u]) : "l"(mine + off) : "memory");
				else if (MODE == 3) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v[u]) : "l"(mine + off) : "memory");
				else asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[u]) : "l"(mine + off) : "memory");
			}
#pragma unroll
			for (int u = 0; u < NL; ++u) acc += v[u];
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char *what, const double *base, double *out, int sms, int iters)
{
	const int wps = 8;
	int grid = sms, block = wps * 32;
	size_t smem = 0;
	if (MODE == 4 || MODE == 5) smem = wps * (NL * 256 + 64);
	if (MODE == 6) { grid = sms * wps; block = 32; }
	if (MODE == 7) smem = 200 * 1024;
	cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	for (int rep = 0; rep < 2; ++rep) {
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		cudaEventRecord(a);
		k<MODE><<<grid, block, smem>>>(base, iters, out);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (rep) printf("%-44s %7.1f GB/s (%s)\n", what, (double)sms * wps * iters * NL * 256.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
	}
}
int main(int argc, char **argv)
{
	const int iters = argc > 1 ? atoi(argv[1]) : 1000;
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	double *base, *out;
	cudaMalloc(&base, (size_t)sms * 8 << 20); cudaMemset(base, 0, (size_t)sms * 8 << 20);
	cudaMalloc(&out, (size_t)sms * 8 * 32 * 8);
	run<0>("0 ld.global", base, out, sms, iters);
	run<1>("1 ld.global.L1::no_allocate", base, out, sms, iters);
	run<2>("2 ld.global.nc", base, out, sms, iters);
	run<3>("3 ld.global.cg", base, out, sms, iters);
	run<4>("4 cp.async 8 B per lane -> shared", base, out, sms, iters);
	run<5>("5 cp.async.bulk 256 B -> shared (TMA)", base, out, sms, iters);
	run<6>("6 ld.global, 8 CTAs of one warp per SM", base, out, sms, iters);
	run<7>("7 ld.global, 200 KB dynamic smem (small L1)", base, out, sms, iters);
	return 0;
}
