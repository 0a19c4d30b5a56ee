#include <cstdint>
__global__ void k(double *out)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = slot;
    const unsigned taddr = base + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 256u;
    double v = 1.0 + threadIdx.x;
    for (int e = 0; e < 128; ++e) {
        const double w = v * (e + 1);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr + 2u * e), "r"(__double2loint(w)), "r"(__double2hiint(w)) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    double acc = 0;
    for (int e = 0; e < 128; ++e) {
        unsigned lo, hi;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr + 2u * e) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __hiloint2double(hi, lo);
    }
    out[threadIdx.x] = acc - v * (128 * 129 / 2);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
#include <cstdio>
int main()
{
    double *d; cudaMalloc(&d, 256 * 8);
    k<<<148, 256>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    double h[256]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double worst = 0; for (int i = 0; i < 256; ++i) worst = fmax(worst, fabs(h[i]));
    printf("tmem roundtrip: %s, max |err| %g\n", cudaGetErrorString(e), worst);
    return 0;
}
