// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (bytes per cycle) for 1..4 warpgroups.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_bw scripts/micro/tmem_bw.cu && build/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: ld x16, 1: st x16, 2: ld x32 (two x16 back to back before the wait)
__global__ void k(long long* out, int iters) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + (threadIdx.x >> 7) * 128u;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t a = base + (uint32_t)((q & 7) * 16);
            if (MODE == 1) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
            } else {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a));
                if (MODE == 0) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[3]; }
            }
        }
        if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (MODE == 2) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[5]; }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 16); long long h[2];
    const int iters = 2000;
    const char* names[3] = {"ld x16 (wait each)", "st x16 (wait per 8)", "ld x16 (wait per 8)"};
    for (int mode = 0; mode < 3; ++mode)
        for (int threads = 128; threads <= 512; threads *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, threads>>>(d, iters);
                if (mode == 1) k<1><<<1, threads>>>(d, iters);
                if (mode == 2) k<2><<<1, threads>>>(d, iters);
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            }
            const double bytes = (double)iters * 8 * 16 * 4 * threads;      // per instruction: 16 columns x 4 B per thread
            printf("%-22s %d warpgroup(s): %lld cycles, %.1f bytes/cycle/SM, %.1f cycles per warp-instruction\n", names[mode], threads / 128,
                   h[0], bytes / h[0], (double)h[0] / (iters * 8));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
