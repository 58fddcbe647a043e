// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M = 128) for the operand forms the kernels use,
// alone and under the shared-memory write traffic of the propagation kernel's producers (cp.async gathers, bulk loads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_rate scripts/micro/mma_rate.cu && build/mma_rate
// Output: cycles per MMA instruction (nominal: N / 2 cycles for K = 16) and the background bytes per cycle.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gnn_branching_b200/csrc/gnnb_umma.cuh"

using namespace gnnb::tcx;

enum Mode { SS_KMAJOR_N256 = 0, SS_MNMAJOR_N256, SS_NOSW_A_N128, TS_N64, TS_N128, SS_KMAJOR_N128, SS_KMAJOR_N64, SS_NOSW_A_N64, N_MODES };
const char* mode_name[N_MODES] = {"SS A,B K-major sw128  N=256", "SS B MN-major sw128   N=256 (prop)", "SS A no-swizzle       N=128 (update gemm1)",
                                  "TS A in TMEM          N=64  (chain)", "TS A in TMEM          N=128", "SS A,B K-major sw128  N=128",
                                  "SS A,B K-major sw128  N=64", "SS A no-swizzle       N=64"};

constexpr uint32_t A_BYTES = 32768, B_BYTES = 65536, BG_BYTES = 98304;

template <int mode, int nacc>
__global__ void __launch_bounds__(192, 1) k(int bg, int chunks, const unsigned char* __restrict__ gsrc, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sa = smem_u32(base), sb = sa + A_BYTES, sbg = sb + B_BYTES;
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    __shared__ long long bg_bytes[5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;   // 1.0 halves
    if (threadIdx.x == 0) { mbar_init(smem_u32(&mbar[0]), 1); mbar_init(smem_u32(&mbar[1]), 1); fence_mbar_init(); done = 0; }
    __syncwarp();
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    fence_proxy_async();
    const uint32_t tmem = slot & 0x0000FFFFu;
    if (warp == 0) {
        if (lane == 0) {
            uint32_t N = 256;
            if (mode == SS_NOSW_A_N128 || mode == TS_N128 || mode == SS_KMAJOR_N128) N = 128;
            if (mode == TS_N64 || mode == SS_KMAJOR_N64 || mode == SS_NOSW_A_N64) N = 64;
            const bool ts = (mode == TS_N64 || mode == TS_N128);
            const uint32_t d0 = ts ? 256u : 0u;          // TS: A operand at columns [0, 64), accumulators from 256 on
            uint32_t idesc = make_idesc(N);
            if (mode == SS_MNMAJOR_N256) idesc |= (1u << 16);
            // every operand descriptor is formed before the timed loop: the loop body is 12 tcgen05.mma and nothing else
            uint64_t ad[12], bd[12];
            uint32_t dd[12], at[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const int pass = i >> 2, kk = i & 3;
                const uint32_t a_addr = (pass == 1) ? sa + 16384 : sa;
                const uint32_t b_addr = (pass == 2) ? sb + 32768 : sb;
                dd[i] = tmem + d0 + (uint32_t)(i % nacc) * N;
                at[i] = tmem + 16 * kk + (pass == 1 ? 8 : 0);
                bd[i] = make_desc(b_addr) + 2 * kk;
                ad[i] = make_desc(a_addr) + 2 * kk;
                if (mode == SS_MNMAJOR_N256) bd[i] = make_desc_mn(b_addr + (kk >> 1) * 16384, 4096) + 128 * (kk & 1);
                if (mode == SS_NOSW_A_N128 || mode == SS_NOSW_A_N64) ad[i] = make_desc_nosw(a_addr, NB_PIECE, 128u) + (uint64_t)(kk * ((2 * NB_PIECE) >> 4));
            }
            const long long t0 = clock64();
#pragma unroll 1
            for (int c = 0; c < chunks; ++c) {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    if (ts) umma_ts(dd[i], at[i], bd[i], idesc, 1u);
                    else umma(dd[i], ad[i], bd[i], idesc, 1u);
                }
            }
            umma_commit(smem_u32(&mbar[0]));
            mbar_wait(smem_u32(&mbar[0]), 0);
            const long long t1 = clock64();
            done = 1;
            out[blockIdx.x * 8 + 0] = t1 - t0;
        }
    } else if (warp < 5 && (bg & 1)) {
        // cp.async gather like the propagation kernel: 4 rows x 8 chunks per instruction, 2 planes, 8 instructions per stage
        long long n = 0;
        uint32_t st = 0, seed = 12345u + warp * 977u + blockIdx.x * 31u + lane * 7919u;
        while (!done) {
            const uint32_t dst0 = sbg + (st & 1) * 32768 + (uint32_t)(warp - 1) * 4096;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                seed = seed * 1664525u + 1013904223u;
                const uint32_t row = __shfl_sync(0xffffffffu, seed >> 12, (lane >> 3) + 4 * (i & 1)) & 16383u;      // 16 K rows x 256 B = 4 MB region
                const unsigned char* src = gsrc + (size_t)row * 256 + (lane & 7) * 16;
                const uint32_t dst = dst0 + swz((uint32_t)(i * 4 + (lane >> 3)), (uint32_t)(lane & 7));
                cp_async16(dst, src, true);
                cp_async16(dst + 16384, src + 128, true);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            n += 8 * 2 * 16 * 32;
            ++st;
        }
        if (lane == 0) bg_bytes[warp - 1] = n;
    } else if (warp == 5 && (bg & 2)) {
        if (lane == 0) {
            long long n = 0;
            uint32_t ph = 0;
            while (!done) {
                mbar_expect_tx(smem_u32(&mbar[1]), 32768);
                bulk_g2s(sbg + 65536, gsrc + (size_t)((n >> 15) & 63) * 32768, 32768, smem_u32(&mbar[1]));
                mbar_wait(smem_u32(&mbar[1]), ph);
                ph ^= 1u;
                n += 32768;
            }
            bg_bytes[4] = n;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long g = 0;
        if (bg & 1) g += bg_bytes[0] + bg_bytes[1] + bg_bytes[2] + bg_bytes[3];
        out[blockIdx.x * 8 + 1] = g;
        out[blockIdx.x * 8 + 2] = (bg & 2) ? bg_bytes[4] : 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int M, int A>
void launch1(int grid, size_t smem, int bg, int chunks, const unsigned char* gsrc, long long* d) {
    cudaFuncSetAttribute(k<M, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<M, A><<<grid, 192, smem>>>(bg, chunks, gsrc, d);
}
template <int M>
void launch_m(int nacc, int grid, size_t smem, int bg, int chunks, const unsigned char* gsrc, long long* d) {
    if (nacc == 1) launch1<M, 1>(grid, smem, bg, chunks, gsrc, d);
    else if (nacc == 2) launch1<M, 2>(grid, smem, bg, chunks, gsrc, d);
    else launch1<M, 4>(grid, smem, bg, chunks, gsrc, d);
}
void launch(int mode, int nacc, int grid, size_t smem, int bg, int chunks, const unsigned char* gsrc, long long* d) {
    switch (mode) {
        case 0: launch_m<0>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 1: launch_m<1>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 2: launch_m<2>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 3: launch_m<3>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 4: launch_m<4>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 5: launch_m<5>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        case 6: launch_m<6>(nacc, grid, smem, bg, chunks, gsrc, d); break;
        default: launch_m<7>(nacc, grid, smem, bg, chunks, gsrc, d); break;
    }
}

int main() {
    const size_t smem = 1024 + A_BYTES + B_BYTES + BG_BYTES;
    unsigned char* gsrc; cudaMalloc(&gsrc, 8 << 20); cudaMemset(gsrc, 0x3c, 8 << 20);
    long long* d; cudaMalloc(&d, 148 * 8 * sizeof(long long));
    long long h[148 * 8];
    const int chunks = 400;
    for (int grid : {1, 148})
        for (int mode = 0; mode < N_MODES; ++mode)
            for (int nacc : {1, 2, 4})
            for (int bg = 0; bg < 4; bg += 3) {
                const int n = (mode == SS_KMAJOR_N256 || mode == SS_MNMAJOR_N256) ? 256 : (mode == SS_NOSW_A_N128 || mode == TS_N128 || mode == SS_KMAJOR_N128) ? 128 : 64;
                const bool ts = (mode == TS_N64 || mode == TS_N128);
                if ((ts ? 256 : 0) + nacc * n > 512) continue;
                if (grid == 1 && bg) continue;
                for (int rep = 0; rep < 2; ++rep) launch(mode, nacc, grid, smem, bg, chunks, gsrc, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, grid * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
                double cyc = 0, g = 0, t = 0;
                for (int b = 0; b < grid; ++b) { cyc += h[b * 8]; g += (double)h[b * 8 + 1] / h[b * 8]; t += (double)h[b * 8 + 2] / h[b * 8]; }
                cyc /= grid; g /= grid; t /= grid;
                printf("grid %3d  %-44s acc x%d bg[%s%s]  %7.1f cycles / MMA   gather %5.1f B/clk  bulk %5.1f B/clk\n", grid, mode_name[mode], nacc,
                       (bg & 1) ? "cp.async " : "", (bg & 2) ? "bulk" : "", cyc / (chunks * 12.0), g, t);
            }
    return 0;
}
