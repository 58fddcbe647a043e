// Microbenchmark 2: does the per-SM rate of scattered-row loads scale with the number of ISSUING WARPS?
// Every issuing warp owns a private ring of `nst` stages of 64 rows x 128 B (8 KB) and loads random runs of rows of a
// DRAM-resident table into it:
//   mode 0  cp.async (LDGSTS) 16 B per lane (all 32 lanes: 16 instructions per lane and stage)
//   mode 1  TMA gather4 (lane 0: 16 instructions per stage)
//   mode 2  TMA 2-D box loads of R rows (lane 0: 64 / R instructions per stage), R = argv[3] (8 .. 64)
//   mode 3  mode 0 + prefetch.global.L2 of the rows of the stage PFD = 8 iterations ahead (2 lines per lane)
//   mode 4  cp.async with commit groups instead of cp.async.mbarrier.arrive: `nst` stages of the warp in flight, a stage is
//           published with a plain mbarrier arrive after cp.async.wait_group nst - 1
//   mode 5  mode 4 + the L2 prefetch of mode 3
// Prints bytes / cycle / SM for nw = 1, 2, 4, 8 warps.  (gather_bw.cu showed that one issuing thread gets the same rate with 4
// or 16 stages in flight: the limit is not the ring depth.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/gather_bw2 scripts/micro/gather_bw2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROWS = 64, STAGE_BYTES = ROWS * 128, MAX_W = 8, MAX_ST = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(a), "r"(ph) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, uint32_t mbar, int c0, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(mbar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void box2d(uint32_t dst, const CUtensorMap* tm, uint32_t mbar, int c0, int r0) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(mbar), "r"(c0), "r"(r0) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive(uint32_t mbar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory"); }

__global__ void __launch_bounds__(256, 1) k(int mode, int nw, int nst, int iters, int R, const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tmbox,
                                            const unsigned char* table, const int* rows /* [grid][MAX_W][iters][64] */, long long* out, int* errs) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[MAX_W][MAX_ST];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int w = 0; w < nw; ++w)
            for (int i = 0; i < nst; ++i) mbar_init(smem_u32(&full[w][i]), (mode == 0 || mode == 3) ? 32 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < nw) {
        const uint32_t ring = smem_u32(base) + warp * nst * STAGE_BYTES;
        const int* myrows = rows + ((size_t)blockIdx.x * MAX_W + warp) * iters * ROWS;
        const long long t0 = clock64();
        if (mode == 0 || mode >= 3 || lane == 0) {
            for (int it = 0; it < iters + nst; ++it) {
                const int s = it % nst;
                const uint32_t mb = smem_u32(&full[warp][s]);
                if (mode >= 4) {
                    // commit groups: before stage `it` overwrites slot s, the copy issued nst iterations ago must have landed
                    if (it >= nst) {
                        if (nst == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
                        else if (nst == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
                        else asm volatile("cp.async.wait_group 7;" ::: "memory");
                    }
                } else if (it >= nst) mbar_wait(mb, ((it / nst) - 1) & 1);
                if (it < iters) {
                    const int* rr = myrows + it * ROWS;
                    const uint32_t st = ring + s * STAGE_BYTES;
                    if ((mode == 3 || mode == 5) && it + 8 < iters) {
                        const int* pr = myrows + (it + 8) * ROWS;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(table + (size_t)pr[lane] * 128));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(table + (size_t)pr[32 + lane] * 128));
                    }
                    if (mode == 0 || mode >= 3) {
                        const int i0 = rr[lane], i1 = rr[32 + lane];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int kr = i * 4 + (lane >> 3);
                            const int r = __shfl_sync(0xffffffffu, i < 8 ? i0 : i1, kr & 31);
                            cp_async16(st + kr * 128 + (((lane & 7) ^ (kr & 7)) << 4), table + (size_t)r * 128 + (lane & 7) * 16);
                        }
                        if (mode >= 4) asm volatile("cp.async.commit_group;" ::: "memory");
                        else cp_async_arrive(mb);
                    } else {
                        mbar_expect_tx(mb, STAGE_BYTES);
                        if (mode == 1) for (int g = 0; g < ROWS / 4; ++g) gather4(st + g * 512, &tm4, mb, 0, rr[4 * g], rr[4 * g + 1], rr[4 * g + 2], rr[4 * g + 3]);
                        else for (int r = 0; r < ROWS; r += R) box2d(st + r * 128, &tmbox, mb, 0, rr[r]);
                    }
                }
            }
        }
        if (mode >= 4) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) out[blockIdx.x * MAX_W + warp] = clock64() - t0;
        // check the last stage
        const int s = (iters - 1) % nst;
        int bad = 0;
        for (int i = lane; i < ROWS * 8; i += 32) {
            const int kr = i >> 3, c = i & 7;
            const int want = myrows[(iters - 1) * ROWS + kr];
            const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (warp * nst + s) * STAGE_BYTES + kr * 128 + ((c ^ (kr & 7)) << 4));
            if (p[0] != (uint32_t)want || p[1] != (uint32_t)c) ++bad;
        }
        if (bad) atomicAdd(errs, bad);
    }
}

int main(int argc, char** argv) {
    const long long nrows = argc > 1 ? atoll(argv[1]) : (8ll << 20);
    const int iters = argc > 2 ? atoi(argv[2]) : 500;
    const int R = argc > 3 ? atoi(argv[3]) : 32;
    const long long window_rows = (argc > 4 ? atoll(argv[4]) : 32) * 8192ll;
    CK(cudaSetDevice(0));
    unsigned char* table;
    CK(cudaMalloc(&table, (size_t)nrows * 128));
    {
        std::vector<uint32_t> h((size_t)nrows * 32);
        for (long long r = 0; r < nrows; ++r)
            for (int c = 0; c < 8; ++c) { h[r * 32 + c * 4] = (uint32_t)r; h[r * 32 + c * 4 + 1] = (uint32_t)c; h[r * 32 + c * 4 + 2] = 0; h[r * 32 + c * 4 + 3] = 0; }
        CK(cudaMemcpy(table, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    const int grid = 148;
    long long* out; int* errs; int* drows;
    CK(cudaMalloc(&out, grid * MAX_W * sizeof(long long)));
    CK(cudaMalloc(&errs, sizeof(int)));
    const size_t nidx = (size_t)grid * MAX_W * iters * ROWS;
    CK(cudaMalloc(&drows, nidx * sizeof(int)));
    {
        std::vector<int> h(nidx);
        uint64_t st = 88172645463325252ull;
        const int run = 64;          // runs of 64 consecutive rows: the same indices serve every mode
        for (size_t i = 0; i < nidx; i += run) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            const long long itn = (long long)((i / ROWS) % iters);
            const long long w0 = (nrows - window_rows - run) * itn / iters;
            const long long r0 = w0 + (long long)(st % (uint64_t)window_rows);
            for (int j = 0; j < run; ++j) h[i + j] = (int)(r0 + j);
        }
        CK(cudaMemcpy(drows, h.data(), nidx * sizeof(int), cudaMemcpyHostToDevice));
    }
    CUtensorMap tm4, tmb;
    const cuuint64_t dims[2] = {64, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t b4[2] = {64, 1}, bb[2] = {64, (cuuint32_t)R};
    if (cuTensorMapEncodeTiled(&tm4, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, b4, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        cuTensorMapEncodeTiled(&tmb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, bb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); return 1; }
    for (int mode : {0, 3, 4, 5})
        for (int nw : {1, 4})
            for (int nst : {2, 4, 8}) {
                if (nw * nst > 24) continue;
                if (mode < 4 && nst == 8) continue;
                CK(cudaMemset(errs, 0, sizeof(int)));
                CK(cudaMemset(out, 0, grid * MAX_W * sizeof(long long)));
                const size_t smem = 1024 + (size_t)nw * nst * STAGE_BYTES;
                CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k<<<grid, 256, smem>>>(mode, nw, nst, iters, R, tm4, tmb, table, drows, out, errs);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                std::vector<long long> cyc(grid * MAX_W);
                int herr = 0;
                CK(cudaMemcpy(cyc.data(), out, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&herr, errs, sizeof(int), cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (long long c : cyc) mx = c > mx ? c : mx;
                printf("mode %d (%s R=%2d)  warps %d  stages/warp %d (%3d KB in flight): %6.2f B/cycle/SM  (%.0f cycles per 8 KB stage and warp)  data errors %d\n", mode,
                       mode == 0 ? "cp.async 16 B / lane" : mode == 1 ? "TMA gather4         " : mode == 2 ? "TMA box             " : mode == 3 ? "cp.async + L2 pf    " :
                       mode == 4 ? "cp.async commit grps" : "commit grps + L2 pf ", mode == 2 ? R : (mode == 1 ? 4 : 0), nw, nst,
                       nw * nst * STAGE_BYTES / 1024, (double)iters * STAGE_BYTES * nw / mx, (double)mx / iters, herr);
            }
    return 0;
}
