// Microbenchmark: how fast can one SM pull scattered 128-byte rows from a DRAM-resident table into shared memory?
//   mode 0  cp.async (LDGSTS) 16 B per lane, 4 warps                     — what the propagation gather does today
//   mode 1  cp.async.bulk.tensor.2d ... tile::gather4 (4 rows per instruction, SWIZZLE_128B), one issuing thread
//   mode 2  cp.async.bulk (linear) of one 128-byte row per instruction, one issuing thread
//   mode 3+ cp.async.bulk.tensor.2d tile (box) loads of R = 8 << (mode - 3) consecutive rows per instruction (R = 8 .. 64,
//           1 .. 8 KB), SWIZZLE_128B, one issuing thread: how the per-instruction cost of the TMA unit amortises
// Rows are 64 fp16 = 128 B; a "stage" is 64 rows (8 KB, one K step of the fused kernel: 16 rows x 2 subdomains x 2 planes);
// a ring of NST stages is kept in flight.  Row indices: runs of `run` consecutive rows at random places of a window of the
// table that slides forward with the iteration (argv[4] MB, 0 = the whole table: TLB-hostile), like a wave of subdomains.
// usage: gather_bw [rows] [iters] [box_rows 1|4] [window_MB]
// Prints bytes / cycle / SM and checks the landed data (row id in the first word of every 16-byte chunk).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/gather_bw scripts/micro/gather_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ROWS_PER_STAGE = 64, STAGE_BYTES = ROWS_PER_STAGE * 128, MAX_ST = 16;

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive(uint32_t mbar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar) : "memory"); }

__global__ void __launch_bounds__(160, 1) k(int mode, int nst, int iters, const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmbox,
                                            const unsigned char* table,
                                            const int* rows /* [grid][iters][64] */, long long* out, int* errs) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ uint64_t full[MAX_ST];
    const uint32_t ring = smem_u32(base);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int* myrows = rows + (size_t)blockIdx.x * iters * ROWS_PER_STAGE;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(smem_u32(&full[i]), mode == 0 ? 128 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0) {
        if (warp < 4) {      // warp w copies rows [16 w, 16 w + 16) of every stage: 4 rows x 8 chunks per instruction
            for (int it = 0; it < iters + nst; ++it) {
                const int s = it % nst;
                if (it >= nst) mbar_wait(smem_u32(&full[s]), ((it / nst) - 1) & 1);       // consume (nothing) = stage free again
                if (it < iters) {
                    const int idx = myrows[it * ROWS_PER_STAGE + warp * 16 + (lane & 15)];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kk = i * 4 + (lane >> 3);
                        const int r = __shfl_sync(0xffffffffu, idx, kk);
                        const int kr = warp * 16 + kk;
                        const uint32_t dst = ring + s * STAGE_BYTES + kr * 128 + (((lane & 7) ^ (kr & 7)) << 4);
                        cp_async16(dst, table + (size_t)r * 128 + (lane & 7) * 16);
                    }
                    cp_async_arrive(smem_u32(&full[s]));
                }
            }
            if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;      // timed by a thread that walked the loop
        }
    } else if (warp == 4) {
      if (lane == 0) {
        t0 = clock64();
        for (int it = 0; it < iters + nst; ++it) {
            const int s = it % nst;
            if (it >= nst) mbar_wait(smem_u32(&full[s]), ((it / nst) - 1) & 1);
            if (it < iters) {
                const uint32_t mb = smem_u32(&full[s]);
                mbar_expect_tx(mb, STAGE_BYTES);
                const int* rr = myrows + it * ROWS_PER_STAGE;
                if (mode == 1) {
                    for (int g4 = 0; g4 < ROWS_PER_STAGE / 4; ++g4)
                        gather4(ring + s * STAGE_BYTES + g4 * 512, &tm, mb, 0, rr[4 * g4], rr[4 * g4 + 1], rr[4 * g4 + 2], rr[4 * g4 + 3]);
                } else if (mode >= 3) {
                    const int R = 8 << (mode - 3);
                    for (int r = 0; r < ROWS_PER_STAGE; r += R) box2d(ring + s * STAGE_BYTES + r * 128, &tmbox, mb, 0, rr[r]);
                } else {
                    for (int r = 0; r < ROWS_PER_STAGE; ++r) bulk_g2s(ring + s * STAGE_BYTES + r * 128, table + (size_t)rr[r] * 128, 128, mb);
                }
            }
        }
        out[blockIdx.x] = clock64() - t0;
      }
      __syncwarp();
    }
    __syncthreads();
    // check the last stage that landed (iteration iters - 1)
    const int s = (iters - 1) % nst;
    int bad = 0;
    for (int i = threadIdx.x; i < ROWS_PER_STAGE * 8; i += blockDim.x) {
        const int kr = i >> 3, c = i & 7;
        const int want_row = myrows[(iters - 1) * ROWS_PER_STAGE + kr];
        const int pos = (mode == 2) ? c : (c ^ (kr & 7));          // every mode but 2 lands SWIZZLE_128B
        const uint32_t* p = reinterpret_cast<const uint32_t*>(base + s * STAGE_BYTES + kr * 128 + pos * 16);
        if (p[0] != (uint32_t)want_row || p[1] != (uint32_t)c) ++bad;
    }
    if (bad) atomicAdd(errs, bad);
}

int main(int argc, char** argv) {
    const long long nrows = argc > 1 ? atoll(argv[1]) : (8ll << 20);       // 8 Mi rows = 1 GiB
    const int iters = argc > 2 ? atoi(argv[2]) : 2000;
    const int boxrows_arg = argc > 3 ? atoi(argv[3]) : 1;
    const long long window_rows = (argc > 4 ? atoll(argv[4]) : 32) * 8192ll;          // MB -> rows of 128 B
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
    long long* out; int* errs;
    CK(cudaMalloc(&out, grid * sizeof(long long)));
    CK(cudaMalloc(&errs, sizeof(int)));
    int* drows;
    CK(cudaMalloc(&drows, (size_t)grid * iters * ROWS_PER_STAGE * sizeof(int)));
    CUtensorMap tm;
    for (int boxrows : {boxrows_arg}) {
        const cuuint64_t dims[2] = {64, (cuuint64_t)nrows};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled(box rows %d) -> %d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int run : {1, 16, 64}) {
            std::vector<int> h((size_t)grid * iters * ROWS_PER_STAGE);
            uint64_t st = 88172645463325252ull;
            for (size_t i = 0; i < h.size(); i += run) {
                st ^= st << 13; st ^= st >> 7; st ^= st << 17;
                long long r0 = (long long)(st % (uint64_t)(nrows - run));
                if (window_rows > 0) {      // window start moves with the iteration index (same for every CTA)
                    const long long itn = (long long)((i / ROWS_PER_STAGE) % iters);
                    const long long w0 = (nrows - window_rows - run) * itn / iters;
                    r0 = w0 + (long long)(st % (uint64_t)window_rows);
                }
                for (int j = 0; j < run && i + j < h.size(); ++j) h[i + j] = (int)(r0 + j);
            }
            CK(cudaMemcpy(drows, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
            for (int mode = 0; mode < 7; ++mode)
                for (int nst : {4, 16}) {
                    if (boxrows != 1 && mode != 1) continue;                // the tensor map only matters for mode 1
                    if (mode >= 3 && run < (8 << (mode - 3))) continue;     // box loads need runs of at least R consecutive rows
                    CUtensorMap tmb = tm;
                    if (mode >= 3) {
                        const cuuint32_t bb[2] = {64, (cuuint32_t)(8 << (mode - 3))};
                        CUresult rb = cuTensorMapEncodeTiled(&tmb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, bb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                        if (rb != CUDA_SUCCESS) { printf("box map R=%d -> %d\n", 8 << (mode - 3), (int)rb); continue; }
                    }
                    CK(cudaMemset(errs, 0, sizeof(int)));
                    const size_t smem = 1024 + (size_t)nst * STAGE_BYTES;
                    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    CK(cudaMemset(out, 0, grid * sizeof(long long)));
                    k<<<grid, 160, smem>>>(mode, nst, iters, tm, tmb, table, drows, out, errs);
                    CK(cudaGetLastError());
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d nst %d run %d box %d: %s\n", mode, nst, run, boxrows, cudaGetErrorString(e)); return 1; }
                    std::vector<long long> cyc(grid);
                    int herr = 0;
                    CK(cudaMemcpy(cyc.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(&herr, errs, sizeof(int), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (long long c : cyc) mx = c > mx ? c : mx;
                    printf("box %d  mode %d (%s)  run %2d  stages %2d (%3d KB in flight): %6.2f B/cycle/SM  (%.0f cycles/stage)  data errors %d\n", boxrows, mode,
                           mode == 0 ? "cp.async 16B x 4 warps" : mode == 1 ? "TMA gather4          " : mode == 2 ? "bulk 128 B per row   " :
                           mode == 3 ? "TMA box  8 rows 1 KB " : mode == 4 ? "TMA box 16 rows 2 KB " : mode == 5 ? "TMA box 32 rows 4 KB " : "TMA box 64 rows 8 KB ", run, nst, nst * STAGE_BYTES / 1024,
                           (double)iters * STAGE_BYTES / mx, (double)mx / iters, herr);
                }
        }
    }
    return 0;
}
