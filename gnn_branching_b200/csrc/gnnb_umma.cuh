// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels (gnnb_tc.cu, gnnb_prop_tc.cu).
#pragma once

#include <cuda_fp16.h>
#include <string.h>

#include "gnnb_common.cuh"

namespace gnnb {
namespace tcx {

constexpr int TILE = 128;                        // nodes per tile = UMMA M
constexpr int WGS = 2;                           // warpgroups (tiles in flight) per CTA
constexpr int NTHREADS = 128 * WGS;
constexpr uint32_t WPLANE = 64 * 64 * 2;         // one 64(n) x 64(k) fp16 weight plane: 8 KB
constexpr uint32_t APLANE = TILE * 64 * 2;       // one A plane: 16 KB
constexpr uint32_t ABUF = 2 * APLANE;            // hi + lo

// byte offset of 16-byte chunk `chunk` (8 fp16 along K) of row `row` inside a K-major SWIZZLE_128B tile whose rows
// are 128 bytes (64 fp16): 8-row groups are 1024 bytes apart (SBO), chunks are XOR-swizzled with the row
__host__ __device__ inline uint32_t swz(uint32_t row, uint32_t chunk) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

// ---- operand precision ----------------------------------------------------------------------------
// Operands are split into two fp16 planes, x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits,
// so the three products hi*hi + lo*hi + hi*lo carry ~2^-22 relative error (a bf16 split gives 2^-17, which left the
// scores only ~1.5x inside the 1e-4 parity bound).  fp16's narrow exponent is handled by keeping every activation
// operand scaled by ASCALE: ReLU and the linears are positively homogeneous, so the whole chain runs in the scaled
// domain with pre-scaled biases, and results are multiplied by AINV when they leave for global memory.
// |activation| >= 65504 / ASCALE = 524 288 overflows to inf -> NaN -> GNNB_ERR_NAN (never silent).
constexpr float ASCALE = 0.125f;
constexpr float AINV = 8.0f;

inline void split_host(float x, uint16_t& hi, uint16_t& lo) {
    const __half h = __float2half_rn(x);
    const float rem = x - __half2float(h);
    const __half l = (rem == rem && rem - rem == 0.f) ? __float2half_rn(rem) : __float2half_rn(0.f);
    memcpy(&hi, &h, 2);
    memcpy(&lo, &l, 2);
}

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 x fp16 -> fp32
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (lane = accumulator row).  The load is asynchronous:
// tmem_wait16 must run on the same array before its values are read (the "+r" operands pin that order for the compiler)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16], float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_sync(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_wait16(r, v);
}

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major), [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;                      // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                                 // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits [4,6) = 1), A/B fp16 (format 0), both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}

__device__ __forceinline__ float relu_nan(float x) { return (x != x) ? x : fmaxf(x, 0.f); }   // F.relu keeps NaN

// (a, b) -> packed fp16x2 hi and lo words; element a sits at the lower address
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const __half2 l = __floats2half2_rn(a - __low2float(h), b - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace tcx
}  // namespace gnnb
