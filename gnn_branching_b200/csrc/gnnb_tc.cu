// Per-node MLP kernels on the 5th-generation tensor cores (GNNB_MATH_TC_FP16X3) — the product path.
//
// Every dense layer of the GNN with K >= 64 is a [128 nodes x K] x [K x N] GEMM per tile, issued as
// tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 64 / 128 / 192, K = 16 per instruction) with fp32
// accumulators in tensor memory.  fp32 accuracy (scores within 1e-4, BASELINE.json) comes from an fp16 hi/lo
// split of BOTH operands and three MMAs per K step:  x*w ~= xh*wh + xl*wh + xh*wl  (error ~2^-22 per product;
// see gnnb_umma.cuh for the scaled operand domain that keeps fp16 in range).
//
// Algebra that shortens the per-tile chain (exact up to fp32 rounding; composed on the host in double):
//   * two linears with nothing but a bias between them are one linear.  The reference computes
//       mu = W4_2 relu(W4 [relax, e] + b4),  e = W3_2 relu(W3 [nb r0, nb r1] + b3) + b3_2,
//       relax = amb (W1_1 h1 + b1_1)                                              (graph_conv.py:153-181, 273-347)
//     so with W4 = [W4a | W4b]:  W4 [relax, e] + b4 = amb (Wr h1 + br) + Wc h3 + bc,  Wr = W4a W1_1, br = W4a b1_1,
//     Wc = W4b W3_2, bc = b4 + W4b b3_2.  The relax kernel therefore emits relax' = amb (Wr h1 + br) directly (the
//     round-independent part of the fc4 / bc4 pre-activation), and the update kernel is three GEMMs per tile:
//       D = nb [W3a; W3b]^T (N = 128) -> h3 = relu(r0 Da + r1 Db + b3) -> D = h3 Wc^T -> g = relu(D + relax' + bc)
//       -> D = g W4_2^T -> mu = (D + b4_2) (r0 != 0)   [-> score head on the last backward sweep];
//   * row scalings of GEMM *inputs* move to the epilogue by linearity: [nb r0, nb r1] W3^T is one
//     N = 128 MMA, bc2's [s1, -d2 s1, d1 s1] input one N = 192 MMA.
//
// Structure of a CTA (1 per SM, persistent over tiles): NWG warpgroups of 128 threads; the stage's weights sit in
// shared memory for the CTA's lifetime as fp16 hi/lo planes in the UMMA K-major SWIZZLE_128B image (repacked once on
// the host, landed with one cp.async.bulk per linear); each warpgroup owns one 128-node tile at a time with its own
// 32 KB A-operand buffer, 512 / NWG TMEM columns and two mbarriers, and walks the chain of its tile sequentially:
//   A operand (TMA of a pre-split tile image, or thread = node row writing hi/lo planes) -> fence.proxy.async ->
//   warpgroup barrier -> one thread issues the MMAs + tcgen05.commit -> everyone waits on the mbarrier ->
//   tcgen05.ld -> bias / ReLU / row scaling in registers -> hi/lo split -> next A ...
// The warpgroups are independent, so one tile's epilogue overlaps the others' MMAs and loads.
//
// Private workspace layouts (produced and consumed only by the tensor-core kernels):
//   nb       per tile of 128 consecutive rows: the A-operand image itself, [hi plane 16 KB][lo plane 16 KB], scaled by
//            ASCALE — written by the propagation kernels, loaded with one 32 KB cp.async.bulk;
//   relax'   [tile][16 channel quads][128 rows][4] fp32 (scaled domain), so that thread = row reads are coalesced.
#include "gnnb_umma.cuh"

namespace gnnb {
namespace {

using namespace tcx;

// ---- warpgroup context ------------------------------------------------------------------------------
struct WG {
    uint32_t a_hi, a_lo;        // shared addresses of this warpgroup's A planes (a_lo = a_hi + APLANE)
    uint32_t mbar_mma, mbar_tma;
    uint32_t ph_mma, ph_tma;
    uint32_t tmem;              // TMEM address: lane base of this warp, first column of this warpgroup
    int t;                      // thread index within the warpgroup = row of the tile this thread owns
    int wg;
};

__device__ __forceinline__ void wg_barrier(const WG& c) { named_bar(1 + c.wg, 128); }

// one 32 KB pre-split tile image (global) -> this warpgroup's A planes.  The A buffer must be free: its last MMA has
// completed and generic-proxy accesses to it were fenced (fence.proxy.async) before the preceding barrier.
__device__ __forceinline__ void tma_tile(const WG& c, const void* src) {
    if (c.t == 0) {
        mbar_expect_tx(c.mbar_tma, ABUF);
        bulk_g2s(c.a_hi, src, ABUF, c.mbar_tma);
    }
}

// make this thread's A-plane writes visible to the tensor core, then one thread issues 3 x 4 MMAs + commit
template <bool WAIT_TMA>
__device__ __forceinline__ void gemm_start(WG& c, uint32_t b_hi, uint32_t b_lo, uint32_t N, uint32_t dcol, bool accumulate) {
    fence_proxy_async();
    tc_fence_before();
    wg_barrier(c);
    if (c.t == 0) {
        tc_fence_after();
        if (WAIT_TMA) mbar_wait(c.mbar_tma, c.ph_tma);
        const uint32_t idesc = make_idesc(N);
        const uint32_t d = (c.tmem & 0x0000FFFFu) + dcol;      // lane 0, column base
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
            const uint64_t ad = make_desc(pass == 1 ? c.a_lo : c.a_hi);
            const uint64_t bd = make_desc(pass == 2 ? b_lo : b_hi);
#pragma unroll
            for (int k = 0; k < 4; ++k)      // 16 fp16 = 32 bytes along K per instruction
                umma(d, ad + 2 * k, bd + 2 * k, idesc, (accumulate || pass > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(c.mbar_mma);
    }
    if (WAIT_TMA) c.ph_tma ^= 1u;
}
__device__ __forceinline__ void gemm_finish(WG& c) {
    mbar_wait(c.mbar_mma, c.ph_mma);
    c.ph_mma ^= 1u;
    tc_fence_after();
}
template <bool WAIT_TMA = false>
__device__ __forceinline__ void gemm(WG& c, uint32_t b_hi, uint32_t b_lo, uint32_t N, uint32_t dcol, bool accumulate) {
    gemm_start<WAIT_TMA>(c, b_hi, b_lo, N, dcol, accumulate);
    gemm_finish(c);
}

// 8 consecutive features [8*chunk, 8*chunk+8) of this thread's row -> A planes
__device__ __forceinline__ void a_store8(const WG& c, int chunk, const float (&v)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const uint32_t off = swz((uint32_t)c.t, (uint32_t)chunk);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(c.a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(c.a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// read-only global loads the compiler must not sink past the MMA waits (they are issued early to hide DRAM latency)
__device__ __forceinline__ float4 ldg4_now(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg1_now(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// accumulator columns [dcol, dcol+64) + bias (optionally ReLU) -> A planes
template <bool RELU>
__device__ __forceinline__ void epilogue_to_a(const WG& c, uint32_t dcol, const float* __restrict__ bias_s) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        tmem_ld16_sync(c.tmem + dcol + q * 16, v);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 b0 = lds4(bias_s + q * 16 + h * 8), b1 = lds4(bias_s + q * 16 + h * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = v[h * 8 + j] + bb[j];
                o[j] = RELU ? relu_nan(x) : x;
            }
            a_store8(c, q * 2 + h, o);
        }
    }
}

// K < 64 first layer on CUDA cores: relu(bias + sum_k feat[k] * wt[k][:]) -> A planes.  wt_s: fp32 [K][64] in smem
template <int K>
__device__ __forceinline__ void first_layer_to_a(const WG& c, const float (&feat)[K], const float* __restrict__ wt_s,
                                                 const float* __restrict__ bias_s) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        float o[8];
        const float4 b0 = lds4(bias_s + ch * 8), b1 = lds4(bias_s + ch * 8 + 4);
        o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 w0 = lds4(wt_s + k * P + ch * 8), w1 = lds4(wt_s + k * P + ch * 8 + 4);
            o[0] = fmaf(feat[k], w0.x, o[0]); o[1] = fmaf(feat[k], w0.y, o[1]); o[2] = fmaf(feat[k], w0.z, o[2]); o[3] = fmaf(feat[k], w0.w, o[3]);
            o[4] = fmaf(feat[k], w1.x, o[4]); o[5] = fmaf(feat[k], w1.y, o[5]); o[6] = fmaf(feat[k], w1.z, o[6]); o[7] = fmaf(feat[k], w1.w, o[7]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = relu_nan(o[j]);
        a_store8(c, ch, o);
    }
}

// (accumulator columns [dcol, dcol+64) + bias) * rowscale * AINV -> global fp32 [rows][64], staged through this
// warpgroup's A buffer (free at this point) so that the global stores are full 256-byte rows.
// Returns true if this thread's (valid) row holds a NaN.
__device__ __forceinline__ bool epilogue_to_global(const WG& c, uint32_t dcol, const float* __restrict__ bias_s, float rowscale,
                                                   float* __restrict__ dst, int64_t row0, int64_t rows) {
    bool bad = false;
    rowscale *= AINV;                                  // leave the scaled operand domain
    const uint32_t rbase = c.a_hi + (uint32_t)c.t * 256u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        tmem_ld16_sync(c.tmem + dcol + q * 16, v);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float4 b4 = lds4(bias_s + q * 16 + h * 4);
            float o[4] = {(v[h * 4 + 0] + b4.x) * rowscale, (v[h * 4 + 1] + b4.y) * rowscale, (v[h * 4 + 2] + b4.z) * rowscale,
                          (v[h * 4 + 3] + b4.w) * rowscale};
            bad |= (o[0] != o[0]) | (o[1] != o[1]) | (o[2] != o[2]) | (o[3] != o[3]);
            const uint32_t chunk = (uint32_t)(q * 4 + h);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + ((chunk ^ ((uint32_t)c.t & 7u)) << 4)), "f"(o[0]),
                         "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
        }
    }
    wg_barrier(c);
    const int lane = c.t & 31, warp = c.t >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int rr = warp * 32 + 2 * i + (lane >> 4);
        const uint32_t chunk = (uint32_t)(lane & 15);
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "r"(c.a_hi + (uint32_t)rr * 256u + ((chunk ^ ((uint32_t)rr & 7u)) << 4)));
        const int64_t grow = row0 + rr;
        if (grow < rows) *(reinterpret_cast<float4*>(dst + grow * P) + chunk) = o;
    }
    fence_proxy_async();       // the next user of the A buffer may be the async proxy (TMA tile load)
    wg_barrier(c);             // staging fully read before the A buffer is written again
    return bad && (row0 + c.t < rows);
}

// (accumulator columns [dcol, dcol+64) + bias) * rowscale * AINV -> this thread's row of global fp32 [rows][64], straight
// from registers (16-byte pieces; the A buffer stays free for the next tile's TMA).  Returns true on NaN in a valid row.
__device__ __forceinline__ bool epilogue_to_global_direct(const WG& c, uint32_t dcol, const float* __restrict__ bias_s, float rowscale,
                                                          float* __restrict__ dst, int64_t row0, int64_t rows) {
    bool bad = false;
    rowscale *= AINV;
    const bool ok = row0 + c.t < rows;
    float4* out = reinterpret_cast<float4*>(dst + (row0 + c.t) * P);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        tmem_ld16_sync(c.tmem + dcol + q * 16, v);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float4 b4 = lds4(bias_s + q * 16 + h * 4);
            const float4 o = make_float4((v[h * 4 + 0] + b4.x) * rowscale, (v[h * 4 + 1] + b4.y) * rowscale,
                                         (v[h * 4 + 2] + b4.z) * rowscale, (v[h * 4 + 3] + b4.w) * rowscale);
            bad |= (o.x != o.x) | (o.y != o.y) | (o.z != o.z) | (o.w != o.w);
            if (ok) out[q * 4 + h] = o;
        }
    }
    return bad && ok;
}

// (accumulator columns [dcol, dcol+64) + bias) * rowscale -> relax' layout [tile][16][128][4] (scaled domain, coalesced)
__device__ __forceinline__ void epilogue_to_rlx(const WG& c, uint32_t dcol, const float* __restrict__ bias_s, float rowscale,
                                                float* __restrict__ dst_tile) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        tmem_ld16_sync(c.tmem + dcol + q * 16, v);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float4 b4 = lds4(bias_s + q * 16 + h * 4);
            const float4 o = make_float4((v[h * 4 + 0] + b4.x) * rowscale, (v[h * 4 + 1] + b4.y) * rowscale,
                                         (v[h * 4 + 2] + b4.z) * rowscale, (v[h * 4 + 3] + b4.w) * rowscale);
            *reinterpret_cast<float4*>(dst_tile + ((size_t)(q * 4 + h) * TILE + c.t) * 4) = o;
        }
    }
}

// ---- shared-memory layout -----------------------------------------------------------------------------
constexpr int MAXWG = 4;
struct Tail {                   // small fp32 data after the weight planes and A buffers
    float bias[6][P];           // pre-scaled by ASCALE
    float w_small[2][8 * P];    // first-layer weights (K <= 7), transposed [K][64], pre-scaled
    float vec[P];
    uint64_t mbar[1 + 2 * MAXWG];
    uint32_t tmem_slot;
};

struct CtaSetup {
    unsigned char* base;        // 1024-aligned dynamic shared memory
    uint32_t w;                 // shared address of the weight planes
    Tail* tail;
    uint32_t tmem_base;
};

// common prologue: carve shared memory, allocate TMEM, init mbarriers, TMA the weight planes in
template <int NWG, int NW>
__device__ __forceinline__ CtaSetup cta_setup(uint32_t wbytes, const uint16_t* const (&wsrc)[NW], const uint32_t (&woff)[NW],
                                              const uint32_t (&wlen)[NW]) {
    extern __shared__ unsigned char smem_raw[];
    CtaSetup s;
    s.base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // pointer arithmetic keeps the shared address space
    s.w = smem_u32(s.base);
    s.tail = reinterpret_cast<Tail*>(s.base + wbytes + NWG * ABUF);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + 2 * NWG; ++i) mbar_init(smem_u32(&s.tail->mbar[i]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s.tail->tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    s.tmem_base = s.tail->tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t mb = smem_u32(&s.tail->mbar[0]);
        uint32_t total = 0;
        for (int i = 0; i < NW; ++i) total += wlen[i];
        mbar_expect_tx(mb, total);
        for (int i = 0; i < NW; ++i) bulk_g2s(s.w + woff[i], wsrc[i], wlen[i], mb);
    }
    return s;
}
template <int NWG>
__device__ __forceinline__ WG make_wg(const CtaSetup& s, uint32_t wbytes) {
    WG c;
    c.wg = threadIdx.x >> 7;
    c.t = threadIdx.x & 127;
    c.a_hi = s.w + wbytes + (uint32_t)c.wg * ABUF;
    c.a_lo = c.a_hi + APLANE;
    c.mbar_mma = smem_u32(&s.tail->mbar[1 + 2 * c.wg]);
    c.mbar_tma = smem_u32(&s.tail->mbar[2 + 2 * c.wg]);
    c.ph_mma = 0;
    c.ph_tma = 0;
    c.tmem = s.tmem_base + ((uint32_t)((c.t >> 5) * 32) << 16) + (uint32_t)c.wg * (512u / NWG);
    return c;
}
__device__ __forceinline__ void cta_teardown(const CtaSetup& s) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s.tmem_base, 512);
}
__device__ __forceinline__ void copy_vec(float* dst, const float* __restrict__ src, int n, float scale = ASCALE) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] * scale;     // biases live in the scaled domain
}

// ---- update: 3 GEMMs per tile (see the header) [+ score head] ---------------------------------------------
constexpr int UPD_WG = 4;
constexpr uint32_t UPD_W3 = 0, UPD_WC = 4 * WPLANE, UPD_W42 = 6 * WPLANE, UPD_FN = 8 * WPLANE, UPD_WBYTES = 10 * WPLANE;   // 80 KB

__global__ void __launch_bounds__(128 * UPD_WG, 1) k_tc_update(GnnParams g, int backward, const float* __restrict__ lb,
                                                               const float* __restrict__ ub, const uint16_t* __restrict__ nb_img,
                                                               const float* __restrict__ rlx, float* __restrict__ mu_out,
                                                               float* __restrict__ scores, int n, int64_t score_stride,
                                                               int64_t score_off, int64_t rows, unsigned long long* nan_count) {
    const int l3 = backward ? BC3 : FC3, l4b = backward ? BC4_1 : FC4_2, lc = backward ? T_BWD_C : T_FWD_C;
    const uint16_t* const wsrc[4] = {g.tc[l3], g.tcx_w[lc], g.tc[l4b], g.tc[FNODE]};
    const uint32_t woff[4] = {UPD_W3, UPD_WC, UPD_W42, UPD_FN};
    const uint32_t wlen[4] = {4 * WPLANE, 2 * WPLANE, 2 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<UPD_WG, 4>(UPD_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[l3], P); copy_vec(tl.bias[1], g.tcx_b[lc], P); copy_vec(tl.bias[2], g.bias[l4b], P);
    copy_vec(tl.bias[3], g.bias[FNODE], P); copy_vec(tl.vec, g.wt[FSCORE], P, 1.0f);
    const float bscore = g.bias[FSCORE][0];
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);                      // weight planes have landed
    WG c = make_wg<UPD_WG>(s, UPD_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    bool bad = false;
    const int64_t tile_step = (int64_t)gridDim.x * UPD_WG;
    int64_t tile = (int64_t)blockIdx.x * UPD_WG + c.wg;
    if (tile < ntiles) tma_tile(c, nb_img + (size_t)tile * (ABUF / 2));          // first tile's nb image
    for (; tile < ntiles; tile += tile_step) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        const bool has_next = tile + tile_step < ntiles;
        float l = 0.f, u = 1.f;
        if (grow < rows) { l = ldg1_now(lb + grow); u = ldg1_now(ub + grow); }
        // D[0:128) = nb [W3a; W3b]^T
        gemm<true>(c, W + UPD_W3, W + UPD_W3 + 2 * WPLANE, 128, 0, false);
        const Ratio q = compute_ratio(l, u);
        const float gate = (q.r0 != 0.0f) ? 1.0f : 0.0f;
        // h3 = relu(r0 * D[0:64) + r1 * D[64:128) + b3) -> A   (graph_conv.py:169-170 / 331-336)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float a[16], b[16];
            {
                uint32_t ra[16], rb[16];
                tmem_ld16(c.tmem + qd * 16, ra);
                tmem_ld16(c.tmem + 64 + qd * 16, rb);
                tmem_wait16(ra, a);
                tmem_wait16(rb, b);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 b0 = lds4(tl.bias[0] + qd * 16 + h * 8), b1 = lds4(tl.bias[0] + qd * 16 + h * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = relu_nan(fmaf(q.r0, a[h * 8 + j], fmaf(q.r1, b[h * 8 + j], bb[j])));
                a_store8(c, qd * 2 + h, o);
            }
        }
        // D[0:64) = h3 Wc^T; meanwhile fetch this row's relax' (coalesced by layout)
        gemm_start<false>(c, W + UPD_WC, W + UPD_WC + WPLANE, 64, 0, false);
        float4 rx[16];
        {
            const float* rt = rlx + (size_t)tile * (TILE * P);
#pragma unroll
            for (int i = 0; i < 16; ++i) rx[i] = ldg4_now(rt + ((size_t)i * TILE + c.t) * 4);
        }
        gemm_finish(c);
        // g = relu(D + relax' + bc) -> A
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float v[16];
            tmem_ld16_sync(c.tmem + qd * 16, v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 b0 = lds4(tl.bias[1] + qd * 16 + h * 8), b1 = lds4(tl.bias[1] + qd * 16 + h * 8 + 4);
                const float4 x0 = rx[qd * 4 + h * 2], x1 = rx[qd * 4 + h * 2 + 1];
                float o[8] = {v[h * 8 + 0] + x0.x + b0.x, v[h * 8 + 1] + x0.y + b0.y, v[h * 8 + 2] + x0.z + b0.z, v[h * 8 + 3] + x0.w + b0.w,
                              v[h * 8 + 4] + x1.x + b1.x, v[h * 8 + 5] + x1.y + b1.y, v[h * 8 + 6] + x1.z + b1.z, v[h * 8 + 7] + x1.w + b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = relu_nan(o[j]);
                a_store8(c, qd * 2 + h, o);
            }
        }
        // D[64:128) = g W4_2^T;  mu = (D + b) * (r0 != 0) -> global
        gemm(c, W + UPD_W42, W + UPD_W42 + WPLANE, 64, 64, false);
        // the A buffer is free again: fetch the next tile's nb image while this tile's results leave
        if (scores == nullptr && has_next) tma_tile(c, nb_img + (size_t)(tile + tile_step) * (ABUF / 2));
        bad |= epilogue_to_global_direct(c, 64, tl.bias[2], gate, mu_out, row0, rows);
        if (scores != nullptr) {      // score head on the new embeddings (graph_conv.py:448-449)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float v[16];
                tmem_ld16_sync(c.tmem + 64 + qd * 16, v);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float4 b0 = lds4(tl.bias[2] + qd * 16 + h * 8), b1 = lds4(tl.bias[2] + qd * 16 + h * 8 + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = (v[h * 8 + j] + bb[j]) * gate;
                    a_store8(c, qd * 2 + h, o);
                }
            }
            gemm(c, W + UPD_FN, W + UPD_FN + WPLANE, 64, 0, false);
            if (has_next) tma_tile(c, nb_img + (size_t)(tile + tile_step) * (ABUF / 2));
            float sc = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float v[16];
                tmem_ld16_sync(c.tmem + qd * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sc = fmaf(relu_nan(v[j] + tl.bias[3][qd * 16 + j]), tl.vec[qd * 16 + j], sc);
            }
            if (grow < rows) scores[(grow / n) * score_stride + score_off + (grow % n)] = fmaf(sc, AINV, bscore);
        }
    }
    if (bad) atomicAdd(nan_count, 1ULL);
    cta_teardown(s);
}

// ---- relax: round-independent part of the fc4 / bc4 pre-activation of a hidden layer --------------------------
constexpr int RLX_WG = 2;
constexpr uint32_t RLX_FR = 0, RLX_BC11 = 2 * WPLANE, RLX_BC12 = 4 * WPLANE, RLX_BC2 = 6 * WPLANE, RLX_BR = 12 * WPLANE;
constexpr uint32_t RLX_WBYTES = 14 * WPLANE;
constexpr uint32_t RD1 = 0, RD2 = 192;         // TMEM columns inside a warpgroup's 256

__global__ void __launch_bounds__(128 * RLX_WG, 1) k_tc_relax(GnnParams g, NodeInputs in, float* __restrict__ rlx_f,
                                                              float* __restrict__ rlx_b) {
    const uint16_t* const wsrc[5] = {g.tcx_w[T_FWD_R], g.tc[BC1_1], g.tc[BC1_2], g.tc[BC2], g.tcx_w[T_BWD_R]};
    const uint32_t woff[5] = {RLX_FR, RLX_BC11, RLX_BC12, RLX_BC2, RLX_BR};
    const uint32_t wlen[5] = {2 * WPLANE, 2 * WPLANE, 2 * WPLANE, 6 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<RLX_WG, 5>(RLX_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.tcx_b[T_FWD_R], P); copy_vec(tl.bias[1], g.bias[BC1_1], P); copy_vec(tl.bias[2], g.bias[BC1_2], P);
    copy_vec(tl.bias[3], g.bias[BC2], P); copy_vec(tl.bias[4], g.tcx_b[T_BWD_R], P);
    copy_vec(tl.w_small[0], g.wt[FC1], 7 * P); copy_vec(tl.w_small[1], g.wt[BC1], 7 * P);
    copy_vec(tl.bias[5], g.bias[FC1], P); copy_vec(tl.vec, g.bias[BC1], P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg<RLX_WG>(s, RLX_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (in.rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * RLX_WG + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * RLX_WG) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float l = 0.f, u = 1.f, d1 = 0.f, d2 = 0.f, pp = 0.f, po = 0.f, bs = 0.f;
        if (grow < in.rows) {
            l = in.lb[grow]; u = in.ub[grow];
            d1 = in.dual[grow * 3 + 1]; d2 = in.dual[grow * 3 + 2];
            pp = in.prim_pre[grow]; po = in.prim_post[grow];
            bs = in.bias_node[grow % in.n];
        }
        const Ratio q = compute_ratio(l, u);
        // forward: relax' = amb * (Wr relu(fc1([beta, l, u, d1-d2, x_pre, x_post, bias])) + br)   (graph_conv.py:153-161, :176)
        {
            const float feat[7] = {q.beta, l, u, d1 - d2, pp, po, bs};
            first_layer_to_a<7>(c, feat, tl.w_small[0], tl.bias[5]);
        }
        gemm(c, W + RLX_FR, W + RLX_FR + WPLANE, 64, RD2, false);
        epilogue_to_rlx(c, RD2, tl.bias[0], q.amb, rlx_f + (size_t)tile * (TILE * P));
        // backward (graph_conv.py:273-293, :344)
        {
            const float feat[7] = {l, u, q.beta, -d2 + d1, po, pp, bs};
            first_layer_to_a<7>(c, feat, tl.w_small[1], tl.vec);
        }
        gemm(c, W + RLX_BC11, W + RLX_BC11 + WPLANE, 64, RD2, false);
        epilogue_to_a<true>(c, RD2, tl.bias[1]);
        gemm(c, W + RLX_BC12, W + RLX_BC12 + WPLANE, 64, RD2, false);
        epilogue_to_a<false>(c, RD2, tl.bias[2]);                                  // s1
        gemm(c, W + RLX_BC2, W + RLX_BC2 + 3 * WPLANE, 192, RD1, false);          // s1 [W2a; W2b; W2c]^T
        {   // relu(Da + (-d2) Db + d1 Dc + b2) -> A
            const float nd2 = -d2;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float a[16], b[16], cc[16];
                {
                    uint32_t ra[16], rb[16], rc[16];
                    tmem_ld16(c.tmem + RD1 + qd * 16, ra);
                    tmem_ld16(c.tmem + RD1 + 64 + qd * 16, rb);
                    tmem_ld16(c.tmem + RD1 + 128 + qd * 16, rc);
                    tmem_wait16(ra, a);
                    tmem_wait16(rb, b);
                    tmem_wait16(rc, cc);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = relu_nan(a[h * 8 + j] + fmaf(nd2, b[h * 8 + j], fmaf(d1, cc[h * 8 + j], tl.bias[3][qd * 16 + h * 8 + j])));
                    a_store8(c, qd * 2 + h, o);
                }
            }
        }
        gemm(c, W + RLX_BR, W + RLX_BR + WPLANE, 64, RD2, false);
        epilogue_to_rlx(c, RD2, tl.bias[4], q.amb, rlx_b + (size_t)tile * (TILE * P));
    }
    cta_teardown(s);
}

// ---- input embedding: mu0 = inp_f_1(relu(inp_f([l0, x, u0])))   (graph_conv.py:90-95) ---------------------------
constexpr int EMB_WG = 4;
constexpr uint32_t EMB_WBYTES = 2 * WPLANE;

__global__ void __launch_bounds__(128 * EMB_WG, 1) k_tc_input_embed(GnnParams g, const float* __restrict__ lb0,
                                                                    const float* __restrict__ x, const float* __restrict__ ub0,
                                                                    float* __restrict__ mu0, int64_t rows) {
    const uint16_t* const wsrc[1] = {g.tc[INP_F_1]};
    const uint32_t woff[1] = {0};
    const uint32_t wlen[1] = {2 * WPLANE};
    CtaSetup s = cta_setup<EMB_WG, 1>(EMB_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[INP_F_1], P); copy_vec(tl.bias[5], g.bias[INP_F], P); copy_vec(tl.w_small[0], g.wt[INP_F], 3 * P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg<EMB_WG>(s, EMB_WBYTES);
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * EMB_WG + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * EMB_WG) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float feat[3] = {0.f, 0.f, 0.f};
        if (grow < rows) { feat[0] = lb0[grow]; feat[1] = x[grow]; feat[2] = ub0[grow]; }
        first_layer_to_a<3>(c, feat, tl.w_small[0], tl.bias[5]);
        gemm(c, s.w, s.w + WPLANE, 64, 0, false);
        epilogue_to_global(c, 0, tl.bias[0], 1.0f, mu0, row0, rows);
    }
    cta_teardown(s);
}

// ---- input update: mu0 = inp_b2_2(relu(inp_b2([inp_b_1(relu(inp_b([l0,u0]))), nb])))   (graph_conv.py:380-385) ----
//      = inp_b2_2(relu(Wi relu(inp_b([l0,u0])) + Wn nb + bi)),  Wi = W_b2[:, :64] W_b1,  Wn = W_b2[:, 64:]
constexpr int INU_WG = 4;
constexpr uint32_t INU_WI = 0, INU_WN = 2 * WPLANE, INU_B22 = 4 * WPLANE, INU_WBYTES = 6 * WPLANE;

__global__ void __launch_bounds__(128 * INU_WG, 1) k_tc_input_update(GnnParams g, const float* __restrict__ lb0,
                                                                     const float* __restrict__ ub0, const uint16_t* __restrict__ nb_img,
                                                                     float* __restrict__ mu0, int64_t rows) {
    const uint16_t* const wsrc[3] = {g.tcx_w[T_INP_C], g.tcx_w[T_INP_NB], g.tc[INP_B2_2]};
    const uint32_t woff[3] = {INU_WI, INU_WN, INU_B22};
    const uint32_t wlen[3] = {2 * WPLANE, 2 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<INU_WG, 3>(INU_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.tcx_b[T_INP_C], P); copy_vec(tl.bias[1], g.bias[INP_B2_2], P);
    copy_vec(tl.bias[5], g.bias[INP_B], P); copy_vec(tl.w_small[0], g.wt[INP_B], 2 * P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg<INU_WG>(s, INU_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * INU_WG + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * INU_WG) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float feat[2] = {0.f, 0.f};
        if (grow < rows) { feat[0] = lb0[grow]; feat[1] = ub0[grow]; }
        first_layer_to_a<2>(c, feat, tl.w_small[0], tl.bias[5]);
        gemm(c, W + INU_WI, W + INU_WI + WPLANE, 64, 0, false);                         // Wi relu(inp_b(.))
        fence_proxy_async();
        wg_barrier(c);
        tma_tile(c, nb_img + (size_t)tile * (ABUF / 2));
        gemm<true>(c, W + INU_WN, W + INU_WN + WPLANE, 64, 0, true);                     // + Wn nb
        epilogue_to_a<true>(c, 0, tl.bias[0]);
        gemm(c, W + INU_B22, W + INU_B22 + WPLANE, 64, 64, false);
        epilogue_to_global(c, 64, tl.bias[1], 1.0f, mu0, row0, rows);
    }
    cta_teardown(s);
}

constexpr size_t smem_bytes(uint32_t wbytes, int nwg) { return 1024 + wbytes + nwg * ABUF + sizeof(Tail); }

int grid_for(int64_t rows, int nwg) {
    const int64_t tiles = (rows + TILE - 1) / TILE, ctas = (tiles + nwg - 1) / nwg;
    return (int)(ctas < 1 ? 1 : (ctas < 148 ? ctas : 148));
}

// fp16 tile image of nb (hi + lo planes, scaled domain) -> fp32 [rows][64]; debugging snapshots only
__global__ void k_unpack_tile_image(const uint16_t* __restrict__ img, float* __restrict__ out, int64_t rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (row, 8-channel chunk)
    if (i >= rows * 8) return;
    const int64_t row = i >> 3;
    const int chunk = (int)(i & 7);
    const int64_t tile = row / TILE;
    const uint32_t r = (uint32_t)(row % TILE);
    const __half* hi = reinterpret_cast<const __half*>(img) + tile * (ABUF / 2) + swz(r, (uint32_t)chunk) / 2;
    const __half* lo = hi + APLANE / 2;
    for (int j = 0; j < 8; ++j) out[row * P + chunk * 8 + j] = (__half2float(hi[j]) + __half2float(lo[j])) * AINV;
}

}  // namespace

bool tc_available() { return true; }

int tc_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_tc_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(UPD_WBYTES, UPD_WG))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(RLX_WBYTES, RLX_WG))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_input_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(EMB_WBYTES, EMB_WG))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_input_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(INU_WBYTES, INU_WG))) != cudaSuccess) return e;
    return 0;
}

// packed layout of one linear weight W[64][K], K = 64 * nblk:  [hi plane of K-block 0 .. nblk-1][lo plane 0 .. nblk-1],
// each plane 64 (n) x 64 (k) fp16 in the K-major SWIZZLE_128B shared-memory image (8 KB), so that the planes of
// consecutive K-blocks also read as one (64 * nblk)-row B tile
int64_t tc_packed_elems(int K) { return (int64_t)2 * (K / 64) * 64 * 64; }

int64_t tc_pack_weight(const float* w, int K, uint16_t* dst) {
    const int nblk = K / 64;
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < K; ++k) {
            uint16_t hi, lo;
            split_host(w[(size_t)n * K + k], hi, lo);
            const int kb = k / 64, kk = k % 64;
            const size_t e = (size_t)(swz((uint32_t)n, (uint32_t)(kk / 8)) + (kk % 8) * 2) / 2;
            dst[(size_t)kb * 4096 + e] = hi;
            dst[(size_t)(nblk + kb) * 4096 + e] = lo;
        }
    return tc_packed_elems(K);
}

void tc_relax(const GnnParams& g, const NodeInputs& in, float* relax_f, float* relax_b, cudaStream_t st, int64_t* launches) {
    k_tc_relax<<<grid_for(in.rows, RLX_WG), 128 * RLX_WG, smem_bytes(RLX_WBYTES, RLX_WG), st>>>(g, in, relax_f, relax_b);
    ++*launches;
}

void tc_update(const GnnParams& g, bool backward, const float* lb, const float* ub, const float* nb, const float* relax,
               float* mu_out, float* scores, int n, int64_t score_stride, int64_t score_off, int64_t rows,
               unsigned long long* nan_count, cudaStream_t st, int64_t* launches) {
    k_tc_update<<<grid_for(rows, UPD_WG), 128 * UPD_WG, smem_bytes(UPD_WBYTES, UPD_WG), st>>>(
        g, backward ? 1 : 0, lb, ub, reinterpret_cast<const uint16_t*>(nb), relax, mu_out, scores, n, score_stride, score_off, rows,
        nan_count);
    ++*launches;
}

void tc_input_embed(const GnnParams& g, const float* lb0, const float* x, const float* ub0, float* mu0, int64_t rows,
                    cudaStream_t st, int64_t* launches) {
    k_tc_input_embed<<<grid_for(rows, EMB_WG), 128 * EMB_WG, smem_bytes(EMB_WBYTES, EMB_WG), st>>>(g, lb0, x, ub0, mu0, rows);
    ++*launches;
}

void tc_input_update(const GnnParams& g, const float* lb0, const float* ub0, const float* nb, float* mu0, int64_t rows,
                     cudaStream_t st, int64_t* launches) {
    k_tc_input_update<<<grid_for(rows, INU_WG), 128 * INU_WG, smem_bytes(INU_WBYTES, INU_WG), st>>>(
        g, lb0, ub0, reinterpret_cast<const uint16_t*>(nb), mu0, rows);
    ++*launches;
}

void tc_unpack_tile_image(const float* img, float* out, int64_t rows, cudaStream_t st) {
    const int64_t n = rows * 8;
    k_unpack_tile_image<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(img), out, rows);
}

}  // namespace gnnb
