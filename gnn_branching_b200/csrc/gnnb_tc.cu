// Per-node MLP kernels on the 5th-generation tensor cores (GNNB_MATH_TC_FP16X3) — the product path.
//
// Every dense layer of the GNN with K >= 64 is a [128 nodes x K] x [K x N] GEMM per tile, issued as
// tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 64 / 128 / 192, K = 16 per instruction) with fp32
// accumulators in tensor memory.  fp32 accuracy (scores within 1e-4, BASELINE.json) comes from an fp16 hi/lo
// split of BOTH operands and three MMAs per K step:  x*w ~= xh*wh + xl*wh + xh*wl  (error ~2^-22 per product;
// see gnnb_umma.cuh for the scaled operand domain that keeps fp16 in range).
//
// Structure of a CTA (256 threads = 2 warpgroups, 1 CTA per SM, persistent over tiles):
//   * the stage's weights sit in shared memory for the CTA's lifetime as fp16 hi/lo planes in the UMMA K-major
//     SWIZZLE_128B layout; they are repacked once on the host (tc_pack_weight) so one cp.async.bulk per linear
//     (TMA, 1-D) lands them;
//   * each warpgroup owns one 128-node tile at a time, its own 32 KB A-operand buffer (hi + lo plane), 256 TMEM
//     columns and one mbarrier, and walks the layer chain of its tile sequentially:
//       write A (thread = node row) -> fence.proxy.async -> warpgroup barrier -> one thread issues the MMAs and
//       tcgen05.commit -> everyone waits on the mbarrier -> tcgen05.ld the accumulators -> bias / ReLU / row
//       scaling in registers -> hi/lo split -> next A ...
//     the two warpgroups are independent, so one tile's epilogue overlaps the other tile's MMAs;
//   * K < 64 first layers (7, 3, 2 input features) run on CUDA cores in fp32, exactly;
//   * row scalings that the reference applies to GEMM *inputs* move to the epilogue by linearity (SURVEY §8a
//     fact 3): [nb*r0, nb*r1] W3^T = r0 (nb W3a^T) + r1 (nb W3b^T) is one N = 128 MMA, and bc2's
//     [s1, -d2 s1, d1 s1] input one N = 192 MMA.
//
// Stage contracts are those of the SIMT twins in gnnb_simt.cu (same inputs, outputs, reference citations).
#include "gnnb_umma.cuh"

namespace gnnb {
namespace {

using namespace tcx;

// ---- warpgroup context ------------------------------------------------------------------------------
struct WG {
    uint32_t a_hi, a_lo;        // shared addresses of this warpgroup's A planes (a_lo = a_hi + APLANE)
    uint32_t mbar;              // this warpgroup's MMA-completion mbarrier
    uint32_t phase;
    uint32_t tmem;              // TMEM address: lane base of this warp, first column of this warpgroup
    int t;                      // thread index within the warpgroup = row of the tile this thread owns
    int wg;
};

__device__ __forceinline__ void wg_barrier(const WG& c) { named_bar(1 + c.wg, 128); }

// make this thread's A-plane writes visible to the tensor core, then one thread issues 3 x 4 MMAs + commit
__device__ __forceinline__ void gemm_start(const WG& c, uint32_t b_hi, uint32_t b_lo, uint32_t N, uint32_t dcol, bool accumulate) {
    fence_proxy_async();
    tc_fence_before();
    wg_barrier(c);
    if (c.t == 0) {
        tc_fence_after();
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
        umma_commit(c.mbar);
    }
}
__device__ __forceinline__ void gemm_finish(WG& c) {
    mbar_wait(c.mbar, c.phase);
    c.phase ^= 1u;
    tc_fence_after();
}
__device__ __forceinline__ void gemm(WG& c, uint32_t b_hi, uint32_t b_lo, uint32_t N, uint32_t dcol, bool accumulate) {
    gemm_start(c, b_hi, b_lo, N, dcol, accumulate);
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

// a [rows][64] fp32 tile in global memory -> A planes, coalesced (a warp instruction covers two 256-byte rows)
struct TileRegs { float4 v[16]; };
__device__ __forceinline__ void tile_load(const WG& c, const float* __restrict__ src, int64_t row0, int64_t rows, TileRegs& r) {
    const int lane = c.t & 31, warp = c.t >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int rr = warp * 32 + 2 * i + (lane >> 4);
        const int64_t grow = row0 + rr;
        r.v[i] = (grow < rows) ? __ldg(reinterpret_cast<const float4*>(src + grow * P) + (lane & 15)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__device__ __forceinline__ void tile_to_a(const WG& c, const TileRegs& r) {
    const int lane = c.t & 31, warp = c.t >> 5;
    const int c4 = lane & 15;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int rr = warp * 32 + 2 * i + (lane >> 4);
        uint32_t h0, h1, l0, l1;
        split2(r.v[i].x * ASCALE, r.v[i].y * ASCALE, h0, l0);
        split2(r.v[i].z * ASCALE, r.v[i].w * ASCALE, h1, l1);
        const uint32_t off = swz((uint32_t)rr, (uint32_t)(c4 >> 1)) + (uint32_t)(c4 & 1) * 8u;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(c.a_hi + off), "r"(h0), "r"(h1) : "memory");
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(c.a_lo + off), "r"(l0), "r"(l1) : "memory");
    }
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
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = v[h * 8 + j] + bias_s[q * 16 + h * 8 + j];
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
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch * 8), b1 = *reinterpret_cast<const float4*>(bias_s + ch * 8 + 4);
        o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 w0 = *reinterpret_cast<const float4*>(wt_s + k * P + ch * 8), w1 = *reinterpret_cast<const float4*>(wt_s + k * P + ch * 8 + 4);
            o[0] = fmaf(feat[k], w0.x, o[0]); o[1] = fmaf(feat[k], w0.y, o[1]); o[2] = fmaf(feat[k], w0.z, o[2]); o[3] = fmaf(feat[k], w0.w, o[3]);
            o[4] = fmaf(feat[k], w1.x, o[4]); o[5] = fmaf(feat[k], w1.y, o[5]); o[6] = fmaf(feat[k], w1.z, o[6]); o[7] = fmaf(feat[k], w1.w, o[7]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = relu_nan(o[j]);
        a_store8(c, ch, o);
    }
}

// (accumulator columns [dcol, dcol+64) + bias) * rowscale -> global [rows][64], staged through this warpgroup's A
// buffer (free at this point) so that the global stores are full 256-byte rows.  Returns true if a NaN was written.
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
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = (v[h * 4 + j] + bias_s[q * 16 + h * 4 + j]) * rowscale;
                bad |= (o[j] != o[j]);
            }
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
    wg_barrier(c);             // staging fully read before the A buffer is written again
    return bad && (row0 + c.t < rows);
}

// ---- shared-memory layout -----------------------------------------------------------------------------
struct Tail {                   // small fp32 data after the weight planes and A buffers
    float bias[6][P];
    float w_small[2][8 * P];    // first-layer weights (K <= 7), transposed [K][64]
    float vec[P];               // fscore weights
    uint64_t mbar[1 + WGS];
    uint32_t tmem_slot;
};

struct CtaSetup {
    unsigned char* base;        // 1024-aligned dynamic shared memory
    uint32_t w;                 // shared address of the weight planes
    Tail* tail;
    uint32_t tmem_base;
};

// common prologue: carve shared memory, allocate TMEM, init mbarriers, TMA the weight planes in
template <int NW>
__device__ __forceinline__ CtaSetup cta_setup(uint32_t wbytes, const uint16_t* const (&wsrc)[NW], const uint32_t (&woff)[NW],
                                              const uint32_t (&wlen)[NW]) {
    extern __shared__ unsigned char smem_raw[];
    CtaSetup s;
    s.base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    s.w = smem_u32(s.base);
    s.tail = reinterpret_cast<Tail*>(s.base + wbytes + WGS * ABUF);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + WGS; ++i) mbar_init(smem_u32(&s.tail->mbar[i]), 1);
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
__device__ __forceinline__ WG make_wg(const CtaSetup& s, uint32_t wbytes) {
    WG c;
    c.wg = threadIdx.x >> 7;
    c.t = threadIdx.x & 127;
    c.a_hi = s.w + wbytes + (uint32_t)c.wg * ABUF;
    c.a_lo = c.a_hi + APLANE;
    c.mbar = smem_u32(&s.tail->mbar[1 + c.wg]);
    c.phase = 0;
    c.tmem = s.tmem_base + ((uint32_t)((c.t >> 5) * 32) << 16) + (uint32_t)c.wg * 256u;
    return c;
}
__device__ __forceinline__ void cta_teardown(const CtaSetup& s) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s.tmem_base, 512);
}
__device__ __forceinline__ void copy_vec(float* dst, const float* __restrict__ src, int n, float scale = ASCALE) {
    for (int i = threadIdx.x; i < n; i += NTHREADS) dst[i] = src[i] * scale;     // biases live in the scaled domain
}

// TMEM column map inside a warpgroup's 256 columns
constexpr uint32_t D1 = 0;      // up to 192 columns (fused N = 128 / 192 products)
constexpr uint32_t D2 = 192;    // 64 columns

// ---- update: e = W3b(relu(r0 nb W3a0^T + r1 nb W3a1^T + b)), mu = W4b(relu([relax, e] W4^T + b)) * (r0 != 0) [+ score] ----
constexpr uint32_t UPD_W3 = 0, UPD_W3B = 4 * WPLANE, UPD_W4 = 6 * WPLANE, UPD_W4B = 10 * WPLANE, UPD_FN = 12 * WPLANE;
constexpr uint32_t UPD_WBYTES = 14 * WPLANE;    // 112 KB

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_update(GnnParams g, int backward, const float* __restrict__ lb,
                                                           const float* __restrict__ ub, const float* __restrict__ nb,
                                                           const float* __restrict__ relax, float* __restrict__ mu_out,
                                                           float* __restrict__ scores, int n, int64_t score_stride,
                                                           int64_t score_off, int64_t rows, unsigned long long* nan_count) {
    const int l3 = backward ? BC3 : FC3, l3b = backward ? BC3_1 : FC3_2, l4 = backward ? BC4 : FC4, l4b = backward ? BC4_1 : FC4_2;
    const uint16_t* const wsrc[5] = {g.tc[l3], g.tc[l3b], g.tc[l4], g.tc[l4b], g.tc[FNODE]};
    const uint32_t woff[5] = {UPD_W3, UPD_W3B, UPD_W4, UPD_W4B, UPD_FN};
    const uint32_t wlen[5] = {4 * WPLANE, 2 * WPLANE, 4 * WPLANE, 2 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<5>(UPD_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[l3], P); copy_vec(tl.bias[1], g.bias[l3b], P); copy_vec(tl.bias[2], g.bias[l4], P);
    copy_vec(tl.bias[3], g.bias[l4b], P); copy_vec(tl.bias[4], g.bias[FNODE], P); copy_vec(tl.vec, g.wt[FSCORE], P, 1.0f);
    const float bscore = g.bias[FSCORE][0];
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);                      // weight planes have landed
    WG c = make_wg(s, UPD_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    bool bad = false;
    for (int64_t tile = (int64_t)blockIdx.x * WGS + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * WGS) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float l = 0.f, u = 1.f;
        if (grow < rows) { l = lb[grow]; u = ub[grow]; }
        const Ratio q = compute_ratio(l, u);
        const float gate = (q.r0 != 0.0f) ? 1.0f : 0.0f;
        TileRegs tr;
        // relax -> A;  D2 = relax W4[:, :64]^T  (runs while the nb tile is fetched)
        tile_load(c, relax, row0, rows, tr);
        tile_to_a(c, tr);
        gemm_start(c, W + UPD_W4, W + UPD_W4 + 2 * WPLANE, 64, D2, false);
        tile_load(c, nb, row0, rows, tr);
        gemm_finish(c);
        // nb -> A;  D1[0:128) = nb [W3a0; W3a1]^T
        tile_to_a(c, tr);
        gemm(c, W + UPD_W3, W + UPD_W3 + 2 * WPLANE, 128, D1, false);
        // h = relu(r0 * D1[0:64) + r1 * D1[64:128) + b3) -> A   (graph_conv.py:169-170 / 331-336)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float a[16], b[16];
            {
                uint32_t ra[16], rb[16];
                tmem_ld16(c.tmem + D1 + qd * 16, ra);
                tmem_ld16(c.tmem + D1 + 64 + qd * 16, rb);
                tmem_wait16(ra, a);
                tmem_wait16(rb, b);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = relu_nan(fmaf(q.r0, a[h * 8 + j], fmaf(q.r1, b[h * 8 + j], tl.bias[0][qd * 16 + h * 8 + j])));
                a_store8(c, qd * 2 + h, o);
            }
        }
        gemm(c, W + UPD_W3B, W + UPD_W3B + WPLANE, 64, D1, false);
        // e = D1 + b -> A;  D2 += e W4[:, 64:]^T
        epilogue_to_a<false>(c, D1, tl.bias[1]);
        gemm(c, W + UPD_W4 + WPLANE, W + UPD_W4 + 3 * WPLANE, 64, D2, true);
        // relu(D2 + b4) -> A;  D1 = . W4b^T
        epilogue_to_a<true>(c, D2, tl.bias[2]);
        gemm(c, W + UPD_W4B, W + UPD_W4B + WPLANE, 64, D1, false);
        // mu = (D1 + b) * (r0 != 0) -> global
        bad |= epilogue_to_global(c, D1, tl.bias[3], gate, mu_out, row0, rows);
        if (scores != nullptr) {      // score head on the new embeddings (graph_conv.py:448-449)
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float v[16];
                tmem_ld16_sync(c.tmem + D1 + qd * 16, v);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = (v[h * 8 + j] + tl.bias[3][qd * 16 + h * 8 + j]) * gate;
                    a_store8(c, qd * 2 + h, o);
                }
            }
            gemm(c, W + UPD_FN, W + UPD_FN + WPLANE, 64, D2, false);
            float sc = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float v[16];
                tmem_ld16_sync(c.tmem + D2 + qd * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) sc = fmaf(relu_nan(v[j] + tl.bias[4][qd * 16 + j]), tl.vec[qd * 16 + j], sc);
            }
            if (grow < rows) scores[(grow / n) * score_stride + score_off + (grow % n)] = fmaf(sc, AINV, bscore);
            tc_fence_before();
            wg_barrier(c);            // all reads of D2 are done before the next tile's first MMA overwrites it
        }
    }
    if (bad) atomicAdd(nan_count, 1ULL);
    cta_teardown(s);
}

// ---- relax: round-independent relaxation features of a hidden layer ------------------------------------------
constexpr uint32_t RLX_FC11 = 0, RLX_BC11 = 2 * WPLANE, RLX_BC12 = 4 * WPLANE, RLX_BC2 = 6 * WPLANE, RLX_BC21 = 12 * WPLANE;
constexpr uint32_t RLX_WBYTES = 14 * WPLANE;

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_relax(GnnParams g, NodeInputs in, float* __restrict__ relax_f,
                                                          float* __restrict__ relax_b) {
    const uint16_t* const wsrc[5] = {g.tc[FC1_1], g.tc[BC1_1], g.tc[BC1_2], g.tc[BC2], g.tc[BC2_1]};
    const uint32_t woff[5] = {RLX_FC11, RLX_BC11, RLX_BC12, RLX_BC2, RLX_BC21};
    const uint32_t wlen[5] = {2 * WPLANE, 2 * WPLANE, 2 * WPLANE, 6 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<5>(RLX_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[FC1_1], P); copy_vec(tl.bias[1], g.bias[BC1_1], P); copy_vec(tl.bias[2], g.bias[BC1_2], P);
    copy_vec(tl.bias[3], g.bias[BC2], P); copy_vec(tl.bias[4], g.bias[BC2_1], P);
    copy_vec(tl.w_small[0], g.wt[FC1], 7 * P); copy_vec(tl.w_small[1], g.wt[BC1], 7 * P);
    copy_vec(tl.bias[5], g.bias[FC1], P); copy_vec(tl.vec, g.bias[BC1], P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg(s, RLX_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (in.rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * WGS + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * WGS) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float l = 0.f, u = 1.f, d1 = 0.f, d2 = 0.f, pp = 0.f, po = 0.f, bs = 0.f;
        if (grow < in.rows) {
            l = in.lb[grow]; u = in.ub[grow];
            d1 = in.dual[grow * 3 + 1]; d2 = in.dual[grow * 3 + 2];
            pp = in.prim_pre[grow]; po = in.prim_post[grow];
            bs = in.bias_node[grow % in.n];
        }
        const Ratio q = compute_ratio(l, u);
        // forward branch: relax_f = fc1_1(relu(fc1([beta, l, u, d1-d2, x_pre, x_post, bias]))) * amb   (graph_conv.py:153-161)
        {
            const float feat[7] = {q.beta, l, u, d1 - d2, pp, po, bs};
            first_layer_to_a<7>(c, feat, tl.w_small[0], tl.bias[5]);
        }
        gemm(c, W + RLX_FC11, W + RLX_FC11 + WPLANE, 64, D2, false);
        epilogue_to_global(c, D2, tl.bias[0], q.amb, relax_f, row0, in.rows);
        // backward branch (graph_conv.py:273-293)
        {
            const float feat[7] = {l, u, q.beta, -d2 + d1, po, pp, bs};
            first_layer_to_a<7>(c, feat, tl.w_small[1], tl.vec);
        }
        gemm(c, W + RLX_BC11, W + RLX_BC11 + WPLANE, 64, D2, false);
        epilogue_to_a<true>(c, D2, tl.bias[1]);
        gemm(c, W + RLX_BC12, W + RLX_BC12 + WPLANE, 64, D2, false);
        epilogue_to_a<false>(c, D2, tl.bias[2]);                                  // s1
        gemm(c, W + RLX_BC2, W + RLX_BC2 + 3 * WPLANE, 192, D1, false);          // s1 [W2a; W2b; W2c]^T
        {   // relu(Da + (-d2) Db + d1 Dc + b2) -> A
            const float nd2 = -d2;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float a[16], b[16], cc[16];
                {
                    uint32_t ra[16], rb[16], rc[16];
                    tmem_ld16(c.tmem + D1 + qd * 16, ra);
                    tmem_ld16(c.tmem + D1 + 64 + qd * 16, rb);
                    tmem_ld16(c.tmem + D1 + 128 + qd * 16, rc);
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
        gemm(c, W + RLX_BC21, W + RLX_BC21 + WPLANE, 64, D2, false);
        epilogue_to_global(c, D2, tl.bias[4], q.amb, relax_b, row0, in.rows);
    }
    cta_teardown(s);
}

// ---- input embedding: mu0 = inp_f_1(relu(inp_f([l0, x, u0])))   (graph_conv.py:90-95) ---------------------------
constexpr uint32_t EMB_WBYTES = 2 * WPLANE;

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_input_embed(GnnParams g, const float* __restrict__ lb0,
                                                                const float* __restrict__ x, const float* __restrict__ ub0,
                                                                float* __restrict__ mu0, int64_t rows) {
    const uint16_t* const wsrc[1] = {g.tc[INP_F_1]};
    const uint32_t woff[1] = {0};
    const uint32_t wlen[1] = {2 * WPLANE};
    CtaSetup s = cta_setup<1>(EMB_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[INP_F_1], P); copy_vec(tl.bias[5], g.bias[INP_F], P); copy_vec(tl.w_small[0], g.wt[INP_F], 3 * P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg(s, EMB_WBYTES);
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * WGS + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * WGS) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float feat[3] = {0.f, 0.f, 0.f};
        if (grow < rows) { feat[0] = lb0[grow]; feat[1] = x[grow]; feat[2] = ub0[grow]; }
        first_layer_to_a<3>(c, feat, tl.w_small[0], tl.bias[5]);
        gemm(c, s.w, s.w + WPLANE, 64, D2, false);
        epilogue_to_global(c, D2, tl.bias[0], 1.0f, mu0, row0, rows);
    }
    cta_teardown(s);
}

// ---- input update: mu0 = inp_b2_2(relu(inp_b2([inp_b_1(relu(inp_b([l0,u0]))), nb])))   (graph_conv.py:380-385) ----
constexpr uint32_t INU_B1 = 0, INU_B2 = 2 * WPLANE, INU_B22 = 6 * WPLANE, INU_WBYTES = 8 * WPLANE;

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_input_update(GnnParams g, const float* __restrict__ lb0,
                                                                 const float* __restrict__ ub0, const float* __restrict__ nb,
                                                                 float* __restrict__ mu0, int64_t rows) {
    const uint16_t* const wsrc[3] = {g.tc[INP_B_1], g.tc[INP_B2], g.tc[INP_B2_2]};
    const uint32_t woff[3] = {INU_B1, INU_B2, INU_B22};
    const uint32_t wlen[3] = {2 * WPLANE, 4 * WPLANE, 2 * WPLANE};
    CtaSetup s = cta_setup<3>(INU_WBYTES, wsrc, woff, wlen);
    Tail& tl = *s.tail;
    copy_vec(tl.bias[0], g.bias[INP_B_1], P); copy_vec(tl.bias[1], g.bias[INP_B2], P); copy_vec(tl.bias[2], g.bias[INP_B2_2], P);
    copy_vec(tl.bias[5], g.bias[INP_B], P); copy_vec(tl.w_small[0], g.wt[INP_B], 2 * P);
    __syncthreads();
    mbar_wait(smem_u32(&tl.mbar[0]), 0);
    WG c = make_wg(s, INU_WBYTES);
    const uint32_t W = s.w;
    const int64_t ntiles = (rows + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * WGS + c.wg; tile < ntiles; tile += (int64_t)gridDim.x * WGS) {
        const int64_t row0 = tile * TILE, grow = row0 + c.t;
        float feat[2] = {0.f, 0.f};
        if (grow < rows) { feat[0] = lb0[grow]; feat[1] = ub0[grow]; }
        TileRegs tr;
        tile_load(c, nb, row0, rows, tr);
        first_layer_to_a<2>(c, feat, tl.w_small[0], tl.bias[5]);
        gemm(c, W + INU_B1, W + INU_B1 + WPLANE, 64, D2, false);
        epilogue_to_a<false>(c, D2, tl.bias[0]);                                          // inp_relax
        gemm(c, W + INU_B2, W + INU_B2 + 2 * WPLANE, 64, D1, false);                      // inp_relax W[:, :64]^T
        tile_to_a(c, tr);
        gemm(c, W + INU_B2 + WPLANE, W + INU_B2 + 3 * WPLANE, 64, D1, true);              // + nb W[:, 64:]^T
        epilogue_to_a<true>(c, D1, tl.bias[1]);
        gemm(c, W + INU_B22, W + INU_B22 + WPLANE, 64, D2, false);
        epilogue_to_global(c, D2, tl.bias[2], 1.0f, mu0, row0, rows);
    }
    cta_teardown(s);
}

constexpr size_t smem_bytes(uint32_t wbytes) { return 1024 + wbytes + WGS * ABUF + sizeof(Tail); }

int grid_for(int64_t rows) {
    const int64_t tiles = (rows + TILE - 1) / TILE, ctas = (tiles + WGS - 1) / WGS;
    return (int)(ctas < 1 ? 1 : (ctas < 148 ? ctas : 148));
}

}  // namespace

bool tc_available() { return true; }

int tc_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_tc_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(UPD_WBYTES))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(RLX_WBYTES))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_input_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(EMB_WBYTES))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_input_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(INU_WBYTES))) != cudaSuccess) return e;
    return 0;
}

// packed layout of one nn.Linear weight W[64][K], K = 64 * nblk:  [hi plane of K-block 0 .. nblk-1][lo plane 0 .. nblk-1],
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
    k_tc_relax<<<grid_for(in.rows), NTHREADS, smem_bytes(RLX_WBYTES), st>>>(g, in, relax_f, relax_b);
    ++*launches;
}

void tc_update(const GnnParams& g, bool backward, const float* lb, const float* ub, const float* nb, const float* relax,
               float* mu_out, float* scores, int n, int64_t score_stride, int64_t score_off, int64_t rows,
               unsigned long long* nan_count, cudaStream_t st, int64_t* launches) {
    k_tc_update<<<grid_for(rows), NTHREADS, smem_bytes(UPD_WBYTES), st>>>(g, backward ? 1 : 0, lb, ub, nb, relax, mu_out, scores,
                                                                          n, score_stride, score_off, rows, nan_count);
    ++*launches;
}

void tc_input_embed(const GnnParams& g, const float* lb0, const float* x, const float* ub0, float* mu0, int64_t rows,
                    cudaStream_t st, int64_t* launches) {
    k_tc_input_embed<<<grid_for(rows), NTHREADS, smem_bytes(EMB_WBYTES), st>>>(g, lb0, x, ub0, mu0, rows);
    ++*launches;
}

void tc_input_update(const GnnParams& g, const float* lb0, const float* ub0, const float* nb, float* mu0, int64_t rows,
                     cudaStream_t st, int64_t* launches) {
    k_tc_input_update<<<grid_for(rows), NTHREADS, smem_bytes(INU_WBYTES), st>>>(g, lb0, ub0, nb, mu0, rows);
    ++*launches;
}

}  // namespace gnnb
