// Propagation of the node embeddings through the verified network's edge sets on the tensor cores.
//
// Every edge set (conv2d, conv_transpose2d / freq, W @ mu, W^T @ mu; graph_conv.py:110-132, 299-322, 361-376) is a
// sparse matrix product  nb[b] = Wmat * mu[b]  with Wmat [n_out x n_in] built from the verified network's own
// weights.  At gnnb_set_network the matrix is cut into dense blocks ("plan"): a tile is 128 output nodes chosen so
// that they share inputs (all channels of a small spatial patch), its K dimension is the sorted list of input nodes
// any of them touches, padded to chunks of 64, and its values are stored as fp16 hi/lo planes in the UMMA K-major
// SWIZZLE_128B image.  conv_transpose's 1/freq tap-count normalisation is folded into the rows.
//
// The kernel is a gather-GEMM: a warpgroup (3 per CTA) owns (tile, pair of subdomains); per K chunk it
//   * TMA-loads the 32 KB weight block (A operand, K-major) with one cp.async.bulk,
//   * gathers the 64 input-node rows of both subdomains (256 B each, coalesced), splits them into fp16 hi/lo and
//     writes them as the MN-major B operand  [k][(subdomain, channel)]  (N = 128),
//   * issues 3 x ksteps tcgen05.mma (M = 128, N = 128, K = 16) into 128 TMEM columns,
// then scatters the accumulator rows (one output node per TMEM lane) into nb, already split into the fp16 hi/lo
// A-operand tile image the node-update kernels load with one cp.async.bulk.
#include <algorithm>
#include <functional>
#include <map>
#include <vector>

#include "gnnb_umma.cuh"

namespace gnnb {

using namespace tcx;

struct PropPlanDev {
    const int32_t* out_rows;      // [ntiles][128] output node of each tile row, -1 = unused
    const int32_t* in_rows;       // [nchunks][64] input node of each K row, -1 = zero row
    const uint16_t* a_planes;     // [nchunks][2][128 * 64] hi plane then lo plane, 32 KB per chunk
    const int32_t* tile_chunk0;   // [ntiles + 1] first chunk of each tile
    const int32_t* ksteps;        // [nchunks] K = 16 steps that hold data (1..4)
    int ntiles, n_in, n_out;
};

struct PropPlan {
    PropPlanDev dev{};
    void* blob = nullptr;
    int nchunks = 0;
    double density = 0.0;         // useful MACs / issued MACs
};

namespace {

constexpr uint32_t PROP_A_BYTES = 2 * APLANE;        // 32 KB: hi + lo weight block of one chunk
constexpr uint32_t PROP_B_BYTES = 2 * 2 * 64 * 128;  // 32 KB: [plane][subdomain][64 rows x 128 B]
constexpr uint32_t PROP_WG_BYTES = PROP_A_BYTES + PROP_B_BYTES;

struct PropTail {
    uint64_t mbar[2 * 4];
    uint32_t tmem_slot;
};

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) rows along N, 8 K-rows per 1024 B group (SBO), the second
// 64-wide N block (second subdomain) `lbo_bytes` further (cute::UMMA::make_umma_desc<Major::MN>)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int PWG = 3;                               // warpgroups per CTA: 3 x (32 KB weights + 32 KB gathered rows)

__device__ __forceinline__ float4 ldg4_now(const float* p) {      // asm volatile: issued where written, never sunk
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(128 * PWG, 1) k_tc_prop(PropPlanDev plan, const float* __restrict__ mu_in,
                                                          uint16_t* __restrict__ nb_img, int Bc) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    PropTail* tail = reinterpret_cast<PropTail*>(base + PWG * PROP_WG_BYTES);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * PWG; ++i) mbar_init(smem_u32(&tail->mbar[i]), 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&tail->tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_slot;

    const int wg = threadIdx.x >> 7, t = threadIdx.x & 127, lane = t & 31, warp = t >> 5;
    const uint32_t a_hi = smem_u32(base) + (uint32_t)wg * PROP_WG_BYTES, a_lo = a_hi + APLANE;
    const uint32_t b_hi = a_hi + PROP_A_BYTES, b_lo = b_hi + PROP_B_BYTES / 2;
    const uint32_t mbar_a = smem_u32(&tail->mbar[2 * wg]), mbar_d = smem_u32(&tail->mbar[2 * wg + 1]);
    const uint32_t tmem = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)wg * 128u;
    const uint32_t tmem_d = (tmem_base & 0x0000FFFFu) + (uint32_t)wg * 128u;
    // D fp32, A fp16 K-major, B fp16 MN-major (bit 16), M = 128, N = 128
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((128u >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    uint32_t ph_a = 0, ph_d = 0;

    const int npairs = (Bc + 1) >> 1;
    const int64_t nitems = (int64_t)plan.ntiles * npairs, item_step = (int64_t)gridDim.x * PWG;

    // the walk over (item = (tile, subdomain pair), K chunk) is flattened so that the rows of the NEXT chunk are
    // requested before waiting for the current chunk's MMAs, across item boundaries too
    int64_t item = (int64_t)blockIdx.x * PWG + wg;
    int tile = 0, d0 = 0, ch = 0, ch0 = 0, ch1 = 0;
    auto open_item = [&]() {
        tile = (int)(item / npairs);
        d0 = 2 * (int)(item % npairs);
        ch0 = plan.tile_chunk0[tile];
        ch1 = plan.tile_chunk0[tile + 1];
        ch = ch0;
    };
    float4 v[16];
    auto gather_issue = [&](int g_ch, int g_d0) {        // 64 input-node rows x 2 subdomains, two 256-byte rows per warp load
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int rr = warp * 32 + 2 * i + (lane >> 4);
            const int dom = rr >> 6, k = rr & 63;
            const int idx = __ldg(plan.in_rows + (size_t)g_ch * 64 + k);
            const int d = g_d0 + dom;
            if (idx >= 0 && d < Bc) v[i] = ldg4_now(mu_in + ((int64_t)d * plan.n_in + idx) * P + (lane & 15) * 4);
            else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (item < nitems) { open_item(); gather_issue(ch, d0); }
    while (item < nitems) {
        if (t == 0) {   // weight block of this chunk (the previous chunk's MMAs have completed: A buffer is free)
            mbar_expect_tx(mbar_a, PROP_A_BYTES);
            bulk_g2s(a_hi, plan.a_planes + (size_t)ch * (PROP_A_BYTES / 2), PROP_A_BYTES, mbar_a);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {                   // gathered rows -> B planes (MN-major image)
            const int rr = warp * 32 + 2 * i + (lane >> 4);
            const int dom = rr >> 6, k = rr & 63, c4 = lane & 15;
            uint32_t h0, h1, l0, l1;
            split2(v[i].x * ASCALE, v[i].y * ASCALE, h0, l0);
            split2(v[i].z * ASCALE, v[i].w * ASCALE, h1, l1);
            const uint32_t off = (uint32_t)dom * 8192u + swz((uint32_t)k, (uint32_t)(c4 >> 1)) + (uint32_t)(c4 & 1) * 8u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(b_hi + off), "r"(h0), "r"(h1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(b_lo + off), "r"(l0), "r"(l1) : "memory");
        }
        fence_proxy_async();
        tc_fence_before();
        named_bar(1 + wg, 128);
        if (t == 0) {
            tc_fence_after();
            mbar_wait(mbar_a, ph_a);
            const int ks_n = plan.ksteps[ch];
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint64_t ad = make_desc(pass == 1 ? a_lo : a_hi);
                const uint64_t bd = make_desc_mn(pass == 2 ? b_lo : b_hi, 8192u);
                for (int ks = 0; ks < ks_n; ++ks)
                    umma(tmem_d, ad + 2 * ks, bd + 128 * ks, idesc, (ch > ch0 || pass > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(mbar_d);
        }
        ph_a ^= 1u;
        // what comes next, and its rows (in flight while the MMAs run)
        const bool last_chunk = (ch + 1 == ch1);
        const int e_tile = tile, e_d0 = d0;
        if (!last_chunk) {
            ++ch;
        } else {
            item += item_step;
            if (item < nitems) open_item();
        }
        if (item < nitems) gather_issue(ch, d0);
        mbar_wait(mbar_d, ph_d);
        ph_d ^= 1u;
        tc_fence_after();
        if (!last_chunk) continue;
        // epilogue: TMEM lane = output node of the tile, columns [64 * dom, 64 * dom + 64) = its channels for subdomain dom.
        // nb leaves as the A-operand image the node kernels TMA in: per 128 consecutive global rows a hi and a lo
        // fp16 plane in the K-major SWIZZLE_128B layout, still in the scaled domain (no split work left for the consumer).
        const int orow = __ldg(plan.out_rows + (size_t)e_tile * TILE + t);
#pragma unroll 1
        for (int dom = 0; dom < 2; ++dom) {
            if (e_d0 + dom >= Bc) break;
            const int64_t grow = (int64_t)(e_d0 + dom) * plan.n_out + orow;
            unsigned char* img = reinterpret_cast<unsigned char*>(nb_img) + (grow / TILE) * (int64_t)ABUF;
            const uint32_t ur = (uint32_t)(grow % TILE);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float x[16];
                tmem_ld16_sync(tmem + (uint32_t)(dom * 64 + q * 16), x);
                if (orow >= 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 hi, lo;
                        split2(x[h * 8 + 0], x[h * 8 + 1], hi.x, lo.x);
                        split2(x[h * 8 + 2], x[h * 8 + 3], hi.y, lo.y);
                        split2(x[h * 8 + 4], x[h * 8 + 5], hi.z, lo.z);
                        split2(x[h * 8 + 6], x[h * 8 + 7], hi.w, lo.w);
                        const uint32_t off = swz(ur, (uint32_t)(q * 2 + h));
                        *reinterpret_cast<uint4*>(img + off) = hi;
                        *reinterpret_cast<uint4*>(img + APLANE + off) = lo;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

constexpr size_t PROP_SMEM = 1024 + PWG * PROP_WG_BYTES + sizeof(PropTail);

// ---- host-side plan construction ------------------------------------------------------------------------
struct Edge { int in; float w; };
using EdgeFn = std::function<void(int, std::vector<Edge>&)>;

// tiles of <= 128 nodes on a (C, H, W) grid: all channels (in blocks of <= 128) of a spatial patch
std::vector<std::vector<int>> grid_tiles(int C, int H, int W) {
    const int Cb = std::min(C, TILE), Pp = std::max(1, TILE / Cb);
    int bestY = 1, bestX = 1;
    long bestTiles = -1;
    for (int Yb = 1; Yb <= std::min(H, Pp); ++Yb)
        for (int Xb = 1; Xb <= std::min(W, Pp / Yb); ++Xb) {
            const long tiles = (long)((H + Yb - 1) / Yb) * ((W + Xb - 1) / Xb);
            const bool better = bestTiles < 0 || tiles < bestTiles ||
                                (tiles == bestTiles && std::abs(Yb - Xb) < std::abs(bestY - bestX));
            if (better) { bestTiles = tiles; bestY = Yb; bestX = Xb; }
        }
    std::vector<std::vector<int>> tiles;
    for (int c0 = 0; c0 < C; c0 += Cb)
        for (int y0 = 0; y0 < H; y0 += bestY)
            for (int x0 = 0; x0 < W; x0 += bestX) {
                std::vector<int> rows;
                for (int c = c0; c < std::min(C, c0 + Cb); ++c)
                    for (int y = y0; y < std::min(H, y0 + bestY); ++y)
                        for (int x = x0; x < std::min(W, x0 + bestX); ++x) rows.push_back((c * H + y) * W + x);
                tiles.push_back(rows);
            }
    return tiles;
}

std::vector<std::vector<int>> range_tiles(int n) {
    std::vector<std::vector<int>> tiles;
    for (int r0 = 0; r0 < n; r0 += TILE) {
        std::vector<int> rows;
        for (int r = r0; r < std::min(n, r0 + TILE); ++r) rows.push_back(r);
        tiles.push_back(rows);
    }
    return tiles;
}

PropPlan* build_plan(const std::vector<std::vector<int>>& tiles, const EdgeFn& edges, int n_in, int n_out) {
    std::vector<int32_t> out_rows, in_rows, tile_chunk0, ksteps;
    std::vector<uint16_t> planes;
    std::vector<Edge> ev;
    double useful = 0.0, issued = 0.0;
    for (const auto& rows : tiles) {
        std::vector<std::vector<Edge>> row_edges(rows.size());
        std::map<int, int> col;                     // input node -> K index (sorted by node)
        for (size_t m = 0; m < rows.size(); ++m) {
            ev.clear();
            edges(rows[m], ev);
            row_edges[m] = ev;
            for (const Edge& e : ev) col[e.in] = 0;
            useful += (double)ev.size();
        }
        int K = 0;
        for (auto& kv : col) kv.second = K++;
        const int nch = std::max(1, (K + 63) / 64);
        tile_chunk0.push_back((int32_t)ksteps.size());
        const size_t chunk_base = ksteps.size();
        for (int c = 0; c < nch; ++c) {
            const int kc = std::min(64, K - c * 64);
            ksteps.push_back(std::max(1, (kc + 15) / 16));
            issued += 128.0 * 16.0 * ksteps.back();
        }
        in_rows.resize((chunk_base + nch) * 64, -1);
        for (const auto& kv : col) in_rows[chunk_base * 64 + kv.second] = kv.first;
        planes.resize((chunk_base + nch) * (size_t)(2 * TILE * 64), 0);
        const int Kp = nch * 64;
        std::vector<float> dense((size_t)TILE * Kp, 0.f);
        for (size_t m = 0; m < rows.size(); ++m)
            for (const Edge& e : row_edges[m]) dense[m * Kp + col[e.in]] += e.w;
        for (size_t m = 0; m < rows.size(); ++m)
            for (int k = 0; k < K; ++k) {
                const float x = dense[m * Kp + k];
                if (x == 0.f) continue;
                const int c = k / 64, kk = k % 64;
                uint16_t hi, lo;
                split_host(x, hi, lo);
                const size_t el = (size_t)(swz((uint32_t)m, (uint32_t)(kk / 8)) + (kk % 8) * 2) / 2;
                const size_t cb = (chunk_base + c) * (size_t)(2 * TILE * 64);
                planes[cb + el] = hi;
                planes[cb + TILE * 64 + el] = lo;
            }
        for (int m = 0; m < TILE; ++m) out_rows.push_back(m < (int)rows.size() ? rows[m] : -1);
    }
    tile_chunk0.push_back((int32_t)ksteps.size());
    PropPlan* p = new PropPlan();
    p->nchunks = (int)ksteps.size();
    p->density = issued > 0 ? useful / issued : 0.0;
    const size_t b_planes = planes.size() * sizeof(uint16_t), b_out = out_rows.size() * 4, b_in = in_rows.size() * 4,
                 b_tc = tile_chunk0.size() * 4, b_ks = ksteps.size() * 4;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t total = up(b_planes) + up(b_out) + up(b_in) + up(b_tc) + up(b_ks);
    if (cudaMalloc(&p->blob, total) != cudaSuccess) { delete p; return nullptr; }
    unsigned char* d = reinterpret_cast<unsigned char*>(p->blob);
    size_t off = 0;
    auto put = [&](const void* src, size_t bytes) {
        void* dst = d + off;
        cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        off += up(bytes);
        return dst;
    };
    p->dev.a_planes = reinterpret_cast<const uint16_t*>(put(planes.data(), b_planes));
    p->dev.out_rows = reinterpret_cast<const int32_t*>(put(out_rows.data(), b_out));
    p->dev.in_rows = reinterpret_cast<const int32_t*>(put(in_rows.data(), b_in));
    p->dev.tile_chunk0 = reinterpret_cast<const int32_t*>(put(tile_chunk0.data(), b_tc));
    p->dev.ksteps = reinterpret_cast<const int32_t*>(put(ksteps.data(), b_ks));
    p->dev.ntiles = (int)tiles.size();
    p->dev.n_in = n_in;
    p->dev.n_out = n_out;
    if (cudaGetLastError() != cudaSuccess) { cudaFree(p->blob); delete p; return nullptr; }
    return p;
}

}  // namespace

int prop_tc_init() {
    return cudaFuncSetAttribute(k_tc_prop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROP_SMEM);
}

// plan for  nb = A_k(mu)  (forward) or  nb = A_k^T(mu) [/ freq]  (backward) of one verified-network layer;
// `w` is the layer's HOST weight (conv [co,ci,k,k] / linear [out,in])
PropPlan* prop_plan_build(const LayerDev& L, const float* w, bool backward, bool normalise) {
    const int k = L.ksize, s = L.stride, p = L.pad;
    if (L.kind == GNNB_LAYER_CONV) {
        const int hw_in = L.h_in * L.w_in, hw_out = L.h_out * L.w_out;
        if (!backward) {
            EdgeFn f = [=](int node, std::vector<Edge>& ev) {
                const int co = node / hw_out, y = (node % hw_out) / L.w_out, x = node % L.w_out;
                for (int ci = 0; ci < L.c_in; ++ci)
                    for (int ky = 0; ky < k; ++ky) {
                        const int yy = y * s + ky - p;
                        if (yy < 0 || yy >= L.h_in) continue;
                        for (int kx = 0; kx < k; ++kx) {
                            const int xx = x * s + kx - p;
                            if (xx < 0 || xx >= L.w_in) continue;
                            ev.push_back({ci * hw_in + yy * L.w_in + xx, w[((co * L.c_in + ci) * k + ky) * k + kx]});
                        }
                    }
            };
            return build_plan(grid_tiles(L.c_out, L.h_out, L.w_out), f, L.n_in, L.n_out);
        }
        EdgeFn f = [=](int node, std::vector<Edge>& ev) {
            const int ci = node / hw_in, yi = (node % hw_in) / L.w_in, xi = node % L.w_in;
            int taps = 0;
            const size_t first = ev.size();
            for (int ky = 0; ky < k; ++ky) {
                const int ty = yi + p - ky;
                if (ty < 0 || ty % s != 0 || ty / s >= L.h_out) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int tx = xi + p - kx;
                    if (tx < 0 || tx % s != 0 || tx / s >= L.w_out) continue;
                    ++taps;
                    for (int co = 0; co < L.c_out; ++co)
                        ev.push_back({co * hw_out + (ty / s) * L.w_out + tx / s, w[((co * L.c_in + ci) * k + ky) * k + kx]});
                }
            }
            if (normalise && taps > 0)      // conv_transpose2d(.) / freq, freq = tap count (graph_conv.py:306-312)
                for (size_t i = first; i < ev.size(); ++i) ev[i].w = ev[i].w / (float)taps;
        };
        return build_plan(grid_tiles(L.c_in, L.h_in, L.w_in), f, L.n_out, L.n_in);
    }
    if (!backward) {
        EdgeFn f = [=](int node, std::vector<Edge>& ev) {
            for (int i = 0; i < L.n_in; ++i) ev.push_back({i, w[(size_t)node * L.n_in + i]});
        };
        return build_plan(range_tiles(L.n_out), f, L.n_in, L.n_out);
    }
    EdgeFn f = [=](int node, std::vector<Edge>& ev) {
        for (int o = 0; o < L.n_out; ++o) ev.push_back({o, w[(size_t)o * L.n_in + node]});
    };
    return build_plan(range_tiles(L.n_in), f, L.n_out, L.n_in);
}

void prop_plan_free(PropPlan* p) {
    if (!p) return;
    if (p->blob) cudaFree(p->blob);
    delete p;
}

double prop_plan_density(const PropPlan* p) { return p ? p->density : 0.0; }

void prop_tc_run(const PropPlan* plan, const float* mu_in, float* nb_img, int Bc, cudaStream_t st, int64_t* launches) {
    const int64_t nitems = (int64_t)plan->dev.ntiles * ((Bc + 1) / 2);
    const int64_t ctas = (nitems + PWG - 1) / PWG;
    const int grid = (int)(ctas < 1 ? 1 : (ctas < 148 ? ctas : 148));
    k_tc_prop<<<grid, 128 * PWG, PROP_SMEM, st>>>(plan->dev, mu_in, reinterpret_cast<uint16_t*>(nb_img), Bc);
    ++*launches;
}

// nb[b, n, :] = Wp[b, n] * mu[L+1][b, :] (graph_conv.py:324-326) written as the fp16 tile image
namespace {
__global__ void k_property_backward_img(const float* __restrict__ wp, const float* __restrict__ mu_out,
                                        uint16_t* __restrict__ nb_img, int nL, int64_t total8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 7);
        const int64_t row = i >> 3;              // b * nL + n
        const int64_t b = row / nL;
        const float w = wp[row] * ASCALE;
        const float4 m0 = *reinterpret_cast<const float4*>(mu_out + b * P + chunk * 8);
        const float4 m1 = *reinterpret_cast<const float4*>(mu_out + b * P + chunk * 8 + 4);
        uint4 hi, lo;
        split2(w * m0.x, w * m0.y, hi.x, lo.x);
        split2(w * m0.z, w * m0.w, hi.y, lo.y);
        split2(w * m1.x, w * m1.y, hi.z, lo.z);
        split2(w * m1.z, w * m1.w, hi.w, lo.w);
        unsigned char* img = reinterpret_cast<unsigned char*>(nb_img) + (row / TILE) * (int64_t)ABUF;
        const uint32_t off = swz((uint32_t)(row % TILE), (uint32_t)chunk);
        *reinterpret_cast<uint4*>(img + off) = hi;
        *reinterpret_cast<uint4*>(img + APLANE + off) = lo;
    }
}
}  // namespace

void prop_tc_property_backward(const float* wp, const float* mu_out, float* nb_img, int nL, int Bc, cudaStream_t st, int64_t* launches) {
    const int64_t total8 = (int64_t)Bc * nL * 8;
    const int64_t blocks = (total8 + 255) / 256;
    k_property_backward_img<<<(unsigned)(blocks < 1 ? 1 : (blocks < 148 * 16 ? blocks : 148 * 16)), 256, 0, st>>>(
        wp, mu_out, reinterpret_cast<uint16_t*>(nb_img), nL, total8);
    ++*launches;
}

}  // namespace gnnb
