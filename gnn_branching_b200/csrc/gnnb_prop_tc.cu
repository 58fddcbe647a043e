// Propagation of the node embeddings through the verified network's edge sets on the tensor cores.
//
// Every edge set (conv2d, conv_transpose2d / freq, W @ mu, W^T @ mu; graph_conv.py:110-132, 299-322, 361-376) is a
// sparse matrix product  nb[b] = Wmat * mu[b]  with Wmat [n_out x n_in] built from the verified network's own
// weights.  At gnnb_set_network the matrix is cut into dense blocks ("plan"): a tile is 128 output nodes chosen so
// that they share inputs (all channels of a small spatial patch), its K dimension is the sorted list of input nodes
// any of them touches, padded to chunks of 64, and its values are stored as fp16 hi/lo planes in the UMMA K-major
// SWIZZLE_128B image.  conv_transpose's 1/freq tap-count normalisation is folded into the rows.
//
// The kernel (k_tc_prop below) is a warp-specialised gather-GEMM over (tile, 4 subdomains) items; its input is the mu
// tile image of the producing layer and its output the piece-major nb tile image the node-update kernels TMA in
// (formats: gnnb_umma.cuh).
#include <algorithm>
#include <functional>
#include <map>
#include <vector>

#include "gnnb_prop_body.cuh"

namespace gnnb {

struct PropPlan {
    PropPlanDev dev{};
    void* blob = nullptr;
    int nchunks = 0, nksteps = 0;
    double density = 0.0;         // useful MACs / issued MACs
};

namespace {

__global__ void __launch_bounds__(prop::PROP_THREADS, 1) k_tc_prop(PropPlanDev plan, const uint16_t* __restrict__ mu_img,
                                                                   uint16_t* __restrict__ nb_img, int Bc) {
    extern __shared__ unsigned char smem_raw[];
    prop::prop_body<false>(plan, mu_img, nb_img, Bc, smem_raw, (int)blockIdx.x, (int)gridDim.x, true);
}

// the same kernel with the gather indices fetched one chunk ahead (option "gather_prefetch": bit-identical results, measured
// no faster, off by default)
__global__ void __launch_bounds__(prop::PROP_THREADS, 1) k_tc_prop_pf(PropPlanDev plan, const uint16_t* __restrict__ mu_img,
                                                                      uint16_t* __restrict__ nb_img, int Bc) {
    extern __shared__ unsigned char smem_raw[];
    prop::prop_body<true>(plan, mu_img, nb_img, Bc, smem_raw, (int)blockIdx.x, (int)gridDim.x, true);
}

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

PropPlan* build_plan(const LayerTiling& out, const LayerTiling& in, const EdgeFn& edges) {
    std::vector<int32_t> in_rows, tile_chunk0, ksteps, tile_ks0, ks_rows;
    std::vector<uint16_t> planes, ks_w;
    std::vector<Edge> ev;
    double useful = 0.0, issued = 0.0;
    for (int t = 0; t < out.ntiles; ++t) {
        const int32_t* rows = &out.node_of_slot[(size_t)t * TILE];       // -1 = padding slot
        std::vector<std::vector<Edge>> row_edges(TILE);
        std::map<int, int> col;                     // input SLOT -> K index (sorted by slot: neighbouring K rows are neighbouring rows of the mu image)
        for (int m = 0; m < TILE; ++m) {
            if (rows[m] < 0) continue;
            ev.clear();
            edges(rows[m], ev);
            for (Edge& e : ev) { e.in = in.slot_of_node[e.in]; col[e.in] = 0; }
            row_edges[m] = ev;
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
        for (int m = 0; m < TILE; ++m)
            for (const Edge& e : row_edges[m]) dense[(size_t)m * Kp + col[e.in]] += e.w;
        // per-K-step form (fused kernel): K steps [0, ceil(K / 16)) of this tile, at least one
        tile_ks0.push_back((int32_t)(ks_rows.size() / 16));
        {
            const int nks = std::max(1, (K + 15) / 16);
            const size_t r0 = ks_rows.size(), w0 = ks_w.size();
            ks_rows.resize(r0 + (size_t)nks * 16, -1);
            for (const auto& kv : col) ks_rows[r0 + kv.second] = kv.first;
            ks_w.resize(w0 + (size_t)nks * 4096, 0);                        // 8 KB = 4096 fp16 per K step
            for (int m = 0; m < TILE; ++m)
                for (int k = 0; k < K; ++k) {
                    const float x = dense[(size_t)m * Kp + k];
                    if (x == 0.f) continue;
                    uint16_t hi, lo;
                    split_host(x, hi, lo);
                    const int j = k / 16, pce = (k % 16) / 8, e = k % 8;
                    const size_t el = w0 + (size_t)j * 4096 + (size_t)pce * 1024 + (size_t)m * 8 + e;      // fp16 elements: piece 2 KB = 1024
                    ks_w[el] = hi;
                    ks_w[el + 2048] = lo;                                                                  // lo plane 4 KB further
                }
        }
        for (int m = 0; m < TILE; ++m)
            for (int k = 0; k < K; ++k) {
                const float x = dense[(size_t)m * Kp + k];
                if (x == 0.f) continue;
                const int c = k / 64, kk = k % 64;
                uint16_t hi, lo;
                split_host(x, hi, lo);
                const size_t el = (size_t)(swz((uint32_t)m, (uint32_t)(kk / 8)) + (kk % 8) * 2) / 2;
                const size_t cb = (chunk_base + c) * (size_t)(2 * TILE * 64);
                planes[cb + el] = hi;
                planes[cb + TILE * 64 + el] = lo;
            }
    }
    tile_chunk0.push_back((int32_t)ksteps.size());
    tile_ks0.push_back((int32_t)(ks_rows.size() / 16));
    PropPlan* p = new PropPlan();
    p->nchunks = (int)ksteps.size();
    p->nksteps = (int)(ks_rows.size() / 16);
    p->density = issued > 0 ? useful / issued : 0.0;
    const size_t b_planes = planes.size() * sizeof(uint16_t), b_in = in_rows.size() * 4,
                 b_tc = tile_chunk0.size() * 4, b_ks = ksteps.size() * 4, b_tk = tile_ks0.size() * 4, b_kr = ks_rows.size() * 4,
                 b_kw = ks_w.size() * sizeof(uint16_t);
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t total = up(b_planes) + up(b_in) + up(b_tc) + up(b_ks) + up(b_tk) + up(b_kr) + up(b_kw);
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
    p->dev.in_rows = reinterpret_cast<const int32_t*>(put(in_rows.data(), b_in));
    p->dev.tile_chunk0 = reinterpret_cast<const int32_t*>(put(tile_chunk0.data(), b_tc));
    p->dev.ksteps = reinterpret_cast<const int32_t*>(put(ksteps.data(), b_ks));
    p->dev.ks_w = reinterpret_cast<const uint16_t*>(put(ks_w.data(), b_kw));
    p->dev.tile_ks0 = reinterpret_cast<const int32_t*>(put(tile_ks0.data(), b_tk));
    p->dev.ks_rows = reinterpret_cast<const int32_t*>(put(ks_rows.data(), b_kr));
    p->dev.ntiles = out.ntiles;
    p->dev.nslots_in = (int)in.node_of_slot.size();
    p->dev.nslots_out = (int)out.node_of_slot.size();
    if (cudaGetLastError() != cudaSuccess) { cudaFree(p->blob); delete p; return nullptr; }
    return p;
}

}  // namespace

LayerTiling make_tiling(int C, int H, int W) {
    const std::vector<std::vector<int>> tiles = (H == 1 && W == 1) ? range_tiles(C) : grid_tiles(C, H, W);
    LayerTiling t;
    t.ntiles = (int)tiles.size();
    t.node_of_slot.assign((size_t)t.ntiles * TILE, -1);
    t.slot_of_node.assign((size_t)C * H * W, -1);
    for (int i = 0; i < t.ntiles; ++i)
        for (size_t m = 0; m < tiles[i].size(); ++m) {
            t.node_of_slot[(size_t)i * TILE + m] = tiles[i][m];
            t.slot_of_node[tiles[i][m]] = i * TILE + (int)m;
        }
    return t;
}

// prefetched indices + 2 weight stages / 5 gather stages (the same shared memory): more gathered rows in flight, the weight
// ring is rarely what the MMA warp waits for (option "gather_prefetch" = 2)
__global__ void __launch_bounds__(prop::PROP_THREADS, 1) k_tc_prop_pf25(PropPlanDev plan, const uint16_t* __restrict__ mu_img,
                                                                        uint16_t* __restrict__ nb_img, int Bc) {
    extern __shared__ unsigned char smem_raw[];
    prop::prop_body<true, 2, 5>(plan, mu_img, nb_img, Bc, smem_raw, (int)blockIdx.x, (int)gridDim.x, true);
}

int prop_tc_init() {
    cudaError_t e = cudaFuncSetAttribute(k_tc_prop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop::PROP_SMEM);
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_tc_prop_pf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop::PROP_SMEM)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_tc_prop_pf25, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop::PROP_SMEM);
}

// plan for  nb = A_k(mu)  (forward) or  nb = A_k^T(mu) [/ freq]  (backward) of one verified-network layer;
// `w` is the layer's HOST weight (conv [co,ci,k,k] / linear [out,in])
PropPlan* prop_plan_build(const LayerDev& L, const float* w, bool backward, bool normalise, const LayerTiling& out,
                          const LayerTiling& in) {
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
            return build_plan(out, in, f);
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
        return build_plan(out, in, f);
    }
    if (!backward) {
        EdgeFn f = [=](int node, std::vector<Edge>& ev) {
            for (int i = 0; i < L.n_in; ++i) ev.push_back({i, w[(size_t)node * L.n_in + i]});
        };
        return build_plan(out, in, f);
    }
    EdgeFn f = [=](int node, std::vector<Edge>& ev) {
        for (int o = 0; o < L.n_out; ++o) ev.push_back({o, w[(size_t)o * L.n_in + node]});
    };
    return build_plan(out, in, f);
}

void prop_plan_free(PropPlan* p) {
    if (!p) return;
    if (p->blob) cudaFree(p->blob);
    delete p;
}

double prop_plan_density(const PropPlan* p) { return p ? p->density : 0.0; }
const PropPlanDev& prop_plan_dev(const PropPlan* p) { return p->dev; }
double prop_plan_ksteps_per_tile(const PropPlan* p) { return p->dev.ntiles > 0 ? (double)p->nksteps / p->dev.ntiles : 1.0; }
double prop_plan_chunks_per_tile(const PropPlan* p) { return p->dev.ntiles > 0 ? (double)p->nchunks / p->dev.ntiles : 1.0; }

void prop_tc_run(const PropPlan* plan, const float* mu_img, float* nb_img, int Bc, cudaStream_t st, int64_t* launches, int gather_prefetch) {
    const int64_t nitems = (int64_t)plan->dev.ntiles * ((Bc + prop::PD - 1) / prop::PD);
    const int grid = (int)(nitems < 1 ? 1 : (nitems < 148 ? nitems : 148));
    launch_pdl(gather_prefetch == 2 ? k_tc_prop_pf25 : gather_prefetch == 1 ? k_tc_prop_pf : k_tc_prop, grid, prop::PROP_THREADS, prop::PROP_SMEM, st, plan->dev,
               reinterpret_cast<const uint16_t*>(mu_img), reinterpret_cast<uint16_t*>(nb_img), Bc);
    ++*launches;
}

// nb[b, n, :] = Wp[b, n] * mu[L+1][b, :] (graph_conv.py:324-326) written as the fp16 tile image
namespace {
__global__ void k_property_backward_img(const float* __restrict__ wp, const float* __restrict__ mu_out,
                                        uint16_t* __restrict__ nb_img, int nL, int nslots, int64_t total8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 7);
        const int64_t row = i >> 3;              // b * nslots + slot; the last hidden layer is flat: slot = node
        const int64_t b = row / nslots;
        const int node = (int)(row - b * nslots);
        const float w = node < nL ? wp[b * nL + node] * ASCALE : 0.f;
        const float4 m0 = *reinterpret_cast<const float4*>(mu_out + b * P + chunk * 8);
        const float4 m1 = *reinterpret_cast<const float4*>(mu_out + b * P + chunk * 8 + 4);
        uint4 hi, lo;
        split2(w * m0.x, w * m0.y, hi.x, lo.x);
        split2(w * m0.z, w * m0.w, hi.y, lo.y);
        split2(w * m1.x, w * m1.y, hi.z, lo.z);
        split2(w * m1.z, w * m1.w, hi.w, lo.w);
        unsigned char* img = reinterpret_cast<unsigned char*>(nb_img) + (row / TILE) * (int64_t)ABUF;
        const uint32_t off = (uint32_t)chunk * NB_PIECE + (uint32_t)(row % TILE) * 16u;     // piece-major nb image
        *reinterpret_cast<uint4*>(img + off) = hi;
        *reinterpret_cast<uint4*>(img + APLANE + off) = lo;
    }
}
}  // namespace

void prop_tc_property_backward(const float* wp, const float* mu_out, float* nb_img, int nL, int nslots, int Bc, cudaStream_t st, int64_t* launches) {
    const int64_t total8 = (int64_t)Bc * nslots * 8;
    const int64_t blocks = (total8 + 255) / 256;
    k_property_backward_img<<<(unsigned)(blocks < 1 ? 1 : (blocks < 148 * 16 ? blocks : 148 * 16)), 256, 0, st>>>(
        wp, mu_out, reinterpret_cast<uint16_t*>(nb_img), nL, nslots, total8);
    ++*launches;
}

}  // namespace gnnb
