// BaBSR / KW branching heuristic, batched over subdomains: choose_node_conv of plnn/kw_score_conv.py:41-156.
//
// Per subdomain the reference walks the verified network backwards with one scalar `ratio` per node:
//   ratio_L = Wp (the property layer's weights, :81-85);  at every ReLU layer k = L..1 (:88-112)
//     r0, icp = compute_ratio(l_k, u_k);  intercept_k = min(ratio, 0) * icp * mask;
//     score_k = | max(b * ratio * (r0 - 1), b * (ratio * r0)) + min(ratio, 0) * icp | * mask;  ratio <- ratio * r0
//   and through A_k^T to the previous layer (conv_transpose2d without bias :115-119, W^T @ ratio :81-85);
// then picks (:123-152) the node with the largest score unless it sits in `sparsest_layer` or is below the threshold,
// else the most negative intercept of the last layer that has one (at most twice in a row), else the first candidate
// of the most preferred non-empty layer of `random_order`.
// One thread block per subdomain; ratios live in shared memory (two buffers of max_k n_k floats); the network's weights
// are read from global memory (a few hundred KB, L2-resident across the blocks).  This is scalar fp32 work of a few
// hundred kFLOP per subdomain — CUDA cores, no tensor pipe.
#include <math.h>

#include "gnnb_common.cuh"

namespace gnnb {
namespace {

constexpr int BT = 256;

struct Best { float v; int i; };

// first index of the maximum (torch.max semantics on ties: lowest index), NaN wins like in torch
__device__ __forceinline__ bool better_max(float a, int ia, float b, int ib) {
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    const bool na = a != a, nb = b != b;
    if (na || nb) return na && (!nb || ia < ib);
    return a > b || (a == b && ia < ib);
}
__device__ __forceinline__ bool better_min(float a, int ia, float b, int ib) {
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    const bool na = a != a, nb = b != b;
    if (na || nb) return na && (!nb || ia < ib);
    return a < b || (a == b && ia < ib);
}

template <bool MAX>
__device__ Best block_reduce(float v, int i, Best* sh) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (MAX ? better_max(ov, oi, v, i) : better_min(ov, oi, v, i)) { v = ov; i = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sh[warp].v = v; sh[warp].i = i; }
    __syncthreads();
    Best r{sh[0].v, sh[0].i};
    for (int w = 1; w < BT / 32; ++w)
        if (MAX ? better_max(sh[w].v, sh[w].i, r.v, r.i) : better_min(sh[w].v, sh[w].i, r.v, r.i)) r = sh[w];
    return r;
}

struct BabsrArgs {
    const LayerDev* layers;      // device array [L]
    int L, n_hidden, nmax;
    const float* const* lb;      // device array of L + 2 device pointers, as gnnb_frontier.lb
    const float* const* ub;
    const float* wp;             // [B, n_L]
    const float* mask;           // [B, n_hidden]
    const int32_t* hidden_off;   // device [L + 1]
    const int32_t* random_order; // device [L]
    const int32_t* counter_in;   // [B]
    int sparsest_layer;
    float threshold;
    int32_t* decision;           // [B, 2]
    int32_t* counter_out;        // [B]
    int32_t* kind;               // [B] or null: 0 score, 1 intercept, 2 preference order, -1 no candidate
    float* scores;               // [B, n_hidden] or null
};

constexpr int MAXL = 32;

__global__ void __launch_bounds__(BT) k_babsr(BabsrArgs a) {
    extern __shared__ float sm[];
    float* ratio = sm;
    float* next = sm + a.nmax;
    __shared__ Best red[BT / 32];
    __shared__ Best smax[MAXL], smin[MAXL];
    __shared__ int sfirst[MAXL];
    __shared__ int fred[BT / 32];
    const int b = blockIdx.x, t = threadIdx.x;
    const int nL = a.layers[a.L - 1].n_out;
    for (int i = t; i < nL; i += BT) ratio[i] = a.wp[(int64_t)b * nL + i];
    __syncthreads();
    for (int k = a.L; k >= 1; --k) {
        const LayerDev Lk = a.layers[k - 1];
        const int n = Lk.n_out, off = a.hidden_off[k - 1];
        const float* lb = a.lb[k] + (int64_t)b * n;
        const float* ub = a.ub[k] + (int64_t)b * n;
        const float* mk = a.mask + (int64_t)b * a.n_hidden + off;
        float bv = -INFINITY, mv = INFINITY;
        int bi = -1, mi = -1, fi = 0x7fffffff;
        for (int i = t; i < n; i += BT) {
            const float l = lb[i], u = ub[i], m = mk[i], r = ratio[i], bias = Lk.bias_node[i];
            // compute_ratio, kw_score_conv.py:23-27, same operation order (IEEE division)
            float lt = l - fmaxf(l, 0.0f), ut = fmaxf(u, 0.0f);
            if (u != u) ut = u;
            if (l != l) lt = l;
            const float r0 = __fdiv_rn(ut, __fsub_rn(ut, lt));
            const float icp = __fmul_rn(__fmul_rn(-1.0f, lt), r0);
            const float icand = __fmul_rn(fminf(r, 0.0f), icp);                        // :92-93
            const float c1 = __fmul_rn(bias, __fmul_rn(r, __fsub_rn(r0, 1.0f)));       // :100-101
            const float rn = __fmul_rn(r, r0);                                         // :102
            const float c2 = __fmul_rn(bias, rn);                                      // :103
            const float sc = __fmul_rn(fabsf(__fadd_rn(fmaxf(c1, c2), icand)), m);     // :104-110
            const float ic = __fmul_rn(icand, m);                                      // :94
            ratio[i] = rn;
            if (a.scores) a.scores[(int64_t)b * a.n_hidden + off + i] = sc;
            if (better_max(sc, i, bv, bi)) { bv = sc; bi = i; }
            if (better_min(ic, i, mv, mi)) { mv = ic; mi = i; }
            if (m != 0.0f && i < fi) fi = i;
        }
        const Best mx = block_reduce<true>(bv, bi, red);
        const Best mn = block_reduce<false>(mv, mi, red);
        // first candidate of the layer
        int f = fi;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
        if ((t & 31) == 0) fred[t >> 5] = f;
        __syncthreads();
        if (t == 0) {
            int ff = fred[0];
            for (int w = 1; w < BT / 32; ++w) ff = min(ff, fred[w]);
            smax[k - 1] = mx; smin[k - 1] = mn; sfirst[k - 1] = ff == 0x7fffffff ? -1 : ff;
        }
        __syncthreads();
        if (k == 1) break;
        // ratio of the previous layer: A_k^T applied to the scalar map (no bias, no tap-count normalisation)
        const int n_in = Lk.n_in;
        if (Lk.kind == GNNB_LAYER_CONV) {
            const int hw_in = Lk.h_in * Lk.w_in, hw_out = Lk.h_out * Lk.w_out, ks = Lk.ksize, s = Lk.stride, p = Lk.pad;
            const int wstride = Lk.c_in * ks * ks;                       // between output channels of the weight tensor
            for (int j = t; j < n_in; j += BT) {
                const int ci = j / hw_in, yi = (j % hw_in) / Lk.w_in, xi = j % Lk.w_in;
                float acc = 0.f;
                for (int ky = 0; ky < ks; ++ky) {                          // taps first: their validity does not depend on co
                    const int ty = yi + p - ky;
                    if (ty < 0 || ty % s != 0 || ty / s >= Lk.h_out) continue;
                    for (int kx = 0; kx < ks; ++kx) {
                        const int tx = xi + p - kx;
                        if (tx < 0 || tx % s != 0 || tx / s >= Lk.w_out) continue;
                        const float* w = Lk.weight + (ci * ks + ky) * ks + kx;
                        const float* r = ratio + (ty / s) * Lk.w_out + tx / s;
#pragma unroll 4
                        for (int co = 0; co < Lk.c_out; ++co) acc = fmaf(__ldg(w + co * wstride), r[co * hw_out], acc);
                    }
                }
                next[j] = acc;
            }
        } else {
            // W^T @ ratio: each thread owns up to 8 input nodes (BT apart, so a warp reads 128 contiguous bytes of a weight
            // row) and walks the rows with 8 independent loads in flight per row, 4 rows unrolled
            for (int j0 = 0; j0 < n_in; j0 += BT * 8) {
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll 4
                for (int o = 0; o < n; ++o) {
                    const float r = ratio[o];
                    const float* wrow = Lk.weight + (int64_t)o * n_in + j0 + t;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (j0 + t + q * BT < n_in) acc[q] = fmaf(__ldg(wrow + q * BT), r, acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (j0 + t + q * BT < n_in) next[j0 + t + q * BT] = acc[q];
            }
        }
        __syncthreads();
        float* tmp = ratio; ratio = next; next = tmp;
    }
    if (t == 0) {                                              // kw_score_conv.py:123-152
        int counter = a.counter_in ? a.counter_in[b] : 0, dl = -1, di = -1, kind = -1;
        bool any = false;
        for (int k = 0; k < a.L; ++k) any |= sfirst[k] >= 0;
        if (any) {
            // max(max_info) compares (value, index) tuples (kw_score_conv.py:123-124): on equal maxima the layer whose in-layer
            // argmax index is larger wins, and max_info.index() returns the first layer among fully equal pairs
            int best = 0;
            for (int k = 1; k < a.L; ++k)
                if (smax[k].v > smax[best].v || (smax[k].v == smax[best].v && smax[k].i > smax[best].i) ||
                    (smax[k].v != smax[k].v && smax[best].v == smax[best].v)) best = k;
            if (best != a.sparsest_layer && smax[best].v > a.threshold) {
                dl = best; di = smax[best].i; kind = 0;
            } else {
                int il = -1;
                for (int k = 0; k < a.L; ++k)
                    if (smin[k].v < -1e-4f) il = k;
                if (il >= 0 && counter < 2) {
                    dl = il; di = smin[il].i; kind = 1;
                    counter += 1;
                    if (il != 0) counter = 0;
                } else {
                    for (int q = a.L - 1; q >= 0 && dl < 0; --q) {
                        const int pl = a.random_order[q];
                        if (pl >= 0 && pl < a.L && sfirst[pl] >= 0) { dl = pl; di = sfirst[pl]; }
                    }
                    kind = 2;
                    counter = 0;
                }
            }
        }
        a.decision[2 * b] = dl; a.decision[2 * b + 1] = di;
        a.counter_out[b] = counter;
        if (a.kind) a.kind[b] = kind;
    }
}

}  // namespace

int babsr_max_layers() { return MAXL; }

int babsr_run(const LayerDev* d_layers, int L, int n_hidden, int nmax, const float* const* d_lb, const float* const* d_ub,
              const float* wp, const float* mask, const int32_t* d_hidden_off, const int32_t* d_random_order,
              const int32_t* counter_in, int sparsest_layer, float threshold, int32_t* decision, int32_t* counter_out,
              int32_t* kind, float* scores, int B, cudaStream_t st, int64_t* launches) {
    BabsrArgs a{d_layers, L, n_hidden, nmax, d_lb, d_ub, wp, mask, d_hidden_off, d_random_order, counter_in, sparsest_layer,
                threshold, decision, counter_out, kind, scores};
    const size_t smem = (size_t)2 * nmax * sizeof(float);
    if (smem > 200 * 1024) return -1;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k_babsr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    k_babsr<<<B, BT, smem, st>>>(a);
    ++*launches;
    return 0;
}

}  // namespace gnnb
