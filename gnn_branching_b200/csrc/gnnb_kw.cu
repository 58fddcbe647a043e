// Batched KW (Wong & Kolter) intermediate bounds on the GPU — the bound producer in front of the scoring path
// (SURVEY §8f rank 3).  NOT VALIDATED ON A GPU YET: written after the GPU budget of round 1 was spent; its parity tests run
// in their own process and are marked xfail until they have passed once.  Nothing on the scoring path calls it.
//
// Reference: DualNetwork.__init__ of the vendored convex_adversarial (dual_network.py:15-101, dual_layers.py:207-312,
// dual_inputs.py:24-70) as called from init_kw_bounds (plnn/dual_network_linear_approximation.py:205-288).  The reference
// pushes x, the n0 x n0 identity, the biases and one scaled unit vector per ambiguous ReLU FORWARD through the layers; the
// number of columns depends on the domain.  Here the same numbers are computed by the transposed recursion, whose column
// count is fixed (one column per output neuron), so B domains x 64-column groups batch into the existing propagation
// kernels (oracle/kw_bounds_oracle.py states the forward form and is pinned to the reference; both forms are the same sums).
//
// For output neuron o of layer k:   s_k = e_o;   for j = k-1 .. 1:  t_j = A_{j+1}^T s_{j+1},  s_j = d_j * t_j;   t_0 = A_1^T s_1
//     centre = t_0 . x + sum_{j<=k} s_j . b_j
//     zl_k[o] = centre - eps |t_0|_1 + sum_{j<k} sum_{i in I_j} zl_j[i] relu(-s_j[i])
//     zu_k[o] = centre + eps |t_0|_1 - sum_{j<k} sum_{i in I_j} zl_j[i] relu(+s_j[i])
// with d_j = [zl_j >= 0] + I_j zu_j / (zu_j - zl_j), I_j = [zl_j < 0 < zu_j] from the (already final) bounds of layer j, then
// intersected with the provided bounds of the parent domain (dual_network.py:85-86).  A_j^T is prop_backward without the
// tap-count normalisation (gnnb_prop.cu), 64 columns riding in the 64 "embedding channels" of a [pairs, n, 64] tensor,
// pair = (domain, column group).
#include <string>
#include <vector>

#include "gnnb_common.cuh"

namespace gnnb {
namespace {

constexpr int KW_COLS = P;          // columns per pair = the channel width of the propagation kernels

// s_k of a column group: buf[pair][node][c] = (node == g * 64 + c)
__global__ void k_kw_onehot(float* __restrict__ buf, int n, int G, int64_t p0, int64_t total4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4 = (int)(i & 15);
    const int64_t row = i >> 4;                      // pair_local * n + node
    const int node = (int)(row % n);
    const int g = (int)((p0 + row / n) % G);
    const int c = node - g * KW_COLS;                // the column whose unit vector has its 1 at this node
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= c4 * 4 && c < c4 * 4 + 4) reinterpret_cast<float*>(&v)[c - c4 * 4] = 1.0f;
    reinterpret_cast<float4*>(buf)[i] = v;
}

// t_L of the property output: column 0 = Wp[b, :], the other columns 0 (one pair per domain)
__global__ void k_kw_wp_col(float* __restrict__ buf, const float* __restrict__ wp, int nL, int64_t b0, int64_t total4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4 = (int)(i & 15);
    const int64_t row = i >> 4;                      // pair_local * nL + node
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 == 0) v.x = wp[(b0 + row / nL) * nL + row % nL];
    reinterpret_cast<float4*>(buf)[i] = v;
}

struct KwAcc { float *cx, *l1, *bias, *low, *up; };      // [pairs][64] each

// t_j -> s_j = d_j * t_j in place, and the three sums over the nodes of layer j.  One block per pair, thread = (part, column).
__global__ void __launch_bounds__(256) k_kw_reduce_layer(float* __restrict__ t, const float* __restrict__ zl, const float* __restrict__ zu,
                                                         const float* __restrict__ bias_node, int n, int G, int64_t p0, KwAcc acc) {
    __shared__ float red[3][4][KW_COLS];
    const int64_t pl = blockIdx.x, b = (p0 + pl) / G;
    const int c = threadIdx.x & 63, part = threadIdx.x >> 6;
    float sb = 0.f, sl = 0.f, su = 0.f;
    for (int i = part; i < n; i += 4) {
        const float l = zl[b * n + i], u = zu[b * n + i];
        const bool I = (u > 0.f) && (l < 0.f);
        float d = (l >= 0.f) ? 1.0f : 0.0f;
        if (I) d += __fdiv_rn(u, u - l);
        float* q = t + ((int64_t)pl * n + i) * KW_COLS + c;
        const float s = *q * d;
        *q = s;
        sb = fmaf(s, bias_node[i], sb);
        if (I) { sl = fmaf(l, fmaxf(-s, 0.f), sl); su = fmaf(l, fmaxf(s, 0.f), su); }
    }
    red[0][part][c] = sb; red[1][part][c] = sl; red[2][part][c] = su;
    __syncthreads();
    if (part == 0) {
        const int64_t o = pl * KW_COLS + c;
        acc.bias[o] += (red[0][0][c] + red[0][1][c]) + (red[0][2][c] + red[0][3][c]);
        acc.low[o] += (red[1][0][c] + red[1][1][c]) + (red[1][2][c] + red[1][3][c]);
        acc.up[o] += (red[2][0][c] + red[2][1][c]) + (red[2][2][c] + red[2][3][c]);
    }
}

// t_0 . x and |t_0|_1
__global__ void __launch_bounds__(256) k_kw_reduce_input(const float* __restrict__ t, const float* __restrict__ x, int n0, int G, int64_t p0,
                                                         KwAcc acc) {
    __shared__ float red[2][4][KW_COLS];
    const int64_t pl = blockIdx.x, b = (p0 + pl) / G;
    const int c = threadIdx.x & 63, part = threadIdx.x >> 6;
    float sx = 0.f, s1 = 0.f;
    for (int i = part; i < n0; i += 4) {
        const float v = t[((int64_t)pl * n0 + i) * KW_COLS + c];
        sx = fmaf(v, x[b * n0 + i], sx);
        s1 += fabsf(v);
    }
    red[0][part][c] = sx; red[1][part][c] = s1;
    __syncthreads();
    if (part == 0) {
        const int64_t o = pl * KW_COLS + c;
        acc.cx[o] = (red[0][0][c] + red[0][1][c]) + (red[0][2][c] + red[0][3][c]);
        acc.l1[o] = (red[1][0][c] + red[1][1][c]) + (red[1][2][c] + red[1][3][c]);
    }
}

// bounds of the group's columns, intersected with the provided ones; own_bias: bias of the layer per node (hidden layers) or
// the property bias per domain (ncols == 1)
__global__ void k_kw_finish(KwAcc acc, int ncols, int G, int64_t p0, int64_t npairs, float eps, const float* __restrict__ bias_node,
                            const float* __restrict__ bp, const float* __restrict__ prov_lb, const float* __restrict__ prov_ub,
                            float* __restrict__ out_lb, float* __restrict__ out_ub) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs * KW_COLS) return;
    const int64_t pl = i / KW_COLS, p = p0 + pl, b = p / G;
    const int o = (int)(p % G) * KW_COLS + (int)(i % KW_COLS);
    if (o >= ncols) return;
    const float own = bias_node ? bias_node[o] : bp[b];
    const float centre = acc.cx[i] + (acc.bias[i] + own);
    float zl = centre - eps * acc.l1[i] + acc.low[i];
    float zu = centre + eps * acc.l1[i] - acc.up[i];
    const int64_t at = b * ncols + o;
    if (prov_lb) zl = fmaxf(zl, prov_lb[at]);
    if (prov_ub) zu = fminf(zu, prov_ub[at]);
    out_lb[at] = zl;
    out_ub[at] = zu;
}

__global__ void k_kw_input_box(const float* __restrict__ x, float eps, int64_t total, float* __restrict__ lb, float* __restrict__ ub) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    lb[i] = x[i] - eps;
    ub[i] = x[i] + eps;
}

unsigned blocks_of(int64_t n, int per) { return (unsigned)((n + per - 1) / per < 1 ? 1 : (n + per - 1) / per); }

}  // namespace

// every pointer is a device pointer.  x [B, n0]; wp [B, n_L]; bp [B]; prov_lb / prov_ub: L arrays [B, n_k] (k = 1..L) or null;
// out_lb / out_ub: L + 2 arrays [B, n_k] (k = 0..L+1).  *ws / *ws_cap (floats): caller-kept scratch, grown when too small.
int kw_bounds(const std::vector<LayerDev>& layers, const std::vector<int>& n, int B, const float* x, float eps, const float* wp, const float* bp,
              const float* const* prov_lb, const float* const* prov_ub, float* const* out_lb, float* const* out_ub, float** ws,
              size_t* ws_cap, cudaStream_t st, int64_t* launches, std::string* err) {
    const int L = (int)layers.size();
    int nmax = 0;
    for (int k = 0; k <= L; ++k) nmax = n[k] > nmax ? n[k] : nmax;
    int64_t max_pairs = 0;
    for (int k = 1; k <= L; ++k) {
        const int64_t pairs = (int64_t)B * ((n[k] + KW_COLS - 1) / KW_COLS);
        max_pairs = pairs > max_pairs ? pairs : max_pairs;
    }
    const int64_t PMAX = max_pairs < 1024 ? (max_pairs < 1 ? 1 : max_pairs) : 1024;      // pairs per pass: two [PMAX, nmax, 64] buffers
    const size_t buf_elems = (size_t)PMAX * nmax * KW_COLS, acc_elems = (size_t)PMAX * KW_COLS;
    const size_t need = 2 * buf_elems + 5 * acc_elems + 64;
    if (*ws_cap < need) {
        if (*ws) { cudaStreamSynchronize(st); cudaFree(*ws); }
        *ws = nullptr; *ws_cap = 0;
        const cudaError_t e = cudaMalloc(ws, need * sizeof(float));
        if (e != cudaSuccess) { *err = std::string("KW bounds workspace: ") + cudaGetErrorString(e); return GNNB_ERR_CUDA; }
        *ws_cap = need;
    }
    float* buf[2] = {*ws, *ws + buf_elems};
    float* accp = *ws + 2 * buf_elems;
    KwAcc acc{accp, accp + acc_elems, accp + 2 * acc_elems, accp + 3 * acc_elems, accp + 4 * acc_elems};

    k_kw_input_box<<<blocks_of((int64_t)B * n[0], 256), 256, 0, st>>>(x, eps, (int64_t)B * n[0], out_lb[0], out_ub[0]);
    ++*launches;
    for (int k = 1; k <= L + 1; ++k) {
        const bool out_layer = k == L + 1;
        const int ncols = out_layer ? 1 : n[k];
        const int G = (ncols + KW_COLS - 1) / KW_COLS;
        const int64_t pairs = (int64_t)B * G;
        for (int64_t p0 = 0; p0 < pairs; p0 += PMAX) {
            const int64_t np = (pairs - p0) < PMAX ? (pairs - p0) : PMAX;
            cudaMemsetAsync(accp, 0, 5 * acc_elems * sizeof(float), st);
            int cur = 0;
            int j = k - 1;                          // layer whose post-activation space buf[cur] is about to hold (t_j)
            if (!out_layer) {
                const int64_t total4 = np * n[k] * 16;
                k_kw_onehot<<<blocks_of(total4, 256), 256, 0, st>>>(buf[cur], n[k], G, p0, total4);        // s_k
                ++*launches;
            } else {
                const int64_t total4 = np * n[L] * 16;
                k_kw_wp_col<<<blocks_of(total4, 256), 256, 0, st>>>(buf[cur], wp, n[L], p0, total4);       // t_L (G = 1: pair = domain)
                ++*launches;
            }
            // walk down: buf[cur] holds s_{j+1} (or, for the output layer's first step, already t_L)
            bool have_t = out_layer;
            for (; j >= 1; --j) {
                if (!have_t) {
                    prop_backward(layers[j], buf[cur], buf[cur ^ 1], (int)np, false, st, launches);          // t_j = A_{j+1}^T s_{j+1}
                    cur ^= 1;
                }
                have_t = false;
                k_kw_reduce_layer<<<(unsigned)np, 256, 0, st>>>(buf[cur], out_lb[j], out_ub[j], layers[j - 1].bias_node, n[j], G, p0, acc);
                ++*launches;
            }
            prop_backward(layers[0], buf[cur], buf[cur ^ 1], (int)np, false, st, launches);                  // t_0 = A_1^T s_1
            cur ^= 1;
            k_kw_reduce_input<<<(unsigned)np, 256, 0, st>>>(buf[cur], x, n[0], G, p0, acc);
            k_kw_finish<<<blocks_of(np * KW_COLS, 256), 256, 0, st>>>(acc, ncols, G, p0, np, eps, out_layer ? nullptr : layers[k - 1].bias_node, bp,
                                                                      (!out_layer && prov_lb) ? prov_lb[k - 1] : nullptr,
                                                                      (!out_layer && prov_ub) ? prov_ub[k - 1] : nullptr, out_lb[k], out_ub[k]);
            *launches += 2;
        }
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = std::string("KW bounds: ") + cudaGetErrorString(e); return GNNB_ERR_CUDA; }
    return GNNB_OK;
}

}  // namespace gnnb
