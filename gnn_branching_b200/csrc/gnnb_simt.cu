// Exact-fp32 per-node MLP kernels on CUDA cores (GNNB_MATH_SIMT_FP32).
//
// These are the validation twins of the tcgen05 kernels in gnnb_tc.cu: same stage decomposition, same
// inputs and outputs, plain fp32 FMA arithmetic.  The tests use them to localise tensor-path errors
// stage by stage; they are not the product path.
//
// Stage decomposition (reference: graphnet/graph_conv.py):
//   relax        rows a9 + a11, the bound/dual/primal feature MLPs.  They do not depend on the round t nor on
//                the embeddings, so they are evaluated once per call and reused by every round:
//                  relax_f = fc1_1(relu(fc1(feat_f))) * amb                               (:153-161)
//                  relax_b = bc2_1(relu(bc2([s1, s1*(-d2), s1*d1]))) * amb,
//                            s1 = bc1_2(relu(bc1_1(relu(bc1(feat_b)))))                   (:273-293)
//   update       e  = W3_2(relu(W3([nb*r0, nb*r1])));  mu = W4_2(relu(W4([relax, e]))) * (r0 != 0)
//                forward: fc3, fc3_2, fc4, fc4_2 (:169-181); backward: bc3, bc3_1, bc4, bc4_1 (:331-347);
//                the last backward sweep also evaluates the score head fscore(relu(fnode(mu))) (:448-449)
//   input_embed  mu0 = inp_f_1(relu(inp_f([l0, x, u0])))                                  (:90-95)
//   input_update mu0 = inp_b2_2(relu(inp_b2([inp_b_1(relu(inp_b([l0,u0]))), nb])))        (:380-385)
//
// Tile: 64 rows (nodes) per CTA iteration, 256 threads, thread (ty, tx) owns rows ty*4..+3 and output
// channels tx*4..+3.  Activations sit in shared memory transposed, actT[k][row], so that both operands
// of the inner product are float4 loads.
#include "gnnb_common.cuh"

namespace gnnb {
namespace {

constexpr int TM = 64;        // rows per tile
constexpr int TMP = 68;       // padded leading dimension of the transposed activation buffers
constexpr int NTH = 256;
constexpr int KMAX = 192;

struct Smem {
    float actA[KMAX * TMP];
    float actB[P * TMP];
    float wbuf[64 * 64];
    float rowv[8][TM];        // per-row scalars (r0, r1, amb, gate, d1, -d2, ...)
};

// acc[i][j] = bias[tx*4+j] + sum_k actT[k][ty*4+i] * Wt[k][tx*4+j]
template <int K>
__device__ __forceinline__ void dense64(const float* __restrict__ Wt, const float* __restrict__ bias,
                                        const float* actT, float* wbuf, float acc[4][4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float4 b4 = *reinterpret_cast<const float4*>(bias + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = b4.x; acc[i][1] = b4.y; acc[i][2] = b4.z; acc[i][3] = b4.w; }
    for (int k0 = 0; k0 < K; k0 += 64) {
        const int kc = (K - k0) < 64 ? (K - k0) : 64;
        __syncthreads();                       // activations written / previous wbuf users done
        for (int idx = threadIdx.x; idx < kc * 16; idx += NTH)
            reinterpret_cast<float4*>(wbuf)[idx] = reinterpret_cast<const float4*>(Wt + (size_t)k0 * 64)[idx];
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(actT + (k0 + k) * TMP + ty * 4);
            const float4 w = *reinterpret_cast<const float4*>(wbuf + k * 64 + tx * 4);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
    }
}

// write relu(acc) (or acc) transposed into an activation buffer at feature offset koff
template <bool RELU>
__device__ __forceinline__ void store_act(float* actT, int koff, const float acc[4][4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 v;
        v.x = RELU ? fmaxf(acc[0][j], 0.f) : acc[0][j];
        v.y = RELU ? fmaxf(acc[1][j], 0.f) : acc[1][j];
        v.z = RELU ? fmaxf(acc[2][j], 0.f) : acc[2][j];
        v.w = RELU ? fmaxf(acc[3][j], 0.f) : acc[3][j];
        if (RELU) {   // F.relu keeps NaN
            if (acc[0][j] != acc[0][j]) v.x = acc[0][j];
            if (acc[1][j] != acc[1][j]) v.y = acc[1][j];
            if (acc[2][j] != acc[2][j]) v.z = acc[2][j];
            if (acc[3][j] != acc[3][j]) v.w = acc[3][j];
        }
        *reinterpret_cast<float4*>(actT + (koff + tx * 4 + j) * TMP + ty * 4) = v;
    }
}

// acc * rowscale -> global [rows][64]
__device__ __forceinline__ void store_global(float* __restrict__ out, int64_t row0, int64_t rows, const float acc[4][4],
                                             const float* rowscale) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (row0 + r < rows) {
            const float s = rowscale ? rowscale[r] : 1.0f;
            float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (rowscale) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
            *reinterpret_cast<float4*>(out + (row0 + r) * P + tx * 4) = v;
        }
    }
}

// global [rows][64] tile -> actT[koff + c][r] * rowscale[r]
__device__ __forceinline__ void load_rows_T(float* actT, int koff, const float* __restrict__ src, int64_t row0, int64_t rows,
                                            const float* rowscale) {
    // lane <-> row keeps the transposed shared-memory stores conflict-free
    const int r = threadIdx.x & 63, cq = threadIdx.x >> 6;     // cq: which quarter of the 16 float4 per row
    const bool ok = row0 + r < rows;
    const float s = rowscale ? rowscale[r] : 1.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c4 = cq * 4 + q;
        float4 v = ok ? *reinterpret_cast<const float4*>(src + (row0 + r) * P + c4 * 4) : make_float4(0, 0, 0, 0);
        if (rowscale) { v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
        actT[(koff + c4 * 4 + 0) * TMP + r] = v.x;
        actT[(koff + c4 * 4 + 1) * TMP + r] = v.y;
        actT[(koff + c4 * 4 + 2) * TMP + r] = v.z;
        actT[(koff + c4 * 4 + 3) * TMP + r] = v.w;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTH) k_simt_relax(GnnParams g, NodeInputs in, float* __restrict__ relax_f,
                                                    float* __restrict__ relax_b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int64_t ntiles = (in.rows + TM - 1) / TM;
    float acc[4][4];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        __syncthreads();
        if (threadIdx.x < TM) {
            const int r = threadIdx.x;
            const int64_t row = row0 + r;
            float l = 0.f, u = 1.f, d1 = 0.f, d2 = 0.f, pp = 0.f, po = 0.f, bs = 0.f;
            if (row < in.rows) {
                l = in.lb[row]; u = in.ub[row];
                d1 = in.dual[row * 3 + 1]; d2 = in.dual[row * 3 + 2];
                pp = in.prim_pre[row]; po = in.prim_post[row];
                bs = in.bias_node[row % in.n];
            }
            const Ratio q = compute_ratio(l, u);
            const float dd = d1 - d2;
            // forward features (graph_conv.py:153-159): [beta, l, u, d1-d2, x_pre, x_post, bias]
            s.actA[0 * TMP + r] = q.beta; s.actA[1 * TMP + r] = l; s.actA[2 * TMP + r] = u; s.actA[3 * TMP + r] = dd;
            s.actA[4 * TMP + r] = pp; s.actA[5 * TMP + r] = po; s.actA[6 * TMP + r] = bs;
            // backward features (graph_conv.py:273-279): [l, u, beta, -d2+d1, x_post, x_pre, bias]
            float* fb = s.actA + 8 * TMP;
            fb[0 * TMP + r] = l; fb[1 * TMP + r] = u; fb[2 * TMP + r] = q.beta; fb[3 * TMP + r] = -d2 + d1;
            fb[4 * TMP + r] = po; fb[5 * TMP + r] = pp; fb[6 * TMP + r] = bs;
            s.rowv[0][r] = q.amb; s.rowv[1][r] = d1; s.rowv[2][r] = -d2;
        }
        // ---- forward relaxation branch ----
        dense64<7>(g.wt[FC1], g.bias[FC1], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[FC1_1], g.bias[FC1_1], s.actB, s.wbuf, acc);
        store_global(relax_f, row0, in.rows, acc, s.rowv[0]);
        // ---- backward relaxation branch ----
        dense64<7>(g.wt[BC1], g.bias[BC1], s.actA + 8 * TMP, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[BC1_1], g.bias[BC1_1], s.actB, s.wbuf, acc);
        __syncthreads();                                  // everyone is done with the feature rows of actA
        store_act<true>(s.actA, 0, acc);
        dense64<P>(g.wt[BC1_2], g.bias[BC1_2], s.actA, s.wbuf, acc);
        __syncthreads();                                  // actA (g2) fully consumed before it is overwritten
        {   // [s1, s1*(-d2), s1*d1]
            const int ty = threadIdx.x >> 4;
            float a1[4][4], a2[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a1[i][j] = acc[i][j] * s.rowv[2][ty * 4 + i];
                    a2[i][j] = acc[i][j] * s.rowv[1][ty * 4 + i];
                }
            store_act<false>(s.actA, 0, acc);
            store_act<false>(s.actA, P, a1);
            store_act<false>(s.actA, 2 * P, a2);
        }
        dense64<3 * P>(g.wt[BC2], g.bias[BC2], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[BC2_1], g.bias[BC2_1], s.actB, s.wbuf, acc);
        store_global(relax_b, row0, in.rows, acc, s.rowv[0]);
    }
}

__global__ void __launch_bounds__(NTH) k_simt_update(GnnParams g, int backward, const float* __restrict__ lb,
                                                     const float* __restrict__ ub, const float* __restrict__ nb,
                                                     const float* __restrict__ relax, float* __restrict__ mu_out,
                                                     float* __restrict__ scores, int n, int64_t score_stride,
                                                     int64_t score_off, int64_t rows, unsigned long long* nan_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int l3 = backward ? BC3 : FC3, l3b = backward ? BC3_1 : FC3_2;
    const int l4 = backward ? BC4 : FC4, l4b = backward ? BC4_1 : FC4_2;
    const int64_t ntiles = (rows + TM - 1) / TM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        __syncthreads();
        if (threadIdx.x < TM) {
            const int r = threadIdx.x;
            float l = 0.f, u = 1.f;
            if (row0 + r < rows) { l = lb[row0 + r]; u = ub[row0 + r]; }
            const Ratio q = compute_ratio(l, u);
            s.rowv[0][r] = q.r0; s.rowv[1][r] = q.r1; s.rowv[2][r] = (q.r0 != 0.0f) ? 1.0f : 0.0f;
        }
        __syncthreads();
        load_rows_T(s.actA, 0, nb, row0, rows, s.rowv[0]);          // nb * r0
        load_rows_T(s.actA, P, nb, row0, rows, s.rowv[1]);          // nb * r1
        dense64<2 * P>(g.wt[l3], g.bias[l3], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[l3b], g.bias[l3b], s.actB, s.wbuf, acc);
        __syncthreads();                                             // actA consumed
        load_rows_T(s.actA, 0, relax, row0, rows, nullptr);         // [relax, e]
        store_act<false>(s.actA, P, acc);
        dense64<2 * P>(g.wt[l4], g.bias[l4], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[l4b], g.bias[l4b], s.actB, s.wbuf, acc);
        // gate (r0 != 0) and NaN watch (graph_conv.py:178, 184; 347)
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float gt = s.rowv[2][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] *= gt;
                bad |= (acc[i][j] != acc[i][j]) && (row0 + ty * 4 + i < rows);
            }
        }
        if (bad) atomicAdd(nan_count, 1ULL);
        store_global(mu_out, row0, rows, acc, nullptr);
        if (scores != nullptr) {   // score head on the final embeddings (graph_conv.py:448-449)
            __syncthreads();
            store_act<false>(s.actA, 0, acc);
            dense64<P>(g.wt[FNODE], g.bias[FNODE], s.actA, s.wbuf, acc);
            const float4 ws = *reinterpret_cast<const float4*>(g.wt[FSCORE] + tx * 4);
            const float wsv[4] = {ws.x, ws.y, ws.z, ws.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float h = fmaxf(acc[i][j], 0.f);
                    if (acc[i][j] != acc[i][j]) h = acc[i][j];
                    part = fmaf(h, wsv[j], part);
                }
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off, 16);
                const int64_t row = row0 + ty * 4 + i;
                if (tx == 0 && row < rows) {
                    const int64_t b = row / n, j = row % n;
                    scores[b * score_stride + score_off + j] = part + g.bias[FSCORE][0];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(NTH) k_simt_input_embed(GnnParams g, const float* __restrict__ lb0,
                                                          const float* __restrict__ x, const float* __restrict__ ub0,
                                                          float* __restrict__ mu0, int64_t rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int64_t ntiles = (rows + TM - 1) / TM;
    float acc[4][4];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        __syncthreads();
        if (threadIdx.x < TM) {
            const int r = threadIdx.x;
            const bool ok = row0 + r < rows;
            s.actA[0 * TMP + r] = ok ? lb0[row0 + r] : 0.f;
            s.actA[1 * TMP + r] = ok ? x[row0 + r] : 0.f;
            s.actA[2 * TMP + r] = ok ? ub0[row0 + r] : 0.f;
        }
        dense64<3>(g.wt[INP_F], g.bias[INP_F], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[INP_F_1], g.bias[INP_F_1], s.actB, s.wbuf, acc);
        store_global(mu0, row0, rows, acc, nullptr);
    }
}

__global__ void __launch_bounds__(NTH) k_simt_input_update(GnnParams g, const float* __restrict__ lb0,
                                                           const float* __restrict__ ub0, const float* __restrict__ nb,
                                                           float* __restrict__ mu0, int64_t rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int64_t ntiles = (rows + TM - 1) / TM;
    float acc[4][4];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        __syncthreads();
        if (threadIdx.x < TM) {
            const int r = threadIdx.x;
            const bool ok = row0 + r < rows;
            s.actA[0 * TMP + r] = ok ? lb0[row0 + r] : 0.f;
            s.actA[1 * TMP + r] = ok ? ub0[row0 + r] : 0.f;
        }
        dense64<2>(g.wt[INP_B], g.bias[INP_B], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[INP_B_1], g.bias[INP_B_1], s.actB, s.wbuf, acc);
        __syncthreads();
        store_act<false>(s.actA, 0, acc);                           // [inp_relax, nb]
        load_rows_T(s.actA, P, nb, row0, rows, nullptr);
        dense64<2 * P>(g.wt[INP_B2], g.bias[INP_B2], s.actA, s.wbuf, acc);
        store_act<true>(s.actB, 0, acc);
        dense64<P>(g.wt[INP_B2_2], g.bias[INP_B2_2], s.actB, s.wbuf, acc);
        store_global(mu0, row0, rows, acc, nullptr);
    }
}

int grid_for(int64_t rows) {
    int64_t tiles = (rows + TM - 1) / TM;
    int64_t cap = 148 * 2 * 8;          // a few waves of resident CTAs; tiles beyond that are looped over
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

}  // namespace

int simt_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_simt_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_simt_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_simt_input_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_simt_input_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem))) != cudaSuccess) return e;
    return 0;
}

void simt_relax(const GnnParams& g, const NodeInputs& in, float* relax_f, float* relax_b, cudaStream_t st, int64_t* launches) {
    k_simt_relax<<<grid_for(in.rows), NTH, sizeof(Smem), st>>>(g, in, relax_f, relax_b);
    ++*launches;
}

void simt_update(const GnnParams& g, bool backward, const float* lb, const float* ub, const float* nb, const float* relax,
                 float* mu_out, float* scores, int n, int64_t score_stride, int64_t score_off, int64_t rows,
                 unsigned long long* nan_count, cudaStream_t st, int64_t* launches) {
    k_simt_update<<<grid_for(rows), NTH, sizeof(Smem), st>>>(g, backward ? 1 : 0, lb, ub, nb, relax, mu_out, scores, n,
                                                             score_stride, score_off, rows, nan_count);
    ++*launches;
}

void simt_input_embed(const GnnParams& g, const float* lb0, const float* x, const float* ub0, float* mu0, int64_t rows,
                      cudaStream_t st, int64_t* launches) {
    k_simt_input_embed<<<grid_for(rows), NTH, sizeof(Smem), st>>>(g, lb0, x, ub0, mu0, rows);
    ++*launches;
}

void simt_input_update(const GnnParams& g, const float* lb0, const float* ub0, const float* nb, float* mu0, int64_t rows,
                       cudaStream_t st, int64_t* launches) {
    k_simt_input_update<<<grid_for(rows), NTH, sizeof(Smem), st>>>(g, lb0, ub0, nb, mu0, rows);
    ++*launches;
}

}  // namespace gnnb
