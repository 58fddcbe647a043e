// Device-resident domain queue of the branch-and-bound loop.
//
// The reference keeps `domains`, a Python list of ReLUDomain objects (mask, lower / upper bound, every layer's bounds, the
// GNN decision; plnn/relu_conv_gnnkwthreshold.py:20-53) sorted by lower bound, and works on it with
//   add_domain(candidate, domains)     bisect.insort_left                                   plnn/branch_and_bound.py:159-164
//   pick_out(domains, threshold)       pop the front until one has lower_bound < threshold                       :167-184
//   prune_domains(domains, threshold)  keep the prefix with lower_bound < threshold                               :264-281
//   domains[0].lower_bound             the global lower bound                     relu_conv_gnnkwthreshold.py:240-244
// Every domain is ~100 KB of bounds, so a GPU loop that scores whole frontiers must not round-trip them through host lists.
//
// Here the payload of a domain lives in a slot of a device pool (struct of arrays: bounds [cap, NB], mask [cap, n_hidden]
// int8, upper bound, decision) and never moves; the queue itself is a sorted array of (key, slot) pairs on the device:
//   key  = order-preserving bits of the lower bound << 32 | ~insertion number    (ties: the newest first, as insort_left)
//   add   scatter the kept children into free slots, append their pairs, one radix sort of the live pairs (CUB)
//   pick  the first n pairs whose key is below the threshold key are the next n pick_out results, in order: gather their
//         payload rows into dense [n, .] arrays (the Frontier layout gnnb_score reads), return the slots to the free stack
//   prune binary search of the threshold key, the tail's slots go back to the free stack
// The host keeps four integers (size, head, free count, insertion counter); per operation it reads back one count.
#include <cub/device/device_radix_sort.cuh>

#include <string>
#include <vector>

#include "gnnb_common.cuh"

namespace gnnb {

struct DomainQueue {
    int device = 0;
    int64_t cap = 0;
    int L = 0, n_hidden = 0;
    std::vector<int> n;                 // n[0..L+1]
    std::vector<int64_t> off;           // offset of layer k inside a bounds row
    int64_t NB = 0;                     // floats per bounds row = sum n[k]
    // payload pool
    float *p_lb = nullptr, *p_ub = nullptr;      // [cap, NB]
    int8_t* p_mask = nullptr;                    // [cap, n_hidden]  -1 undecided, 0 / 1 fixed
    float *p_lower = nullptr, *p_upper = nullptr;   // [cap]
    int32_t* p_dec = nullptr;                    // [cap, 2]
    // sorted (key, slot) pairs: two buffers, the live ones are buf[cur][head .. head + size)
    uint64_t* keys[2] = {nullptr, nullptr};
    int32_t* slots[2] = {nullptr, nullptr};
    int cur = 0;
    int64_t size = 0, head = 0;
    int32_t* free_stack = nullptr;               // [cap] free payload slots, top = n_free - 1
    int64_t n_free = 0;
    uint32_t seq = 0;                            // insertion counter
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int32_t *d_sel = nullptr, *d_count = nullptr;   // [max batch] kept children of an add; one counter
    int64_t sel_cap = 0;
    float* stage = nullptr;                      // staging of host-buffer calls
    size_t stage_bytes = 0;
    std::string err;
};

namespace {

__host__ __device__ inline uint32_t f2u(float f) {      // order-preserving float -> uint
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    if ((u << 1) == 0) u = 0;                            // -0.0 == +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline uint64_t make_key(float lb, uint32_t seq) { return ((uint64_t)f2u(lb) << 32) | (uint64_t)(0xFFFFFFFFu - seq); }

// one block: indices of the children with keep != 0 (all when keep is null), in order; *count = how many
__global__ void __launch_bounds__(1024) k_q_select(const uint8_t* __restrict__ keep, int B, int32_t* __restrict__ sel, int32_t* __restrict__ count) {
    __shared__ int32_t part[1024];
    const int per = (B + 1023) / 1024, lo = threadIdx.x * per, hi = min(B, lo + per);
    int32_t c = 0;
    for (int i = lo; i < hi; ++i) c += (keep == nullptr || keep[i]) ? 1 : 0;
    part[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int32_t run = part[threadIdx.x] - c;
    for (int i = lo; i < hi; ++i)
        if (keep == nullptr || keep[i]) sel[run++] = i;
    if (threadIdx.x == 1023) *count = part[1023];
}

// pairs of the kept children: slot from the top of the free stack, key from the lower bound and the insertion number
__global__ void k_q_append(const int32_t* __restrict__ sel, int count, const float* __restrict__ lower, const int32_t* __restrict__ free_stack,
                           int64_t n_free, uint32_t seq0, uint64_t* __restrict__ keys, int32_t* __restrict__ slots) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    slots[j] = free_stack[n_free - 1 - j];
    keys[j] = make_key(lower[sel[j]], seq0 + (uint32_t)j);
}

struct RowCopy {       // one bounds layer: dense [B, n] <-> pool rows [cap, NB] at column `off`
    float* pool;
    const float* dense_in;
    float* dense_out;
    int n;
    int64_t off;
};
constexpr int MAX_COPY = 2 * 18;

struct ScatterArgs {
    RowCopy c[MAX_COPY];
    int ncopy;
    int64_t NB;
    const int32_t* sel;        // add: child index of entry j (null on pick)
    const int32_t* slots;      // pool slot of entry j
    int count;
    // scalars and small rows
    const float *lower_in, *upper_in;
    const int8_t* mask_in;
    const int32_t* dec_in;
    float *lower_out, *upper_out;
    int8_t* mask_out;
    int32_t* dec_out;
    float *p_lower, *p_upper;
    int8_t* p_mask;
    int32_t* p_dec;
    int n_hidden;
};

// grid (count, ncopy + 1): block (j, c) moves layer c of entry j; the extra block row moves the mask, bounds scalars, decision
template <bool TO_POOL>
__global__ void __launch_bounds__(256) k_q_rows(ScatterArgs a) {
    const int j = blockIdx.x;
    const int64_t slot = a.slots[j];
    const int64_t src = TO_POOL ? (int64_t)a.sel[j] : (int64_t)j;       // dense row
    if ((int)blockIdx.y < a.ncopy) {
        const RowCopy c = a.c[blockIdx.y];
        float* pool = c.pool + slot * a.NB + c.off;
        if (TO_POOL) { const float* d = c.dense_in + src * c.n; for (int i = threadIdx.x; i < c.n; i += blockDim.x) pool[i] = d[i]; }
        else { float* d = c.dense_out + src * c.n; for (int i = threadIdx.x; i < c.n; i += blockDim.x) d[i] = pool[i]; }
        return;
    }
    if (TO_POOL) {
        for (int i = threadIdx.x; i < a.n_hidden; i += blockDim.x) a.p_mask[slot * a.n_hidden + i] = a.mask_in[src * a.n_hidden + i];
        if (threadIdx.x == 0) {
            a.p_lower[slot] = a.lower_in[src]; a.p_upper[slot] = a.upper_in[src];
            a.p_dec[slot * 2] = a.dec_in ? a.dec_in[src * 2] : -1; a.p_dec[slot * 2 + 1] = a.dec_in ? a.dec_in[src * 2 + 1] : -1;
        }
    } else {
        if (a.mask_out) for (int i = threadIdx.x; i < a.n_hidden; i += blockDim.x) a.mask_out[src * a.n_hidden + i] = a.p_mask[slot * a.n_hidden + i];
        if (threadIdx.x == 0) {
            if (a.lower_out) a.lower_out[src] = a.p_lower[slot];
            if (a.upper_out) a.upper_out[src] = a.p_upper[slot];
            if (a.dec_out) { a.dec_out[src * 2] = a.p_dec[slot * 2]; a.dec_out[src * 2 + 1] = a.p_dec[slot * 2 + 1]; }
        }
    }
}

// number of live pairs whose key is below the threshold key (the pairs are sorted): binary search by one thread
__global__ void k_q_lower_bound(const uint64_t* __restrict__ keys, int64_t size, uint64_t thr, int32_t* __restrict__ out) {
    int64_t lo = 0, hi = size;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < thr) lo = mid + 1; else hi = mid;
    }
    *out = (int32_t)lo;
}

__global__ void k_q_release(const int32_t* __restrict__ slots, int64_t count, int32_t* __restrict__ free_stack, int64_t n_free) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) free_stack[n_free + j] = slots[j];
}

__global__ void k_q_iota_desc(int32_t* __restrict__ free_stack, int64_t cap) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cap) free_stack[j] = (int32_t)(cap - 1 - j);       // slot 0 on top
}

#define QCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) { q->err = std::string(#call) + ": " + cudaGetErrorString(e__); return GNNB_ERR_CUDA; } \
    } while (0)

int ensure_sel(DomainQueue* q, int64_t B) {
    if (q->sel_cap >= B) return GNNB_OK;
    if (q->d_sel) cudaFree(q->d_sel);
    q->d_sel = nullptr;
    QCU(cudaMalloc(&q->d_sel, (size_t)B * sizeof(int32_t)));
    q->sel_cap = B;
    return GNNB_OK;
}

int ensure_stage(DomainQueue* q, size_t bytes) {
    if (q->stage_bytes >= bytes) return GNNB_OK;
    if (q->stage) cudaFree(q->stage);
    q->stage = nullptr; q->stage_bytes = 0;
    QCU(cudaMalloc(&q->stage, bytes));
    q->stage_bytes = bytes;
    return GNNB_OK;
}

}  // namespace

int queue_create(int device, const std::vector<int>& n, int n_hidden, int64_t capacity, DomainQueue** out, std::string* err) {
    DomainQueue* q = new DomainQueue();
    q->device = device; q->cap = capacity; q->n = n; q->L = (int)n.size() - 2; q->n_hidden = n_hidden;
    q->off.resize(n.size());
    for (size_t k = 0; k < n.size(); ++k) { q->off[k] = q->NB; q->NB += n[k]; }
    auto fail = [&](const char* what, cudaError_t e) { *err = std::string(what) + ": " + cudaGetErrorString(e); queue_destroy(q); return GNNB_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if (2 * (int)n.size() > MAX_COPY) { *err = "too many layers for the domain queue"; delete q; return GNNB_ERR_UNSUPPORTED; }
    const size_t c = (size_t)capacity;
    if ((e = cudaMalloc(&q->p_lb, c * q->NB * sizeof(float))) != cudaSuccess) return fail("queue pool (lower bounds)", e);
    if ((e = cudaMalloc(&q->p_ub, c * q->NB * sizeof(float))) != cudaSuccess) return fail("queue pool (upper bounds)", e);
    if ((e = cudaMalloc(&q->p_mask, c * n_hidden)) != cudaSuccess) return fail("queue pool (mask)", e);
    if ((e = cudaMalloc(&q->p_lower, c * sizeof(float))) != cudaSuccess) return fail("queue pool", e);
    if ((e = cudaMalloc(&q->p_upper, c * sizeof(float))) != cudaSuccess) return fail("queue pool", e);
    if ((e = cudaMalloc(&q->p_dec, c * 2 * sizeof(int32_t))) != cudaSuccess) return fail("queue pool", e);
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaMalloc(&q->keys[i], c * sizeof(uint64_t))) != cudaSuccess) return fail("queue keys", e);
        if ((e = cudaMalloc(&q->slots[i], c * sizeof(int32_t))) != cudaSuccess) return fail("queue slots", e);
    }
    if ((e = cudaMalloc(&q->free_stack, c * sizeof(int32_t))) != cudaSuccess) return fail("queue free stack", e);
    if ((e = cudaMalloc(&q->d_count, sizeof(int32_t))) != cudaSuccess) return fail("queue counter", e);
    cub::DeviceRadixSort::SortPairs(nullptr, q->cub_bytes, q->keys[0], q->keys[1], q->slots[0], q->slots[1], (int64_t)capacity);
    if ((e = cudaMalloc(&q->cub_tmp, q->cub_bytes ? q->cub_bytes : 16)) != cudaSuccess) return fail("queue sort scratch", e);
    k_q_iota_desc<<<(unsigned)((capacity + 255) / 256), 256>>>(q->free_stack, capacity);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("queue init", e);
    q->n_free = capacity;
    *out = q;
    return GNNB_OK;
}

void queue_destroy(DomainQueue* q) {
    if (!q) return;
    cudaSetDevice(q->device);
    void* ptrs[] = {q->p_lb, q->p_ub, q->p_mask, q->p_lower, q->p_upper, q->p_dec, q->keys[0], q->keys[1], q->slots[0], q->slots[1],
                    q->free_stack, q->d_count, q->cub_tmp, q->d_sel, q->stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete q;
}

const std::string& queue_error(const DomainQueue* q) { return q->err; }
int64_t queue_size(const DomainQueue* q) { return q->size; }

int queue_global_lb(DomainQueue* q, float* out, cudaStream_t st) {
    if (q->size == 0) return GNNB_ERR_STATE;
    int32_t slot = 0;
    QCU(cudaMemcpyAsync(&slot, q->slots[q->cur] + q->head, sizeof slot, cudaMemcpyDeviceToHost, st));
    QCU(cudaStreamSynchronize(st));
    QCU(cudaMemcpyAsync(out, q->p_lower + slot, sizeof(float), cudaMemcpyDeviceToHost, st));
    QCU(cudaStreamSynchronize(st));
    return GNNB_OK;
}

// every pointer is a device pointer; returns the number added through *added
int queue_add(DomainQueue* q, int B, const float* lower, const float* upper, const float* const* lb, const float* const* ub,
              const int8_t* mask, const int32_t* decision, const uint8_t* keep, int32_t* added, cudaStream_t st, int64_t* launches) {
    if (B <= 0) { *added = 0; return GNNB_OK; }
    int rc = ensure_sel(q, B);
    if (rc != GNNB_OK) return rc;
    k_q_select<<<1, 1024, 0, st>>>(keep, B, q->d_sel, q->d_count);
    int32_t count = 0;
    QCU(cudaMemcpyAsync(&count, q->d_count, sizeof count, cudaMemcpyDeviceToHost, st));
    QCU(cudaStreamSynchronize(st));
    ++*launches;
    *added = count;
    if (count == 0) return GNNB_OK;
    if (count > q->n_free) { q->err = "domain queue is full"; return GNNB_ERR_STATE; }
    // move the live pairs to the front of the buffer if the appended ones would not fit behind them
    if (q->head + q->size + count > q->cap) {
        QCU(cudaMemcpyAsync(q->keys[q->cur ^ 1], q->keys[q->cur] + q->head, (size_t)q->size * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        QCU(cudaMemcpyAsync(q->slots[q->cur ^ 1], q->slots[q->cur] + q->head, (size_t)q->size * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        q->cur ^= 1; q->head = 0;
    }
    uint64_t* kin = q->keys[q->cur] + q->head;
    int32_t* sin = q->slots[q->cur] + q->head;
    k_q_append<<<(count + 255) / 256, 256, 0, st>>>(q->d_sel, count, lower, q->free_stack, q->n_free, q->seq, kin + q->size, sin + q->size);
    ScatterArgs a{};
    a.ncopy = 0;
    for (int k = 0; k <= q->L + 1; ++k) {
        a.c[a.ncopy++] = RowCopy{q->p_lb, lb[k], nullptr, q->n[k], q->off[k]};
        a.c[a.ncopy++] = RowCopy{q->p_ub, ub[k], nullptr, q->n[k], q->off[k]};
    }
    a.NB = q->NB; a.sel = q->d_sel; a.slots = sin + q->size; a.count = count;
    a.lower_in = lower; a.upper_in = upper; a.mask_in = mask; a.dec_in = decision;
    a.p_lower = q->p_lower; a.p_upper = q->p_upper; a.p_mask = q->p_mask; a.p_dec = q->p_dec; a.n_hidden = q->n_hidden;
    k_q_rows<true><<<dim3((unsigned)count, (unsigned)a.ncopy + 1), 256, 0, st>>>(a);
    // one sort of the live pairs (the old ones are already in order; a radix sort over 64-bit keys is a few passes over 12 B / pair)
    const int64_t total = q->size + count;
    size_t bytes = q->cub_bytes;
    QCU(cub::DeviceRadixSort::SortPairs(q->cub_tmp, bytes, kin, q->keys[q->cur ^ 1], sin, q->slots[q->cur ^ 1], total, 0, 64, st));
    *launches += 3;
    q->cur ^= 1; q->head = 0; q->size = total; q->n_free -= count; q->seq += (uint32_t)count;
    QCU(cudaGetLastError());
    return GNNB_OK;
}

// pick_out repeated up to max_B times; outputs are dense device arrays [max_B, .]; *picked = how many
int queue_pick(DomainQueue* q, int max_B, float threshold, bool discard, int32_t* picked, float* lower, float* upper, float* const* lb, float* const* ub,
               int8_t* mask, int32_t* decision, cudaStream_t st, int64_t* launches) {
    *picked = 0;
    if (max_B <= 0 || q->size == 0) return GNNB_OK;
    const uint64_t* keys = q->keys[q->cur] + q->head;
    const int32_t* slots = q->slots[q->cur] + q->head;
    k_q_lower_bound<<<1, 1, 0, st>>>(keys, q->size, make_key(threshold, 0xFFFFFFFFu), q->d_count);    // smallest key with this bound
    int32_t n_valid = 0;
    QCU(cudaMemcpyAsync(&n_valid, q->d_count, sizeof n_valid, cudaMemcpyDeviceToHost, st));
    QCU(cudaStreamSynchronize(st));
    ++*launches;
    const int n = n_valid < max_B ? n_valid : max_B;
    if (n > 0) {
        ScatterArgs a{};
        a.ncopy = 0;
        for (int k = 0; k <= q->L + 1; ++k) {
            if (lb && lb[k]) a.c[a.ncopy++] = RowCopy{q->p_lb, nullptr, lb[k], q->n[k], q->off[k]};
            if (ub && ub[k]) a.c[a.ncopy++] = RowCopy{q->p_ub, nullptr, ub[k], q->n[k], q->off[k]};
        }
        a.NB = q->NB; a.sel = nullptr; a.slots = slots; a.count = n;
        a.lower_out = lower; a.upper_out = upper; a.mask_out = mask; a.dec_out = decision;
        a.p_lower = q->p_lower; a.p_upper = q->p_upper; a.p_mask = q->p_mask; a.p_dec = q->p_dec; a.n_hidden = q->n_hidden;
        k_q_rows<false><<<dim3((unsigned)n, (unsigned)a.ncopy + 1), 256, 0, st>>>(a);
        ++*launches;
    }
    // the picked domains leave the queue; when fewer than max_B are below the threshold the next pick_out of the reference pops
    // (discards) everything that is left (`discard`), or the rest is kept for a later prune
    const int64_t gone = (n < max_B && discard) ? q->size : n;
    if (gone == 0) return GNNB_OK;
    k_q_release<<<(unsigned)((gone + 255) / 256), 256, 0, st>>>(slots, gone, q->free_stack, q->n_free);
    ++*launches;
    q->n_free += gone; q->head += gone; q->size -= gone;
    if (q->size == 0) q->head = 0;
    *picked = n;
    QCU(cudaGetLastError());
    return GNNB_OK;
}

int queue_prune(DomainQueue* q, float threshold, cudaStream_t st, int64_t* launches) {
    if (q->size == 0) return GNNB_OK;
    const uint64_t* keys = q->keys[q->cur] + q->head;
    const int32_t* slots = q->slots[q->cur] + q->head;
    k_q_lower_bound<<<1, 1, 0, st>>>(keys, q->size, make_key(threshold, 0xFFFFFFFFu), q->d_count);
    int32_t n_valid = 0;
    QCU(cudaMemcpyAsync(&n_valid, q->d_count, sizeof n_valid, cudaMemcpyDeviceToHost, st));
    QCU(cudaStreamSynchronize(st));
    ++*launches;
    const int64_t gone = q->size - n_valid;
    if (gone > 0) {
        k_q_release<<<(unsigned)((gone + 255) / 256), 256, 0, st>>>(slots + n_valid, gone, q->free_stack, q->n_free);
        ++*launches;
        q->n_free += gone; q->size = n_valid;
        if (q->size == 0) q->head = 0;
    }
    QCU(cudaGetLastError());
    return GNNB_OK;
}

float* queue_stage(DomainQueue* q, size_t bytes) { return ensure_stage(q, bytes) == GNNB_OK ? q->stage : nullptr; }
int64_t queue_capacity(const DomainQueue* q) { return q->cap; }

}  // namespace gnnb
