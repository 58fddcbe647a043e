// tcgen05 node kernels — placeholder until the tensor path lands (the API refuses math = TC meanwhile).
#include "gnnb_common.cuh"

namespace gnnb {
bool tc_available() { return false; }
int tc_init() { return 0; }
int64_t tc_packed_elems(int K) { (void)K; return 0; }
int64_t tc_pack_weight(const float*, int, uint16_t*) { return 0; }
void tc_relax(const GnnParams&, const NodeInputs&, float*, float*, cudaStream_t, int64_t*) {}
void tc_update(const GnnParams&, bool, const float*, const float*, const float*, const float*, float*, float*, int, int64_t,
               int64_t, int64_t, unsigned long long*, cudaStream_t, int64_t*) {}
void tc_input_embed(const GnnParams&, const float*, const float*, const float*, float*, int64_t, cudaStream_t, int64_t*) {}
void tc_input_update(const GnnParams&, const float*, const float*, const float*, float*, int64_t, cudaStream_t, int64_t*) {}
}  // namespace gnnb
