// tcgen05 path (placeholder until the kernels land): every entry reports B200_ERR_UNSUPPORTED.
#include "tc_api.cuh"

namespace b200 {
int64_t tc_plan(const MlpShape&, const MlpShape&, int64_t rows_map, int64_t rows_atlas, char* base, TcPlan* out) {
  if (out) { out->base = base; out->bytes = 0; out->rows_map = rows_map; out->rows_atlas = rows_atlas; }
  return 0;
}
static int unsupported() { set_error("tensor-core path not built"); return B200_ERR_UNSUPPORTED; }
int tc_atlas_forward(const TcStep&, cudaStream_t) { return unsupported(); }
int tc_atlas_backward(const TcStep&, cudaStream_t) { return unsupported(); }
int tc_mapping_forward(const TcStep&, cudaStream_t) { return unsupported(); }
int tc_mapping_backward(const TcStep&, cudaStream_t) { return unsupported(); }
}  // namespace b200
