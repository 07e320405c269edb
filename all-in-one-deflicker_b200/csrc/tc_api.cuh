// Interface between c_api.cu and the tcgen05 path (mlp_tc.cu).
#pragma once
#include "common.cuh"

namespace b200 {

// Buffers of the tensor-core path, carved from the caller's workspace (see mlp_tc.cu).
struct TcPlan {
  char* base = nullptr;
  int64_t bytes = 0;
  int64_t rows_map = 0, rows_atlas = 0;
};

struct TcStep {
  const MlpShape* ms; const MlpShape* as;
  const TcPlan* plan;
  const float* params; float* grads;
  const float* x_map;        // [n_groups*cap][4]
  float* uv;                 // [n_groups*cap][2]  mapping output
  float* y_atlas;            // [3*cap][3]         atlas output
  const float* d_uv;         // [n_groups*cap][2]  direct gradient of the loss head
  const float* d_y;          // [3*cap][3]
  int cap, n_groups;
  const int* counters;
  int flow_groups;           // 1: groups 5 / 6 of the mapping batch hold counters[5] / counters[6] compacted rows
};

int64_t tc_plan(const MlpShape& ms, const MlpShape& as, int64_t rows_map, int64_t rows_atlas, char* base,
                TcPlan* out);
int tc_begin_step(const TcStep& s, cudaStream_t st);       // optional: start the weight-image preparation early (side stream)
int tc_atlas_forward(const TcStep& s, cudaStream_t st);    // mapping on all groups, atlas on groups 0..2
int tc_atlas_backward(const TcStep& s, cudaStream_t st);   // all parameter gradients
int tc_mapping_forward(const TcStep& s, cudaStream_t st);  // pre-training: mapping only
int tc_mapping_backward(const TcStep& s, cudaStream_t st);

// inference (render): x_map [rows][4] -> uv [rows][2] -> y [rows][3]; rows a multiple of 128; ws >= tc_infer_workspace_bytes
int64_t tc_infer_workspace_bytes(const MlpShape& ms, const MlpShape& as);
int tc_infer_forward(const MlpShape& ms, const MlpShape& as, const float* params, const float* x_map, float* uv,
                     float* y, int64_t rows, char* ws, cudaStream_t st);

// stand-alone IMLP (one network, autograd): see mlp_tc.cu
int64_t tc_single_workspace_bytes(const MlpShape& sh, bool is_atlas, int64_t rows);
// persistent: the caller keeps this workspace and these parameter / gradient buffers across calls -> job tables are
// cached in their own device allocations and the calls become graph-capturable after one eager call
int tc_single_forward(const MlpShape& sh, bool is_atlas, const float* params, const float* x, float* y, int64_t rows,
                      bool training, char* ws, bool persistent, cudaStream_t st);
int tc_single_backward(const MlpShape& sh, bool is_atlas, const float* params, float* grads, const float* x,
                       const float* y, const float* dy, float* d_in, int* gmax2, int64_t rows, char* ws, bool persistent,
                       cudaStream_t st);

// Scope guard used by entry points that own a persistent workspace (the segmentation step): while alive,
// b200_mlp_forward / backward calls made by this thread use the cached-table path above.
struct PersistentWorkspaceScope {
  PersistentWorkspaceScope();
  ~PersistentWorkspaceScope();
  bool prev;
};

int tc_debug_wgrad(long long* cycles, int* shapes, int max_ctas);

}  // namespace b200
