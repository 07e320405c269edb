// Stand-alone loss heads behind the reference's three loss *functions* (the drop-in boundary of SURVEY.md §8b):
//   get_gradient_loss_single   src/models/stage_1/loss_utils.py:134-170
//   get_rigidity_loss          src/models/stage_1/loss_utils.py:227-278
//   get_optical_flow_loss      src/models/stage_1/loss_utils.py:299-322  (one direction per call)
// The fused training step (b200_atlas_loss_grad) does not use these: it evaluates all four terms in one
// kernel.  These exist so that a caller who keeps the reference's function-level structure (its own loop, its
// own IMLP objects) still runs the loss arithmetic — value AND gradient — on this library: each call writes the
// scalar the reference function returns and d(scalar)/d(inputs); the Python wrapper hands those to autograd.
// The per-sample arithmetic is loss_math.h, the same code the fused kernel runs.
#include "common.cuh"
#include "loss_math.h"

namespace b200 {

template <int NV>
__device__ __forceinline__ void block_sum_to(float (&v)[NV], float* dst, float scale) {
  __shared__ float red[NV][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    float x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[q][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[q][w];
      atomicAdd(dst + q, t * scale);
    }
  }
}

// mean_b( ||dx_gt - (rgb_xp - rgb)||^2 + ||dy_gt - (rgb_yp - rgb)||^2 )      loss_utils.py:165-170
__global__ void gradient_head_kernel(const float* __restrict__ rgb, const float* __restrict__ rgb_xp,
                                     const float* __restrict__ rgb_yp, const float* __restrict__ dx_gt,
                                     const float* __restrict__ dy_gt, int64_t n, float inv_n, float* __restrict__ loss,
                                     float* __restrict__ d_rgb, float* __restrict__ d_xp, float* __restrict__ d_yp) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float part[1] = {0.f};
  if (s < n) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float o = rgb[s * 3 + c];
      const float ex = dx_gt[s * 3 + c] - (rgb_xp[s * 3 + c] - o);
      const float ey = dy_gt[s * 3 + c] - (rgb_yp[s * 3 + c] - o);
      part[0] += ex * ex + ey * ey;
      d_rgb[s * 3 + c] = 2.0f * (ex + ey) * inv_n;
      d_xp[s * 3 + c] = -2.0f * ex * inv_n;
      d_yp[s * 3 + c] = -2.0f * ey * inv_n;
    }
  }
  block_sum_to(part, loss, inv_n);
}

// uv_p rows [0, n): mapping at (x, y-d, t); rows [n, 2n): mapping at (x-d, y, t)       loss_utils.py:230-240
__global__ void rigidity_head_kernel(const float* __restrict__ uv, const float* __restrict__ uv_p, int64_t n, float L,
                                     float uv_scale, float d, float inv_n, float* __restrict__ per_sample,
                                     float* __restrict__ loss, float* __restrict__ d_uv, float* __restrict__ d_uv_p) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float part[1] = {0.f};
  if (s < n) {
    const float u0[2] = {uv[s * 2], uv[s * 2 + 1]};
    const float ua[2] = {uv_p[s * 2], uv_p[s * 2 + 1]};
    const float ub[2] = {uv_p[(n + s) * 2], uv_p[(n + s) * 2 + 1]};
    float g0[2] = {0.f, 0.f}, ga[2] = {0.f, 0.f}, gb[2] = {0.f, 0.f};
    const float v = rigidity_term(u0, ua, ub, L, uv_scale, d, inv_n, g0, ga, gb);
    part[0] = v;
    if (per_sample) per_sample[s] = v;
    d_uv[s * 2] = g0[0]; d_uv[s * 2 + 1] = g0[1];
    d_uv_p[s * 2] = ga[0]; d_uv_p[s * 2 + 1] = ga[1];
    d_uv_p[(n + s) * 2] = gb[0]; d_uv_p[(n + s) * 2 + 1] = gb[1];
  }
  block_sum_to(part, loss, inv_n);
}

// mean_rows ||uv_match - uv_rel||_2 * resx / (2 uv_mapping_scale)                      loss_utils.py:306-309
__global__ void flow_head_kernel(const float* __restrict__ uv_rel, const float* __restrict__ uv_match, int64_t n, float L,
                                 float uv_scale, float inv_n, float* __restrict__ loss, float* __restrict__ d_rel,
                                 float* __restrict__ d_match) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float part[1] = {0.f};
  if (s < n) {
    const float u0[2] = {uv_rel[s * 2], uv_rel[s * 2 + 1]};
    const float um[2] = {uv_match[s * 2], uv_match[s * 2 + 1]};
    float g0[2] = {0.f, 0.f}, gm[2] = {0.f, 0.f};
    part[0] = flow_term(u0, um, L, uv_scale, inv_n, g0, gm);
    d_rel[s * 2] = g0[0]; d_rel[s * 2 + 1] = g0[1];
    d_match[s * 2] = gm[0]; d_match[s * 2 + 1] = gm[1];
  }
  block_sum_to(part, loss, inv_n);
}

// mean_rows w * ||uv_match - uv_rel||_2 * resx / (2 uv_mapping_scale): the alpha-weighted form of the segmentation
// variant (loss_utils.py:316-318, use_alpha=True); also d/d w
__global__ void flow_head_weighted_kernel(const float* __restrict__ uv_rel, const float* __restrict__ uv_match,
                                          const float* __restrict__ w, int64_t n, float L, float uv_scale, float inv_n,
                                          float* __restrict__ loss, float* __restrict__ d_rel, float* __restrict__ d_match,
                                          float* __restrict__ d_w) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float part[1] = {0.f};
  if (s < n) {
    const float u0[2] = {uv_rel[s * 2], uv_rel[s * 2 + 1]};
    const float um[2] = {uv_match[s * 2], uv_match[s * 2 + 1]};
    const float ws = w[s];
    float g0[2] = {0.f, 0.f}, gm[2] = {0.f, 0.f};
    const float l = flow_term(u0, um, L, uv_scale, inv_n * ws, g0, gm);
    part[0] = l * ws;
    d_rel[s * 2] = g0[0]; d_rel[s * 2 + 1] = g0[1];
    d_match[s * 2] = gm[0]; d_match[s * 2 + 1] = gm[1];
    d_w[s] = l * inv_n;
  }
  block_sum_to(part, loss, inv_n);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_gradient_loss_head(const float* rgb, const float* rgb_xp, const float* rgb_yp, const float* dx_gt,
                            const float* dy_gt, int64_t n, float* loss, float* d_rgb, float* d_rgb_xp,
                            float* d_rgb_yp, void* stream) {
  B200_REQUIRE(rgb && rgb_xp && rgb_yp && dx_gt && dy_gt && loss && d_rgb && d_rgb_xp && d_rgb_yp, "null pointer");
  B200_REQUIRE(n > 0 && n < (1ll << 31), "row count out of range: %lld", (long long)n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, st));
  gradient_head_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rgb, rgb_xp, rgb_yp, dx_gt, dy_gt, n,
                                                                    1.0f / (float)n, loss, d_rgb, d_rgb_xp, d_rgb_yp);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_rigidity_loss_head(const float* uv, const float* uv_p, int64_t n, float resx, float uv_mapping_scale,
                            float derivative_amount, float* per_sample, float* loss, float* d_uv, float* d_uv_p,
                            void* stream) {
  B200_REQUIRE(uv && uv_p && loss && d_uv && d_uv_p, "null pointer");
  B200_REQUIRE(n > 0 && n < (1ll << 30), "row count out of range: %lld", (long long)n);
  B200_REQUIRE(resx > 0.f && uv_mapping_scale != 0.f && derivative_amount != 0.f, "bad rigidity geometry");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, st));
  rigidity_head_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(uv, uv_p, n, resx, uv_mapping_scale,
                                                                    derivative_amount, 1.0f / (float)n, per_sample, loss,
                                                                    d_uv, d_uv_p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_flow_loss_head(const float* uv_rel, const float* uv_match, int64_t n, float resx, float uv_mapping_scale,
                        float* loss, float* d_uv_rel, float* d_uv_match, void* stream) {
  B200_REQUIRE(loss, "null pointer");
  B200_REQUIRE(n >= 0 && n < (1ll << 31), "row count out of range: %lld", (long long)n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {                 // mean over an empty set: NaN, exactly as torch (loss_utils.py:320-322)
    B200_CHECK_CUDA(cudaMemsetAsync(loss, 0xFF, 4, st));     // 0xFFFFFFFF is a quiet NaN; capturable
    return B200_OK;
  }
  B200_REQUIRE(uv_rel && uv_match && d_uv_rel && d_uv_match, "null pointer");
  B200_REQUIRE(resx > 0.f && uv_mapping_scale != 0.f, "bad flow geometry");
  B200_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, st));
  flow_head_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(uv_rel, uv_match, n, resx, uv_mapping_scale,
                                                                1.0f / (float)n, loss, d_uv_rel, d_uv_match);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_flow_loss_head_weighted(const float* uv_rel, const float* uv_match, const float* w, int64_t n, float resx,
                                 float uv_mapping_scale, float* loss, float* d_uv_rel, float* d_uv_match, float* d_w,
                                 void* stream) {
  B200_REQUIRE(loss, "null pointer");
  B200_REQUIRE(n >= 0 && n < (1ll << 31), "row count out of range: %lld", (long long)n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {
    B200_CHECK_CUDA(cudaMemsetAsync(loss, 0xFF, 4, st));
    return B200_OK;
  }
  B200_REQUIRE(uv_rel && uv_match && w && d_uv_rel && d_uv_match && d_w, "null pointer");
  B200_REQUIRE(resx > 0.f && uv_mapping_scale != 0.f, "bad flow geometry");
  B200_CHECK_CUDA(cudaMemsetAsync(loss, 0, 4, st));
  flow_head_weighted_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(uv_rel, uv_match, w, n, resx, uv_mapping_scale,
                                                                         1.0f / (float)n, loss, d_uv_rel, d_uv_match, d_w);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
