// Per-sample arithmetic of the stage-1 atlas loss head, shared by the CUDA kernels
// (atlas_kernels.cu) and by the host-side check library (hostcheck.cpp) that the CPU tests use
// to compare this hand-derived forward + backward against autograd of the oracle.
//
// Reference being restated (paths relative to the reference root):
//   RGB loss            src/stage1_neural_atlas.py:181,194
//   gradient loss       src/models/stage_1/loss_utils.py:134-170
//   rigidity loss       src/models/stage_1/loss_utils.py:227-278
//   optical-flow loss   src/models/stage_1/loss_utils.py:299-356
//   coordinate scaling  src/stage1_neural_atlas.py:168-171 (int64 / numpy double -> fp32 divide)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

// Row groups of the mapping network batch.  Group g occupies rows [g*cap, (g+1)*cap).
enum RowGroup {
  G_BASE = 0,   // (x, y, t)
  G_XP1 = 1,    // (x+1, y, t)            gradient loss
  G_YP1 = 2,    // (x, y+1, t)            gradient loss
  G_YMD = 3,    // (x, y-d, t)            rigidity, d = derivative_amount
  G_XMD = 4,    // (x-d, y, t)
  G_FWD = 5,    // (x+fx, y+fy, t+1)      forward flow match
  G_BWD = 6,    // (x+bx, y+by, t-1)      backward flow match
  G_YMG = 7,    // (x, y-D, t)            global rigidity, D = global_rigidity_derivative_amount_fg
  G_XMG = 8,    // (x-D, y, t)
  G_COUNT = 9
};

// fp32 division / subtraction without contraction, the rounding sequence torch's CPU kernels
// produce for  int64_tensor / np.float64(L/2) - 1  (checked bit-exactly in tests/test_host_math.py)
B200_HD float norm_coord(float v, float half_extent) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(__fdiv_rn(v, half_extent), 1.0f);
#else
  volatile float q = v / half_extent;
  return q - 1.0f;
#endif
}

struct LossConfig {
  float larger_dim;     // max(resx, resy) as float
  float uv_scale;       // uv_mapping_scale (0.8)
  float d_local;        // derivative_amount
  float d_global;       // global_rigidity_derivative_amount_fg
  float c_rgb, c_grad, c_rig, c_rig_global, c_flow;
  int with_global;      // i <= stop_global_rigidity
  float inv_batch;      // 1 / samples_batch (the global batch, also under frame sharding)
  float inv_nf, inv_nb; // 1 / number of valid forward / backward flow rows (0 when none)
};

struct SampleIn {
  float uv[G_COUNT][2]; // mapping outputs (tanh applied)
  float y[3][3];        // atlas outputs (tanh applied) for groups BASE, XP1, YP1
  float rgb[3], dx[3], dy[3];
  float wf, wb;         // 1 if the forward / backward flow of this pixel is valid
};

struct SampleOut {
  float duv[G_COUNT][2]; // d total / d uv   (direct paths only; the atlas path is added by backprop)
  float dy[3][3];        // d total / d y
  float rgb, grad, rig, rig_global, flow_f, flow_b;  // un-normalised per-sample loss values
};

// rigidity term for one derivative amount.  uv0: base, uva: (x, y-d), uvb: (x-d, y).
// Adds coeff * d(loss)/d(uv*) into g0/ga/gb and returns the per-sample loss value.
B200_HD float rigidity_term(const float* uv0, const float* uva, const float* uvb, float L, float uv_scale,
                            float d, float coeff, float* g0, float* ga, float* gb) {
  // loss_utils.py:246-256 : (u(x,y) - u(neighbour)) * resx / 2 / uv_mapping_scale / derivative_amount
  const float p = (uv0[0] - uvb[0]) * L / 2.0f / uv_scale / d;   // du/dx
  const float q = (uv0[0] - uva[0]) * L / 2.0f / uv_scale / d;   // du/dy
  const float r = (uv0[1] - uvb[1]) * L / 2.0f / uv_scale / d;   // dv/dx
  const float s = (uv0[1] - uva[1]) * L / 2.0f / uv_scale / d;   // dv/dy
  const float S = L / 2.0f / uv_scale / d;
  // JtJ = J^T J, loss_utils.py:259
  const float A = p * p + r * r, Bc = p * q + r * s, D = q * q + s * s;
  const float a = A + 0.001f, dd = D + 0.001f;                   // :261-264
  const float det = a * dd - Bc * Bc;
  const float n1 = sqrtf(A * A + 2.0f * Bc * Bc + D * D);        // ||JtJ||_F        :273
  const float n2 = sqrtf(a * a + 2.0f * Bc * Bc + dd * dd);      // ||adj||_F
  const float inv_det = 1.0f / det;
  const float value = n1 + n2 * fabsf(inv_det);
  // backward.  sqrt at exactly 0 has no finite derivative (torch gives NaN there); we use 0.
  const float in1 = n1 > 0.0f ? 1.0f / n1 : 0.0f;
  const float in2 = n2 > 0.0f ? 1.0f / n2 : 0.0f;
  const float sgn = det >= 0.0f ? 1.0f : -1.0f;
  // F2 = n2 / |det| ; d det/da = dd, d det/ddd = a, d det/dBc = -2 Bc
  const float k = n2 * inv_det * inv_det * sgn;
  const float gA = A * in1 + (a * in2) * fabsf(inv_det) - k * dd;
  const float gD = D * in1 + (dd * in2) * fabsf(inv_det) - k * a;
  const float gB = 2.0f * Bc * in1 + (2.0f * Bc * in2) * fabsf(inv_det) + k * 2.0f * Bc;
  const float gp = 2.0f * p * gA + q * gB;
  const float gr = 2.0f * r * gA + s * gB;
  const float gq = 2.0f * q * gD + p * gB;
  const float gs = 2.0f * s * gD + r * gB;
  const float w = coeff * S;
  g0[0] += w * (gp + gq);  g0[1] += w * (gr + gs);
  gb[0] -= w * gp;         gb[1] -= w * gr;
  ga[0] -= w * gq;         ga[1] -= w * gs;
  return value;
}

// one direction of the flow term.  uvm: mapping output at the flow-matched point.
B200_HD float flow_term(const float* uv0, const float* uvm, float L, float uv_scale, float coeff,
                        float* g0, float* gm) {
  const float ex = uvm[0] - uv0[0], ey = uvm[1] - uv0[1];
  const float n = sqrtf(ex * ex + ey * ey);
  const float scale = L / (2.0f * uv_scale);                      // loss_utils.py:308
  const float inv = n > 0.0f ? 1.0f / n : 0.0f;                   // norm backward at 0 is 0 in torch
  const float w = coeff * scale * inv;
  gm[0] += w * ex;  gm[1] += w * ey;
  g0[0] -= w * ex;  g0[1] -= w * ey;
  return n * scale;
}

B200_HD void sample_loss(const SampleIn& in, const LossConfig& c, SampleOut& out) {
  for (int g = 0; g < G_COUNT; ++g) out.duv[g][0] = out.duv[g][1] = 0.0f;
  // ---- colour terms.  o = (y + 1) * 0.5  (src/stage1_neural_atlas.py:181)
  float v_rgb = 0.0f, v_grad = 0.0f;
  const float w_rgb = c.c_rgb * c.inv_batch, w_grad = c.c_grad * c.inv_batch;
  for (int ch = 0; ch < 3; ++ch) {
    const float o = (in.y[0][ch] + 1.0f) * 0.5f;
    const float ox = (in.y[1][ch] + 1.0f) * 0.5f;
    const float oy = (in.y[2][ch] + 1.0f) * 0.5f;
    const float e = o - in.rgb[ch];
    const float ex = in.dx[ch] - (ox - o);                       // loss_utils.py:165-169
    const float ey = in.dy[ch] - (oy - o);
    v_rgb += e * e;
    v_grad += ex * ex + ey * ey;
    // d/d o, chain 0.5 for y
    out.dy[0][ch] = 0.5f * (w_rgb * 2.0f * e + w_grad * 2.0f * (ex + ey));
    out.dy[1][ch] = 0.5f * (-w_grad * 2.0f * ex);
    out.dy[2][ch] = 0.5f * (-w_grad * 2.0f * ey);
  }
  out.rgb = v_rgb;
  out.grad = v_grad;
  // ---- rigidity (local always, global while i <= stop_global_rigidity)
  out.rig = rigidity_term(in.uv[G_BASE], in.uv[G_YMD], in.uv[G_XMD], c.larger_dim, c.uv_scale, c.d_local,
                          c.c_rig * c.inv_batch, out.duv[G_BASE], out.duv[G_YMD], out.duv[G_XMD]);
  out.rig_global = 0.0f;
  if (c.with_global)
    out.rig_global = rigidity_term(in.uv[G_BASE], in.uv[G_YMG], in.uv[G_XMG], c.larger_dim, c.uv_scale,
                                   c.d_global, c.c_rig_global * c.inv_batch, out.duv[G_BASE],
                                   out.duv[G_YMG], out.duv[G_XMG]);
  // ---- optical flow: 0.5 * mean_fwd + 0.5 * mean_bwd over the valid rows (loss_utils.py:316-322)
  out.flow_f = out.flow_b = 0.0f;
  if (in.wf != 0.0f)
    out.flow_f = flow_term(in.uv[G_BASE], in.uv[G_FWD], c.larger_dim, c.uv_scale,
                           0.5f * c.c_flow * c.inv_nf, out.duv[G_BASE], out.duv[G_FWD]);
  if (in.wb != 0.0f)
    out.flow_b = flow_term(in.uv[G_BASE], in.uv[G_BWD], c.larger_dim, c.uv_scale,
                           0.5f * c.c_flow * c.inv_nb, out.duv[G_BASE], out.duv[G_BWD]);
}

// pre-training loss of the mapping network (src/models/stage_1/unwrap_utils.py:192):
// mean_b || xy * uv_scale - uv ||_2.  Returns the per-sample value, writes d/d uv (times inv_batch).
B200_HD float pretrain_term(float xn, float yn, const float* uv, float uv_scale, float inv_batch, float* g) {
  const float ex = xn * uv_scale - uv[0], ey = yn * uv_scale - uv[1];
  const float n = sqrtf(ex * ex + ey * ey);
  const float inv = n > 0.0f ? 1.0f / n : 0.0f;
  g[0] = -inv_batch * ex * inv;
  g[1] = -inv_batch * ey * inv;
  return n;
}

// Positional encoding (src/models/stage_1/implicit_neural_networks.py:9-13,34).  Column layout for
// input dim D and F frequencies: col = k*2D + j -> sin(x_j * b_k), col = k*2D + D + j -> cos(x_j * b_k)
// with b_k = fp32(2^k * pi).
B200_HD float pe_freq(int k) {
  return (float)(3.141592653589793 * (double)(1 << k));
}

}  // namespace b200
