// Per-sample arithmetic of the loss head of the SEGMENTATION variant of the stage-1 loop (two mapping networks,
// one alpha network, one atlas sampled in two quadrants), shared by seg.cu and the host check library.
//
// Reference being restated (paths relative to the reference root):
//   alpha range, composite, RGB + sparsity     src/stage1_neural_atlas_seg.py:226-252
//   two-layer gradient loss                    src/models/stage_1/loss_utils.py:173-224
//   rigidity (x2, local and global)            src/models/stage_1/loss_utils.py:227-278
//   alpha-weighted optical-flow loss (x2)      src/models/stage_1/loss_utils.py:299-322 (use_alpha=True)
//   alpha optical-flow loss                    src/models/stage_1/loss_utils.py:385-408
//   bootstrapping BCE                          src/stage1_neural_atlas_seg.py:306-307
//   total                                      src/stage1_neural_atlas_seg.py:309-315
#pragma once
#include "loss_math.h"

namespace b200 {

// alpha-network evaluations of one sample: at the pixel, its two +1 neighbours and its two flow-matched points
enum AlphaRow { A_BASE = 0, A_XP1 = 1, A_YP1 = 2, A_FWD = 3, A_BWD = 4, A_COUNT = 5 };
// atlas evaluations: layer 1 (foreground mapping, quadrant +0.5) then layer 2 (background, -0.5), each at BASE/XP1/YP1
constexpr int SEG_ATLAS_ROWS = 6;

struct SegLossConfig {
  float larger_dim, uv_scale, d_local, d_global;
  float c_rgb, c_grad, c_rig, c_rig_global1, c_rig_global2, c_flow, c_alpha_flow, c_sparsity, c_boot;
  int with_global;
  float inv_batch, inv_nf, inv_nb;
};

struct SegSampleIn {
  float uv1[G_COUNT][2], uv2[G_COUNT][2];   // tanh outputs of the two mapping networks
  float ar[A_COUNT];                        // tanh outputs of the alpha network
  float y[SEG_ATLAS_ROWS][3];               // tanh outputs of the atlas network
  float rgb[3], dx[3], dy[3];
  float a_gt;                               // bootstrapping mask value
  float wf, wb;
};

enum SegValue { SV_RGB = 0, SV_GRAD, SV_SPARSITY, SV_RIG1, SV_RIG2, SV_RIGG1, SV_RIGG2, SV_FLOW1_F, SV_FLOW1_B,
                SV_FLOW2_F, SV_FLOW2_B, SV_AFLOW_F, SV_AFLOW_B, SV_BCE, SV_COUNT };

struct SegSampleOut {
  float duv1[G_COUNT][2], duv2[G_COUNT][2];  // direct paths only (the atlas path is added by backprop)
  float dar[A_COUNT];
  float dy[SEG_ATLAS_ROWS][3];
  float val[SV_COUNT];                       // un-normalised per-sample values
};

// tanh output -> (0.001, 0.991), three separately rounded steps (stage1_neural_atlas_seg.py:229-232)
B200_HD float seg_alpha(float raw) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(__fmul_rn(__fmul_rn(0.5f, __fadd_rn(raw, 1.0f)), 0.99f), 0.001f);
#else
  volatile float a = 0.5f * (raw + 1.0f);
  volatile float b = a * 0.99f;
  return b + 0.001f;
#endif
}
constexpr float SEG_DALPHA = 0.495f;         // d alpha / d raw

B200_HD void seg_sample_loss(const SegSampleIn& in, const SegLossConfig& c, SegSampleOut& out) {
  for (int g = 0; g < G_COUNT; ++g) out.duv1[g][0] = out.duv1[g][1] = out.duv2[g][0] = out.duv2[g][1] = 0.0f;
  for (int k = 0; k < SV_COUNT; ++k) out.val[k] = 0.0f;
  float a[A_COUNT], da[A_COUNT];
  for (int k = 0; k < A_COUNT; ++k) { a[k] = seg_alpha(in.ar[k]); da[k] = 0.0f; }
  const float al = a[A_BASE], ax = a[A_XP1], ay = a[A_YP1];
  // ---- colour terms
  const float w_rgb = c.c_rgb * c.inv_batch, w_grad = c.c_grad * c.inv_batch, w_sp = c.c_sparsity * c.inv_batch;
  for (int ch = 0; ch < 3; ++ch) {
    const float c1 = (in.y[0][ch] + 1.0f) * 0.5f, c1x = (in.y[1][ch] + 1.0f) * 0.5f, c1y = (in.y[2][ch] + 1.0f) * 0.5f;
    const float c2 = (in.y[3][ch] + 1.0f) * 0.5f, c2x = (in.y[4][ch] + 1.0f) * 0.5f, c2y = (in.y[5][ch] + 1.0f) * 0.5f;
    const float o = c1 * al + c2 * (1.0f - al);                 // :240
    const float ox = c1x * ax + c2x * (1.0f - ax);              // loss_utils.py:214-217
    const float oy = c1y * ay + c2y * (1.0f - ay);
    const float n = c1 * (1.0f - al);                           // :249
    const float e = o - in.rgb[ch];
    const float ex = in.dx[ch] - (ox - o), ey = in.dy[ch] - (oy - o);
    out.val[SV_RGB] += e * e;
    out.val[SV_GRAD] += ex * ex + ey * ey;
    out.val[SV_SPARSITY] += n * n;
    const float g_o = w_rgb * 2.0f * e + w_grad * 2.0f * (ex + ey);
    const float g_ox = -w_grad * 2.0f * ex, g_oy = -w_grad * 2.0f * ey;
    const float g_n = w_sp * 2.0f * n;
    out.dy[0][ch] = 0.5f * (g_o * al + g_n * (1.0f - al));
    out.dy[3][ch] = 0.5f * (g_o * (1.0f - al));
    out.dy[1][ch] = 0.5f * (g_ox * ax);
    out.dy[4][ch] = 0.5f * (g_ox * (1.0f - ax));
    out.dy[2][ch] = 0.5f * (g_oy * ay);
    out.dy[5][ch] = 0.5f * (g_oy * (1.0f - ay));
    da[A_BASE] += g_o * (c1 - c2) - g_n * c1;
    da[A_XP1] += g_ox * (c1x - c2x);
    da[A_YP1] += g_oy * (c1y - c2y);
  }
  // ---- rigidity of both mappings
  out.val[SV_RIG1] = rigidity_term(in.uv1[G_BASE], in.uv1[G_YMD], in.uv1[G_XMD], c.larger_dim, c.uv_scale, c.d_local,
                                   c.c_rig * c.inv_batch, out.duv1[G_BASE], out.duv1[G_YMD], out.duv1[G_XMD]);
  out.val[SV_RIG2] = rigidity_term(in.uv2[G_BASE], in.uv2[G_YMD], in.uv2[G_XMD], c.larger_dim, c.uv_scale, c.d_local,
                                   c.c_rig * c.inv_batch, out.duv2[G_BASE], out.duv2[G_YMD], out.duv2[G_XMD]);
  if (c.with_global) {
    out.val[SV_RIGG1] = rigidity_term(in.uv1[G_BASE], in.uv1[G_YMG], in.uv1[G_XMG], c.larger_dim, c.uv_scale,
                                      c.d_global, c.c_rig_global1 * c.inv_batch, out.duv1[G_BASE], out.duv1[G_YMG],
                                      out.duv1[G_XMG]);
    out.val[SV_RIGG2] = rigidity_term(in.uv2[G_BASE], in.uv2[G_YMG], in.uv2[G_XMG], c.larger_dim, c.uv_scale,
                                      c.d_global, c.c_rig_global2 * c.inv_batch, out.duv2[G_BASE], out.duv2[G_YMG],
                                      out.duv2[G_XMG]);
  }
  // ---- flow terms: mapping 1 weighted by alpha, mapping 2 by (1 - alpha); alpha agreement along the flow
  for (int dir = 0; dir < 2; ++dir) {
    const bool on = dir == 0 ? in.wf != 0.0f : in.wb != 0.0f;
    if (!on) continue;
    const int g = dir == 0 ? G_FWD : G_BWD, ak = dir == 0 ? A_FWD : A_BWD;
    const float inv_n = dir == 0 ? c.inv_nf : c.inv_nb;
    const float wm = 0.5f * c.c_flow * inv_n;
    const float l1 = flow_term(in.uv1[G_BASE], in.uv1[g], c.larger_dim, c.uv_scale, wm * al, out.duv1[G_BASE], out.duv1[g]);
    const float l2 = flow_term(in.uv2[G_BASE], in.uv2[g], c.larger_dim, c.uv_scale, wm * (1.0f - al), out.duv2[G_BASE],
                               out.duv2[g]);
    out.val[dir == 0 ? SV_FLOW1_F : SV_FLOW1_B] = l1 * al;
    out.val[dir == 0 ? SV_FLOW2_F : SV_FLOW2_B] = l2 * (1.0f - al);
    da[A_BASE] += wm * (l1 - l2);
    // forward: |alpha - alpha_match|, backward: |alpha_match - alpha| (loss_utils.py:395,406)
    const float d = dir == 0 ? al - a[ak] : a[ak] - al;
    const float sg = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
    const float wa = 0.5f * c.c_alpha_flow * inv_n * sg;
    out.val[dir == 0 ? SV_AFLOW_F : SV_AFLOW_B] = fabsf(d);
    if (dir == 0) { da[A_BASE] += wa; da[ak] -= wa; } else { da[ak] += wa; da[A_BASE] -= wa; }
  }
  // ---- bootstrapping BCE against the segmentation mask
  out.val[SV_BCE] = -in.a_gt * logf(al) - (1.0f - in.a_gt) * logf(1.0f - al);
  da[A_BASE] += c.c_boot * c.inv_batch * (-in.a_gt / al + (1.0f - in.a_gt) / (1.0f - al));
  for (int k = 0; k < A_COUNT; ++k) out.dar[k] = SEG_DALPHA * da[k];
}

}  // namespace b200
