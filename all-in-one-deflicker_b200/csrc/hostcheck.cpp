// Host build of the loss-head arithmetic (loss_math.h) for the CPU tests: lets pytest compare the
// hand-derived forward/backward with autograd of the oracle without a GPU.  Not used by the product.
#include <stdint.h>
#include "loss_math.h"
#include "seg_loss_math.h"

extern "C" {

// in: float[37] = uv[9][2], y[3][3], rgb[3], dx[3], dy[3], wf, wb ; cfg: float[12]
// out: float[33] = duv[9][2], dy[3][3], rgb, grad, rig, rig_global, flow_f, flow_b
void b200_host_sample_loss(const float* in, const float* cfg, float* out) {
  b200::SampleIn s;
  int k = 0;
  for (int g = 0; g < 9; ++g) for (int c = 0; c < 2; ++c) s.uv[g][c] = in[k++];
  for (int g = 0; g < 3; ++g) for (int c = 0; c < 3; ++c) s.y[g][c] = in[k++];
  for (int c = 0; c < 3; ++c) s.rgb[c] = in[k++];
  for (int c = 0; c < 3; ++c) s.dx[c] = in[k++];
  for (int c = 0; c < 3; ++c) s.dy[c] = in[k++];
  s.wf = in[k++]; s.wb = in[k++];
  b200::LossConfig c;
  c.larger_dim = cfg[0]; c.uv_scale = cfg[1]; c.d_local = cfg[2]; c.d_global = cfg[3];
  c.c_rgb = cfg[4]; c.c_grad = cfg[5]; c.c_rig = cfg[6]; c.c_rig_global = cfg[7]; c.c_flow = cfg[8];
  c.with_global = cfg[7] != 0.0f; c.inv_batch = cfg[9]; c.inv_nf = cfg[10]; c.inv_nb = cfg[11];
  b200::SampleOut o;
  b200::sample_loss(s, c, o);
  k = 0;
  for (int g = 0; g < 9; ++g) for (int cc = 0; cc < 2; ++cc) out[k++] = o.duv[g][cc];
  for (int g = 0; g < 3; ++g) for (int cc = 0; cc < 3; ++cc) out[k++] = o.dy[g][cc];
  out[k++] = o.rgb; out[k++] = o.grad; out[k++] = o.rig; out[k++] = o.rig_global;
  out[k++] = o.flow_f; out[k++] = o.flow_b;
}

// in: float[71] = uv1[9][2], uv2[9][2], ar[5], y[6][3], rgb[3], dx[3], dy[3], a_gt, wf, wb
// cfg: float[17] = larger_dim, uv_scale, d_local, d_global, c_rgb, c_grad, c_rig, c_rig_global1, c_rig_global2, c_flow,
//                  c_alpha_flow, c_sparsity, c_boot, with_global, inv_batch, inv_nf, inv_nb
// out: float[73] = duv1[9][2], duv2[9][2], dar[5], dy[6][3], val[14]
void b200_host_seg_sample_loss(const float* in, const float* cfg, float* out) {
  b200::SegSampleIn s;
  int k = 0;
  for (int g = 0; g < 9; ++g) for (int c = 0; c < 2; ++c) s.uv1[g][c] = in[k++];
  for (int g = 0; g < 9; ++g) for (int c = 0; c < 2; ++c) s.uv2[g][c] = in[k++];
  for (int a = 0; a < 5; ++a) s.ar[a] = in[k++];
  for (int g = 0; g < 6; ++g) for (int c = 0; c < 3; ++c) s.y[g][c] = in[k++];
  for (int c = 0; c < 3; ++c) s.rgb[c] = in[k++];
  for (int c = 0; c < 3; ++c) s.dx[c] = in[k++];
  for (int c = 0; c < 3; ++c) s.dy[c] = in[k++];
  s.a_gt = in[k++]; s.wf = in[k++]; s.wb = in[k++];
  b200::SegLossConfig c;
  c.larger_dim = cfg[0]; c.uv_scale = cfg[1]; c.d_local = cfg[2]; c.d_global = cfg[3];
  c.c_rgb = cfg[4]; c.c_grad = cfg[5]; c.c_rig = cfg[6]; c.c_rig_global1 = cfg[7]; c.c_rig_global2 = cfg[8];
  c.c_flow = cfg[9]; c.c_alpha_flow = cfg[10]; c.c_sparsity = cfg[11]; c.c_boot = cfg[12];
  c.with_global = cfg[13] != 0.0f; c.inv_batch = cfg[14]; c.inv_nf = cfg[15]; c.inv_nb = cfg[16];
  b200::SegSampleOut o;
  b200::seg_sample_loss(s, c, o);
  k = 0;
  for (int g = 0; g < 9; ++g) for (int cc = 0; cc < 2; ++cc) out[k++] = o.duv1[g][cc];
  for (int g = 0; g < 9; ++g) for (int cc = 0; cc < 2; ++cc) out[k++] = o.duv2[g][cc];
  for (int a = 0; a < 5; ++a) out[k++] = o.dar[a];
  for (int g = 0; g < 6; ++g) for (int cc = 0; cc < 3; ++cc) out[k++] = o.dy[g][cc];
  for (int v = 0; v < b200::SV_COUNT; ++v) out[k++] = o.val[v];
}

void b200_host_norm_coords(const float* v, int64_t n, float half_extent, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = b200::norm_coord(v[i], half_extent);
}

float b200_host_pe_freq(int k) { return b200::pe_freq(k); }

float b200_host_pretrain(float xn, float yn, const float* uv, float uv_scale, float inv_batch, float* g) {
  return b200::pretrain_term(xn, yn, uv, uv_scale, inv_batch, g);
}

}
