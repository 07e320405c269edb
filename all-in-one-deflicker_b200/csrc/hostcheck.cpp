// Host build of the loss-head arithmetic (loss_math.h) for the CPU tests: lets pytest compare the
// hand-derived forward/backward with autograd of the oracle without a GPU.  Not used by the product.
#include <stdint.h>
#include "loss_math.h"

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

void b200_host_norm_coords(const float* v, int64_t n, float half_extent, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = b200::norm_coord(v[i], half_extent);
}

float b200_host_pe_freq(int k) { return b200::pe_freq(k); }

float b200_host_pretrain(float xn, float yn, const float* uv, float uv_scale, float inv_batch, float* g) {
  return b200::pretrain_term(xn, yn, uv, uv_scale, inv_batch, g);
}

}
