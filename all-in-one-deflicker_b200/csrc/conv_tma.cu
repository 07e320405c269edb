// TMA-fed tcgen05 convolution (fp16 operands, fp32 accumulation in TMEM) — no im2col gather at all.
//
// With a channels-last fp16 copy of the input, the A operand of the implicit GEMM for ONE filter tap (ky,kx)
// and one block of 64 input channels is a plain 2-D box of the tensor: 128 consecutive output pixels of one
// image row x 64 channels.  A tiled TMA load with the 128-byte swizzle lands it in shared memory in exactly the
// canonical K-major layout tcgen05.mma reads; the tap only shifts the box coordinates (the pixel dimension is
// not the innermost one, so any shift is legal for the TMA unit).  The reduction runs over (tap, channel block).
//
// Two kernels per convolution:
//   conv_pack_input_kernel  fp32 NCHW (channel slice) -> fp16 NHWC (channels padded to 64) through a shared-
//                           memory transpose, with the padding (zeros / reflection), the nearest x2 upsampling
//                           and — for stride 2 — a split into the 4 pixel phases materialised, so that the
//                           convolution proper is a stride-1 "valid" one (HBM-bound)
//   conv2d_tma_kernel       persistent, warp 0 = TMA producer (1 activation box + 1 weight image per stage),
//                           warp 1 = tcgen05.mma issuer (two TMEM accumulators), warps 2-9 = epilogue
//                           (bias, activation, scale, residual, coalesced NCHW fp32 stores)
//
// Same operator semantics as b200_conv2d (reference: nn.Conv2d / ReflectionPad2d / nn.Upsample call sites in
// src/models/stage_1/core/update.py, src/models/network_filter.py, src/models/network_local.py).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200 {
using namespace ptx;

constexpr int TM_THREADS = 320;
constexpr int TM_MAX_A = 8, TM_MAX_B = 16, TM_MAX_ACC = 8;
constexpr int TM_BAR_BYTES = 1024;              // mbarriers + TMEM slot
constexpr int TM_SMEM_BUDGET = 225 * 1024;

struct ConvTmaArgs {
  B200ConvDesc d;
  const char* w_img; const float* bias; const float* res; float* y;
  int OH, OW, x_tiles, cchunks, n_chunks, n_tile, n_tiles_n, phases, shift, total_tiles;
  int a_rows, a_stage, n_a, n_b, n_acc, acc_stride;   // box rows, bytes per A stage, ring depths, TMEM ring
  int b_group, b_stage;                                // weight chunks (x taps) per B stage, bytes per B stage
  int split;                                           // 1: both operands are (hi, lo) fp16 pairs, 3 MMAs per product
  int resident;                                        // 1: the whole weight image stays in shared memory (narrow layers)
  // chained output: the result is ALSO (or only, y == nullptr) written as fp16 into the packed input of the next
  // convolution: [n][yp_hp2][yp_wp2][yp_cp] with its zero halo (yp_pad_h, yp_pad_w), at channel offset yp_c_off
  __half* yp; int yp_hp2, yp_wp2, yp_cp, yp_pad_h, yp_pad_w, yp_c_off;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int ACT>
__device__ __forceinline__ float tm_act(float v) {
  if (ACT == 1) return fmaxf(v, 0.f);
  if (ACT == 2) return v > 0.f ? v : 0.2f * v;
  if (ACT == 3) return 1.0f / (1.0f + expf(-v));
  if (ACT == 4) return tanhf(v);
  return v;
}

// 32 accumulator columns of one pixel -> bias, activation, scale, residual; NCHW fp32 stores (one channel plane apart,
// yq may be null) and / or 32 consecutive fp16 channels of the pixel's vector in the next layer's packed input (ypk)
template <int ACT>
__device__ __forceinline__ void tm_store32(const uint32_t (&raw)[32], const float* __restrict__ bias, float scale,
                                           const float* __restrict__ resp, float* __restrict__ yq, int64_t oplane,
                                           int ncols, bool live, int lane, __half* __restrict__ ypk) {
  const float bl = (bias && lane < ncols) ? __ldg(bias + lane) : 0.f;      // lane i holds bias[i] (ncols is warp-uniform)
  const int nvalid = live ? ncols : 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float radd[16];
    if (resp) {                                            // residual loads in flight before the first store
      const float* rp = resp + (int64_t)(h * 16) * oplane;
#pragma unroll
      for (int i = 0; i < 16; ++i, rp += oplane) radd[i] = (h * 16 + i) < nvalid ? __ldg(rp) : 0.f;
    }
    float* p = yq ? yq + (int64_t)(h * 16) * oplane : nullptr;
    float v16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float val = __uint_as_float(raw[h * 16 + i]) + __shfl_sync(0xffffffffu, bl, h * 16 + i);
      val = tm_act<ACT>(val) * scale;
      if (resp) val += radd[i];
      v16[i] = val;
      if (p && h * 16 + i < nvalid) p[(int64_t)i * oplane] = val;
    }
    if (ypk) {                                             // channels are consecutive in NHWC: two 16-byte stores
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        if (h * 16 + g8 * 8 < nvalid) {                    // ncols is a multiple of 8 on this path
          uint4 w;
          w.x = pack_half2(v16[g8 * 8 + 0], v16[g8 * 8 + 1]); w.y = pack_half2(v16[g8 * 8 + 2], v16[g8 * 8 + 3]);
          w.z = pack_half2(v16[g8 * 8 + 4], v16[g8 * 8 + 5]); w.w = pack_half2(v16[g8 * 8 + 6], v16[g8 * 8 + 7]);
          *reinterpret_cast<uint4*>(ypk + h * 16 + g8 * 8) = w;
        }
      }
    }
  }
}

// Reduction order shared by the producer, the MMA issuer and the weight images:
//   for ky, for xpar in [0, stride), for cc (64-channel block):   one activation box (all x shifts of that row)
//     for kx = xpar, xpar + stride, ... < KW:                      one weight chunk, A start shifted by kx>>shift rows
__global__ void __launch_bounds__(TM_THREADS, 1) conv2d_tma_kernel(const __grid_constant__ ConvTmaArgs a,
                                                                   const __grid_constant__ CUtensorMap xmap,
                                                                   const __grid_constant__ CUtensorMap xmap_lo) {
  extern __shared__ __align__(1024) char smem[];
  const int b_half = a.n_tile * 128;                    // one term of one weight chunk
  const int b_bytes = a.split ? 2 * b_half : b_half;
  const int a_half = a.a_stage >> 1;                    // split: hi box at +0, lo box at +a_half
  char* sA = smem;
  char* sB = smem + a.n_a * a.a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + a.n_b * a.b_stage);
  uint64_t* a_full = bars;                         // tx
  uint64_t* a_empty = a_full + TM_MAX_A;           // commit
  uint64_t* b_full = a_empty + TM_MAX_A;           // tx
  uint64_t* b_empty = b_full + TM_MAX_B;           // commit
  uint64_t* d_full = b_empty + TM_MAX_B;           // commit
  uint64_t* d_empty = d_full + TM_MAX_ACC;         // 128 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + TM_MAX_ACC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const B200ConvDesc& d = a.d;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("b200: conv smem not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < TM_MAX_A; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TM_MAX_B; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < TM_MAX_ACC; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int s = d.stride;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      if (a.resident) {                                    // one cout tile, all chunks: loaded once per CTA
        mbar_expect_tx(&b_full[0], a.n_chunks * b_bytes);
        for (int j = 0; j < a.n_chunks; ++j) bulk_g2s(sB + j * b_bytes, a.w_img + (int64_t)j * b_bytes, b_bytes, &b_full[0]);
      }
      for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        const int nt = t % a.n_tiles_n;
        int r = t / a.n_tiles_n;
        const int xb = r % a.x_tiles; r /= a.x_tiles;
        const int oy = r % a.OH, n = r / a.OH;
        const char* wsrc = a.w_img + (int64_t)nt * a.n_chunks * b_bytes;
        for (int ky = 0; ky < d.KH; ++ky)
          for (int xpar = 0; xpar < s && xpar < d.KW; ++xpar) {
            const int nsub = (d.KW - xpar + s - 1) / s;
            const int ph = a.phases == 4 ? ((ky & 1) * 2 + xpar) : 0;
            for (int cc = 0; cc < a.cchunks; ++cc) {
              mbar_wait(&a_empty[sa], pha ^ 1);
              mbar_expect_tx(&a_full[sa], a.a_rows * 128 * (1 + a.split));
              tma_load_4d(sA + sa * a.a_stage, &xmap, cc * 64, xb * 128, oy + (ky >> a.shift), n * a.phases + ph, &a_full[sa]);
              if (a.split)
                tma_load_4d(sA + sa * a.a_stage + a_half, &xmap_lo, cc * 64, xb * 128, oy + (ky >> a.shift), n * a.phases + ph,
                            &a_full[sa]);
              if (++sa == a.n_a) { sa = 0; pha ^= 1; }
              if (a.resident) continue;
              for (int i0 = 0; i0 < nsub; i0 += a.b_group) {
                const int gn = nsub - i0 < a.b_group ? nsub - i0 : a.b_group;
                mbar_wait(&b_empty[sb], phb ^ 1);
                mbar_expect_tx(&b_full[sb], gn * b_bytes);
                bulk_g2s(sB + sb * a.b_stage, wsrc, gn * b_bytes, &b_full[sb]);
                wsrc += gn * b_bytes;
                if (++sb == a.n_b) { sb = 0; phb ^= 1; }
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the (uniform) loop control so that descriptors stay in uniform registers; one elected
    // lane issues.  Ring positions are kept incrementally (no runtime divisions on this latency-critical path).
    const uint32_t idesc = make_idesc(128, a.n_tile, 0, 0);
    const uint64_t desc_hi = make_desc(0, 16, 1024);              // everything but the start address
    const bool leader = elect_one();
    int sa = 0, sb = 0, acc = 0;
    uint32_t pha = 0, phb = 0, phd = 0;
    if (a.resident) mbar_wait(&b_full[0], 0);
    const uint64_t db_res = desc_hi + (uint64_t)(smem_u32(sB) >> 4);
    const uint64_t db_step = (uint64_t)(b_bytes >> 4);
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
      mbar_wait(&d_empty[acc], phd ^ 1);
      tc_fence_after();
      const uint32_t dcol = tmem + acc * a.acc_stride;
      uint32_t accum = 0;
      uint64_t db_run = db_res;                            // resident mode: descriptor of the next weight chunk
      for (int ky = 0; ky < d.KH; ++ky)
        for (int xpar = 0; xpar < s && xpar < d.KW; ++xpar) {
          const int nsub = (d.KW - xpar + s - 1) / s;
          for (int cc = 0; cc < a.cchunks; ++cc) {
            mbar_wait(&a_full[sa], pha);
            const uint32_t pa = smem_u32(sA + sa * a.a_stage);
            const int left = d.Cin - cc * 64;
            const int ksteps = ((left < 64 ? left : 64) + 15) >> 4;
            if (a.resident) {
              tc_fence_after();
              if (leader) {
                uint64_t da = desc_hi + (uint64_t)(pa >> 4);
                for (int i = 0; i < nsub; ++i, da += 8, db_run += db_step)
                  for (int ks = 0; ks < ksteps; ++ks, accum = 1u) mma_ss(dcol, da + 2 * ks, db_run + 2 * ks, idesc, accum);
                mma_commit(&a_empty[sa]);
              } else {
                db_run += db_step * nsub;
              }
              __syncwarp();
              if (++sa == a.n_a) { sa = 0; pha ^= 1; }
              continue;
            }
            for (int i0 = 0; i0 < nsub; i0 += a.b_group) {
              const int gn = nsub - i0 < a.b_group ? nsub - i0 : a.b_group;
              mbar_wait(&b_full[sb], phb);
              tc_fence_after();
              const uint32_t pb = smem_u32(sB + sb * a.b_stage);
              if (leader) {
                for (int i = 0; i < gn; ++i) {
                  const uint64_t da = desc_hi + (uint64_t)((pa + (i0 + i) * 128) >> 4);
                  const uint64_t db = desc_hi + (uint64_t)((pb + i * b_bytes) >> 4);
                  if (!a.split) {
                    for (int ks = 0; ks < ksteps; ++ks, accum = 1u) mma_ss(dcol, da + 2 * ks, db + 2 * ks, idesc, accum);
                  } else {                                 // (hi + lo)(hi + lo) without the lo*lo term
                    const uint64_t da_lo = da + (uint64_t)(a_half >> 4), db_lo = db + (uint64_t)(b_half >> 4);
                    for (int ks = 0; ks < ksteps; ++ks, accum = 1u) {
                      mma_ss(dcol, da + 2 * ks, db + 2 * ks, idesc, accum);
                      mma_ss(dcol, da + 2 * ks, db_lo + 2 * ks, idesc, 1u);
                      mma_ss(dcol, da_lo + 2 * ks, db + 2 * ks, idesc, 1u);
                    }
                  }
                }
                mma_commit(&b_empty[sb]);
              }
              __syncwarp();
              if (++sb == a.n_b) { sb = 0; phb ^= 1; }
            }
            if (leader) mma_commit(&a_empty[sa]);
            __syncwarp();
            if (++sa == a.n_a) { sa = 0; pha ^= 1; }
          }
        }
      if (leader) mma_commit(&d_full[acc]);
      __syncwarp();
      if (++acc == a.n_acc) { acc = 0; phd ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: two sets of 4 warps, alternate tiles
    const int q = warp & 3, set = (warp - 2) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t oplane = (int64_t)a.OH * a.OW;
    uint32_t tile_i = 0;
    int acc = -1;
    uint32_t phd = 1;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tile_i) {
      if (++acc == a.n_acc) acc = 0;
      if (acc == 0) phd ^= 1;
      if ((int)(tile_i & 1) != set) continue;
      const int nt = t % a.n_tiles_n;
      int r = t / a.n_tiles_n;
      const int xb = r % a.x_tiles; r /= a.x_tiles;
      const int oy = r % a.OH, n = r / a.OH;
      const int x = xb * 128 + q * 32 + lane;
      const bool live = x < a.OW;
      const int64_t sp = (int64_t)oy * a.OW + x;
      const float* resp = a.res ? a.res + ((int64_t)n * d.res_c_total + d.res_c_off) * oplane + sp : nullptr;
      float* yp = a.y ? a.y + ((int64_t)n * d.out_c_total + d.out_c_off) * oplane + sp : nullptr;
      __half* ypix = (a.yp && live) ? a.yp + ((((int64_t)n * a.yp_hp2 + oy + a.yp_pad_h) * a.yp_wp2 + x + a.yp_pad_w) * a.yp_cp +
                                               a.yp_c_off) : nullptr;
      mbar_wait(&d_full[acc], phd);
      tc_fence_after();
      for (int c0 = 0; c0 < a.n_tile; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(tlane + acc * a.acc_stride + c0, raw);
        tmem_ld_wait();
        const int jb = nt * a.n_tile + c0;
        int nvalid = a.n_tile - c0;
        if (nvalid > d.Cout - jb) nvalid = d.Cout - jb;
        if (nvalid > 32) nvalid = 32;
        const float* bq = a.bias ? a.bias + jb : nullptr;
        const float* rq = resp ? resp + (int64_t)jb * oplane : nullptr;
        float* yq = yp ? yp + (int64_t)jb * oplane : nullptr;
        __half* yk = ypix ? ypix + jb : nullptr;
        switch (d.act) {
          case 1: tm_store32<1>(raw, bq, d.out_scale, rq, yq, oplane, nvalid, live, lane, yk); break;
          case 2: tm_store32<2>(raw, bq, d.out_scale, rq, yq, oplane, nvalid, live, lane, yk); break;
          case 3: tm_store32<3>(raw, bq, d.out_scale, rq, yq, oplane, nvalid, live, lane, yk); break;
          case 4: tm_store32<4>(raw, bq, d.out_scale, rq, yq, oplane, nvalid, live, lane, yk); break;
          default: tm_store32<0>(raw, bq, d.out_scale, rq, yq, oplane, nvalid, live, lane, yk); break;
        }
      }
      tc_fence_before();
      mbar_arrive(&d_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

__device__ __forceinline__ int tm_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// fp32 NCHW slice -> fp16 [N*phases][HP2][WP2][Cp] (Cp = channels padded to 64) with padding / upsampling /
// phase split applied.  One block = 32 pixels of one row x 64 channels, transposed through shared memory.
__global__ void __launch_bounds__(256) conv_pack_input_kernel(const float* __restrict__ x, __half* __restrict__ xp,
                                                              B200ConvDesc d, int HP2, int WP2, int Cp, int phases,
                                                              float scale, int64_t lo_off, int fold_cf) {
  __shared__ float tile[64][33];
  const int cblocks = Cp / 64;
  const int cb = blockIdx.z % cblocks, nph = blockIdx.z / cblocks;
  const int ph = nph % phases, n = nph / phases;
  const int yy = blockIdx.y, bx = blockIdx.x * 32;
  const int s = d.stride, HU = d.H * d.upsample, WU = d.W * d.upsample;
  {
    const int lx_ = threadIdx.x & 31, crow = threadIdx.x >> 5;
    if (fold_cf) {
      // channel slot k = kx * fold_cf + c holds input channel c at x + kx (stride 1, nearest sampling only)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = crow + 8 * i, kx = k / fold_cf, c = k - kx * fold_cf;
        const int Y = yy - d.pad_h, X = bx + lx_ + kx - d.pad_w;
        int u = Y, w = X;
        bool ok = kx < d.KW && c < d.Cin && bx + lx_ < WP2;
        if (d.pad_mode == 1) { u = tm_reflect(Y, HU); w = tm_reflect(X, WU); ok = ok && Y >= -d.pad_h && Y < HU + d.pad_h; }
        else ok = ok && Y >= 0 && Y < HU && X >= 0 && X < WU;
        if (d.upsample > 1) { u >>= 1; w >>= 1; }
        tile[k][lx_] = ok ? __ldg(x + ((int64_t)n * d.in_c_total + d.in_c_off + c) * d.H * d.W + (int64_t)u * d.W + w) : 0.f;
      }
    } else {
    const int Y = yy * s + (phases == 4 ? (ph >> 1) : 0) - d.pad_h;
    const int X = (bx + lx_) * s + (phases == 4 ? (ph & 1) : 0) - d.pad_w;
    int u = Y, w = X;
    bool ok;
    if (d.pad_mode == 1) {
      ok = Y >= -d.pad_h && Y < HU + d.pad_h && X >= -d.pad_w && X < WU + d.pad_w && bx + lx_ < WP2;
      u = tm_reflect(Y, HU); w = tm_reflect(X, WU);
    } else {
      ok = Y >= 0 && Y < HU && X >= 0 && X < WU;
    }
    const int64_t plane = (int64_t)d.H * d.W;
    const float* src0 = x + ((int64_t)n * d.in_c_total + d.in_c_off) * plane;
    if (d.upsample > 1 && d.upsample_mode == 1) {
      // bilinear x2, align_corners=True (ATen area_pixel_compute_source_index): src = dst * (in-1)/(out-1)
      const float sh = HU > 1 ? (float)(d.H - 1) / (float)(HU - 1) : 0.f;
      const float sw = WU > 1 ? (float)(d.W - 1) / (float)(WU - 1) : 0.f;
      const float fy = sh * u, fx = sw * w;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < d.H - 1 ? 1 : 0), x1 = x0 + (x0 < d.W - 1 ? 1 : 0);
      const float ly = fy - y0, lx = fx - x0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cb * 64 + crow + 8 * i;
        float v = 0.f;
        if (ok && c < d.Cin) {
          const float* s = src0 + (int64_t)c * plane;
          v = (1.f - ly) * ((1.f - lx) * __ldg(s + y0 * d.W + x0) + lx * __ldg(s + y0 * d.W + x1)) +
              ly * ((1.f - lx) * __ldg(s + y1 * d.W + x0) + lx * __ldg(s + y1 * d.W + x1));
        }
        tile[crow + 8 * i][lx_] = v;
      }
    } else {
      if (d.upsample > 1) { u >>= 1; w >>= 1; }
      const float* src = src0 + (int64_t)u * d.W + w;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cb * 64 + crow + 8 * i;
        tile[crow + 8 * i][lx_] = (ok && c < d.Cin) ? __ldg(src + (int64_t)c * plane) : 0.f;
      }
    }
    }
  }
  __syncthreads();
  const int px = threadIdx.x >> 3, c8 = threadIdx.x & 7;
  if (bx + px < WP2) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = tile[c8 * 8 + q][px] * scale;
    uint4 hi, lo;
    split2_f16(v[0], v[1], hi.x, lo.x); split2_f16(v[2], v[3], hi.y, lo.y);
    split2_f16(v[4], v[5], hi.z, lo.z); split2_f16(v[6], v[7], hi.w, lo.w);
    const int64_t o = (((int64_t)nph * HP2 + yy) * WP2 + bx + px) * Cp + cb * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(xp + o) = hi;
    if (lo_off) *reinterpret_cast<uint4*>(xp + lo_off + o) = lo;
  }
}

// nn.ReflectionPad2d on a packed input whose interior was written by the producing convolutions: every halo pixel of
// [N][HP2][WP2][Cp] copies the vector of its mirror pixel.  One thread per (halo pixel, 8-channel group).
__global__ void conv_reflect_halo_kernel(__half* __restrict__ xp, int N, int H, int W, int pad_h, int pad_w, int Cp) {
  const int HP2 = H + 2 * pad_h, WP2 = W + 2 * pad_w;
  const int c8n = Cp / 8;
  const int64_t halo_px = (int64_t)HP2 * WP2 - (int64_t)H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * halo_px * c8n) return;
  const int c8 = (int)(i % c8n);
  int64_t r = i / c8n;
  const int n = (int)(r / halo_px);
  r -= (int64_t)n * halo_px;
  // halo pixels in order: pad_h full rows on top, pad_h full rows at the bottom, then the side columns of the H rows
  int Y, X;
  const int64_t band = (int64_t)pad_h * WP2;
  if (r < band) { Y = (int)(r / WP2); X = (int)(r % WP2); }
  else if (r < 2 * band) { r -= band; Y = pad_h + H + (int)(r / WP2); X = (int)(r % WP2); }
  else {
    r -= 2 * band;
    Y = pad_h + (int)(r / (2 * pad_w));
    const int k = (int)(r % (2 * pad_w));
    X = k < pad_w ? k : W + k;                     // left columns 0..pad_w-1, right columns pad_w+W ..
  }
  const int sy = tm_reflect(Y - pad_h, H) + pad_h, sx = tm_reflect(X - pad_w, W) + pad_w;
  const uint4 v = *reinterpret_cast<const uint4*>(xp + (((int64_t)n * HP2 + sy) * WP2 + sx) * Cp + c8 * 8);
  *reinterpret_cast<uint4*>(xp + (((int64_t)n * HP2 + Y) * WP2 + X) * Cp + c8 * 8) = v;
}

// weights [Cout][Cin][KH][KW] fp32 -> per cout tile, per (tap, channel block): [n_tile rows x 64 channels] fp16,
// K-major, 128-byte swizzle, zero padded
// (w_co, w_ci, w_tap = element strides of the weight tensor; split: chunk = [hi rows | lo rows], values pre-scaled)
__global__ void conv_tma_weight_images_kernel(const float* __restrict__ w, char* __restrict__ img, int Cout, int Cin,
                                              int KH, int KW, int stride, int n_tile, int n_tiles_n, int cchunks,
                                              int64_t w_co, int64_t w_ci, int64_t w_tap, float scale, int split,
                                              int fold_cf) {
  const int KHW = KH * KW, n_chunks = fold_cf ? KH : KHW * cchunks;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk each
  if (e >= (int64_t)n_tiles_n * n_tile * n_chunks * 8) return;
  const int c8 = (int)(e % 8), j = (int)((e / 8) % n_chunks), grow = (int)(e / (8 * n_chunks));
  const int nt = grow / n_tile, lrow = grow % n_tile, row = nt * n_tile + lrow;
  // chunk j -> (ky, xpar, cc, kx) in the kernel's reduction order
  int tap = 0, cc = 0;
  if (!fold_cf) {
    int cnt = 0;
    bool found = false;
    for (int ky = 0; ky < KH && !found; ++ky)
      for (int xpar = 0; xpar < stride && xpar < KW && !found; ++xpar)
        for (int c = 0; c < cchunks && !found; ++c)
          for (int kx = xpar; kx < KW; kx += stride, ++cnt)
            if (cnt == j) { tap = ky * KW + kx; cc = c; found = true; break; }
  }
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    int ci = cc * 64 + c8 * 8 + q, tp = tap;
    bool ok = row < Cout;
    if (fold_cf) {                                         // chunk j = filter row, slot k = kx * fold_cf + c
      const int kx = ci / fold_cf;
      ci -= kx * fold_cf; tp = j * KW + kx; ok = ok && kx < KW;
    }
    v[q] = (ok && ci < Cin) ? w[row * w_co + ci * w_ci + tp * w_tap] * scale : 0.f;
  }
  uint4 hi, lo;
  split2_f16(v[0], v[1], hi.x, lo.x); split2_f16(v[2], v[3], hi.y, lo.y);
  split2_f16(v[4], v[5], hi.z, lo.z); split2_f16(v[6], v[7], hi.w, lo.w);
  const int rr = lrow & 7;
  const int64_t chunk_bytes = (int64_t)n_tile * 128 * (1 + split);
  char* dst = img + ((int64_t)nt * n_chunks + j) * chunk_bytes + (lrow >> 3) * 1024 + rr * 128 + ((c8 ^ rr) << 4);
  *reinterpret_cast<uint4*>(dst) = hi;
  if (split) *reinterpret_cast<uint4*>(dst + (int64_t)n_tile * 128) = lo;
}

struct TmaGeom {
  int HU, WU, OH, OW, phases, shift, HP2, WP2, Cp, cchunks, n_chunks, n_tile, n_tiles_n;
  int fold_cf;      // > 0: the KW x taps are folded into the channel dimension, fold_cf channels per tap (narrow inputs)
  int64_t pack_bytes;
};

static int tma_geometry(const B200ConvDesc* d, TmaGeom* g) {
  B200_REQUIRE(d, "null descriptor");
  B200_REQUIRE(d->N > 0 && d->Cin > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
               (d->stride == 1 || d->stride == 2) && (d->upsample == 1 || d->upsample == 2) &&
               (d->pad_mode == 0 || d->pad_mode == 1) && d->act >= 0 && d->act <= 4 && d->pad_h >= 0 && d->pad_w >= 0 &&
               (d->upsample_mode == 0 || (d->upsample_mode == 1 && d->upsample == 2)),
               "invalid convolution descriptor (b200_conv2d_tma supports stride 1 and 2)");
  g->HU = d->H * d->upsample; g->WU = d->W * d->upsample;
  B200_REQUIRE(d->pad_mode == 0 || (d->pad_h < g->HU && d->pad_w < g->WU), "reflection padding larger than the input");
  const int HP = g->HU + 2 * d->pad_h, WP = g->WU + 2 * d->pad_w;
  B200_REQUIRE(HP >= d->KH && WP >= d->KW, "empty output");
  g->OH = (HP - d->KH) / d->stride + 1; g->OW = (WP - d->KW) / d->stride + 1;
  g->phases = d->stride == 2 ? 4 : 1; g->shift = d->stride == 2 ? 1 : 0;
  g->HP2 = (HP + d->stride - 1) / d->stride;
  g->WP2 = (WP + d->stride - 1) / d->stride;
  g->cchunks = (d->Cin + 63) / 64;
  g->n_chunks = d->KH * d->KW * g->cchunks;
  g->Cp = g->cchunks * 64;
  // narrow inputs (Cin <= 8 for 7-wide filters, <= 16 for 3-wide ...): pack the KW horizontally shifted copies of a
  // pixel next to each other in its 64-channel vector; the convolution then has KW = 1 and one chunk per filter row
  g->fold_cf = 0;
  const int cf = (d->Cin + 7) / 8 * 8;
  if (d->stride == 1 && d->KW > 1 && cf * d->KW <= 64 && d->upsample_mode == 0) {
    g->fold_cf = cf;
    g->WP2 = g->OW;
    g->n_chunks = d->KH;
  }
  g->n_tiles_n = (d->Cout + 255) / 256;
  g->n_tile = ((d->Cout + g->n_tiles_n - 1) / g->n_tiles_n + 15) / 16 * 16;
  g->pack_bytes = (int64_t)d->N * g->phases * g->HP2 * g->WP2 * g->Cp * 2;
  return B200_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int launch_weight_images(const B200ConvDesc* d, const TmaGeom& g, const float* w, int64_t w_co, int64_t w_ci,
                                int64_t w_tap, float scale, int split, void* images, cudaStream_t st) {
  const int64_t total = (int64_t)g.n_tiles_n * g.n_tile * g.n_chunks * 8;
  conv_tma_weight_images_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      w, reinterpret_cast<char*>(images), d->Cout, d->Cin, d->KH, d->KW, d->stride, g.n_tile, g.n_tiles_n, g.cchunks, w_co,
      w_ci, w_tap, scale, split, g.fold_cf);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// repack x into `base` (fp16, hi then lo when split) and run the convolution
struct ChainOut { __half* yp = nullptr; int hp2 = 0, wp2 = 0, cp = 0, pad_h = 0, pad_w = 0, c_off = 0; };

// x == nullptr: `base` already holds the packed input (written by the producing convolution's epilogue)
static int launch_conv_tma(const B200ConvDesc* d, const TmaGeom& g, const float* x, char* base, const void* w_images,
                           const float* bias, const float* residual, float* y, int split, float in_scale, cudaStream_t st,
                           const ChainOut* chain = nullptr) {
  B200_REQUIRE(g.HP2 <= 65535 && (int64_t)d->N * g.phases * g.cchunks <= 65535, "input too large for the repack grid");
  if (x) {
    conv_pack_input_kernel<<<dim3((g.WP2 + 31) / 32, g.HP2, d->N * g.phases * g.cchunks), 256, 0, st>>>(
        x, reinterpret_cast<__half*>(base), *d, g.HP2, g.WP2, g.Cp, g.phases, in_scale, split ? g.pack_bytes / 2 : 0, g.fold_cf);
    B200_CHECK_LAUNCH();
  }

  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200_ERR_UNSUPPORTED; }
  alignas(64) CUtensorMap map, map_lo;
  const cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.WP2, (cuuint64_t)g.HP2, (cuuint64_t)d->N * g.phases};
  const cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 2, (cuuint64_t)g.WP2 * g.Cp * 2, (cuuint64_t)g.HP2 * g.WP2 * g.Cp * 2};
  const int kw_eff = g.fold_cf ? 1 : d->KW;               // folded: the x taps live in the channel dimension
  const int a_rows = 128 + ((kw_eff - 1) >> g.shift);
  B200_REQUIRE(a_rows <= 256, "filter too wide");
  const cuuint32_t box[4] = {64, (cuuint32_t)a_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int term = 0; term <= split; ++term) {
    const CUresult cr = enc(term ? &map_lo : &map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base + term * g.pack_bytes, dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return B200_ERR_CUDA; }
  }
  if (!split) map_lo = map;

  static bool attr_done = false;
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(conv2d_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BUDGET));
    attr_done = true;
  }
  ConvTmaArgs a{};
  a.d = *d; a.w_img = reinterpret_cast<const char*>(w_images); a.bias = bias; a.res = residual; a.y = y;
  if (chain && chain->yp) {
    a.yp = chain->yp; a.yp_hp2 = chain->hp2; a.yp_wp2 = chain->wp2; a.yp_cp = chain->cp; a.yp_pad_h = chain->pad_h;
    a.yp_pad_w = chain->pad_w; a.yp_c_off = chain->c_off;
  }
  if (g.fold_cf) { a.d.KW = 1; a.d.Cin = g.fold_cf * d->KW; }
  a.OH = g.OH; a.OW = g.OW; a.x_tiles = (g.OW + 127) / 128; a.cchunks = g.cchunks; a.n_chunks = g.n_chunks;
  a.n_tile = g.n_tile; a.n_tiles_n = g.n_tiles_n; a.phases = g.phases; a.shift = g.shift;
  a.a_rows = a_rows; a.split = split;
  a.a_stage = (a_rows * 128 + 1023) / 1024 * 1024 * (1 + split);
  const int b_bytes = g.n_tile * 128 * (1 + split);
  // x taps that share one activation box are fetched as one bulk copy while that stays <= 32 KB
  const int nsub_max = (kw_eff + d->stride - 1) / d->stride;
  int b_group = 32768 / b_bytes; if (b_group < 1) b_group = 1; if (b_group > nsub_max) b_group = nsub_max;
  a.b_group = b_group; a.b_stage = b_group * b_bytes;
  // shared memory: at least 3 B stages (2 for the widest tiles), up to 6 A stages, the rest goes to the B ring;
  // narrow layers keep the whole weight image resident instead (no per-chunk weight traffic or barriers)
  const int avail = TM_SMEM_BUDGET - TM_BAR_BYTES;
  a.resident = (!split && g.n_tiles_n == 1 && (int64_t)g.n_chunks * b_bytes <= 100 * 1024) ? 1 : 0;
  int n_a, n_b;
  if (a.resident) {
    a.b_group = 1; a.b_stage = b_bytes;
    n_b = g.n_chunks;                                      // "ring" = the resident image
    n_a = (avail - n_b * a.b_stage) / a.a_stage;
    if (n_a > TM_MAX_A) n_a = TM_MAX_A;
  } else {
    n_a = (avail - 3 * a.b_stage) / a.a_stage;
    if (n_a > 6) n_a = 6;
    if (n_a < 2) n_a = 2;
    n_b = (avail - n_a * a.a_stage) / a.b_stage;
    if (n_b > TM_MAX_B) n_b = TM_MAX_B;
  }
  B200_REQUIRE(n_a >= 2 && (a.resident || n_b >= 2), "tile does not fit in shared memory");
  a.n_a = n_a; a.n_b = n_b;
  const int cols = (g.n_tile + 31) / 32 * 32;
  a.n_acc = 512 / cols; if (a.n_acc > TM_MAX_ACC) a.n_acc = TM_MAX_ACC; a.n_acc &= ~1;
  a.acc_stride = 512 / a.n_acc;
  int smem_bytes = n_a * a.a_stage + n_b * a.b_stage + TM_BAR_BYTES;
  // each CTA allocates all 512 TMEM columns: never let two CTAs share an SM (a second tcgen05.alloc would wait forever)
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  const int64_t tiles = (int64_t)d->N * g.OH * a.x_tiles * g.n_tiles_n;
  B200_REQUIRE(tiles < (1ll << 31), "too many tiles");
  a.total_tiles = (int)tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  conv2d_tma_kernel<<<(unsigned)(tiles < sms ? tiles : sms), TM_THREADS, smem_bytes, st>>>(a, map, map_lo);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // namespace b200

namespace b200 { int launch_avgpool2(const float* x, float* y, int64_t planes, int H, int W, cudaStream_t st); }   // raft_kernels.cu
using namespace b200;

extern "C" {

int64_t b200_conv_tma_workspace_bytes(const B200ConvDesc* d) {
  TmaGeom g;
  if (tma_geometry(d, &g) != B200_OK) return -1;
  return g.pack_bytes + 256;
}

int64_t b200_conv_tma_weight_image_bytes(const B200ConvDesc* d) {
  TmaGeom g;
  if (tma_geometry(d, &g) != B200_OK) return -1;
  return (int64_t)g.n_tiles_n * g.n_chunks * g.n_tile * 128;
}

int b200_conv_tma_weight_images(const B200ConvDesc* d, const float* w, void* images, void* stream) {
  B200_REQUIRE(w && images, "null pointer");
  TmaGeom g;
  if (int rc = tma_geometry(d, &g)) return rc;
  const int KHW = d->KH * d->KW;
  return launch_weight_images(d, g, w, (int64_t)d->Cin * KHW, KHW, 1, 1.0f, 0, images, reinterpret_cast<cudaStream_t>(stream));
}

int b200_conv2d_tma(const B200ConvDesc* d, const float* x, const void* w_images, const float* bias,
                    const float* residual, float* y, void* workspace, int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(x && w_images && y && workspace, "null pointer");
  TmaGeom g;
  if (int rc = tma_geometry(d, &g)) return rc;
  B200_REQUIRE(d->in_c_off >= 0 && d->in_c_off + d->Cin <= d->in_c_total && d->out_c_off >= 0 &&
               d->out_c_off + d->Cout <= d->out_c_total, "channel slice out of range");
  if (!b200_device_supports_tc()) { set_error("b200_conv2d_tma needs a compute-capability 10.x device"); return B200_ERR_UNSUPPORTED; }
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  if (base + g.pack_bytes > reinterpret_cast<char*>(workspace) + workspace_bytes) {
    set_error("b200_conv2d_tma: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)(g.pack_bytes + 256));
    return B200_ERR_WORKSPACE;
  }
  return launch_conv_tma(d, g, x, base, w_images, bias, residual, y, 0, 1.0f, reinterpret_cast<cudaStream_t>(stream));
}

/* A convolution can consume a packed input written by its producers when its packing is the plain one: stride 1, no
 * upsampling, x taps not folded into the channel vector.  Zero padding: the halo of the buffer stays zero; reflection
 * padding: the consumer call first mirrors the interior into the halo (conv_reflect_halo_kernel). */
int b200_conv_tma_chainable(const B200ConvDesc* next) {
  TmaGeom g;
  if (!next || tma_geometry(next, &g) != B200_OK) return 0;
  return (next->stride == 1 && next->upsample == 1 && g.fold_cf == 0 && next->in_c_off == 0 &&
          next->in_c_total == next->Cin) ? 1 : 0;
}

int b200_conv2d_tma_chain(const B200ConvDesc* d, const float* x, void* in_packed, const void* w_images, const float* bias,
                          const float* residual, float* y, void* out_packed, const B200ConvDesc* next, int32_t next_c_off,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(d && w_images && (x || in_packed) && (y || out_packed), "null pointer");
  TmaGeom g;
  if (int rc = tma_geometry(d, &g)) return rc;
  B200_REQUIRE(d->in_c_off >= 0 && d->in_c_off + d->Cin <= d->in_c_total && d->out_c_off >= 0 &&
               d->out_c_off + d->Cout <= d->out_c_total, "channel slice out of range");
  if (!b200_device_supports_tc()) { set_error("b200_conv2d_tma_chain needs a compute-capability 10.x device"); return B200_ERR_UNSUPPORTED; }
  char* base = nullptr;
  if (in_packed) {
    B200_REQUIRE(b200_conv_tma_chainable(d), "this convolution cannot take a pre-packed input");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(in_packed) & 255) == 0, "packed input must be 256-byte aligned");
    base = reinterpret_cast<char*>(in_packed);
  } else {
    B200_REQUIRE(workspace, "null workspace");
    base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    if (base + g.pack_bytes > reinterpret_cast<char*>(workspace) + workspace_bytes) {
      set_error("b200_conv2d_tma_chain: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)(g.pack_bytes + 256));
      return B200_ERR_WORKSPACE;
    }
  }
  ChainOut co;
  if (out_packed) {
    B200_REQUIRE(next && b200_conv_tma_chainable(next), "the consumer convolution cannot take a pre-packed input");
    TmaGeom gn;
    if (int rc = tma_geometry(next, &gn)) return rc;
    B200_REQUIRE(next->N == d->N && next->H == g.OH && next->W == g.OW, "consumer geometry does not match this output");
    B200_REQUIRE(next_c_off >= 0 && next_c_off + d->Cout <= next->Cin && (next_c_off & 7) == 0 && (d->Cout & 7) == 0,
                 "chained channel slice must lie inside the consumer's input and be a multiple of 8 channels");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(out_packed) & 255) == 0, "packed output must be 256-byte aligned");
    co.yp = reinterpret_cast<__half*>(out_packed); co.hp2 = gn.HP2; co.wp2 = gn.WP2; co.cp = gn.Cp;
    co.pad_h = next->pad_h; co.pad_w = next->pad_w; co.c_off = next_c_off;
  }
  if (in_packed && d->pad_mode == 1 && (d->pad_h > 0 || d->pad_w > 0)) {
    const int64_t threads = (int64_t)d->N * ((int64_t)g.HP2 * g.WP2 - (int64_t)d->H * d->W) * (g.Cp / 8);
    conv_reflect_halo_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__half*>(base), d->N, d->H, d->W, d->pad_h, d->pad_w, g.Cp);
    B200_CHECK_LAUNCH();
  }
  return launch_conv_tma(d, g, in_packed ? nullptr : x, base, w_images, bias, residual, y, 0, 1.0f,
                         reinterpret_cast<cudaStream_t>(stream), &co);
}

/* ---- all-pairs correlation (RAFT CorrBlock level 0, src/models/stage_1/core/corr.py:56-64) as a 1x1 "convolution":
 * input = fmap2 (pixels p2), weights[co = p1][ci] = fmap1[ci][p1], output [p1][p2].  The reference computes it in
 * fp32 (fmaps are cast with .float() and TF32 matmul is off by default), so both operands are split into
 * (hi, lo) fp16 pairs, pre-scaled by 2^4: 3 tensor-core products per element, fp32 accumulation — the same
 * fp32-grade emulation as the stage-1 MLPs. */
static void corr_desc(int dim, int H8, int W8, B200ConvDesc* d) {
  *d = B200ConvDesc{};
  d->N = 1; d->Cin = dim; d->H = H8; d->W = W8; d->in_c_total = dim; d->in_c_off = 0;
  d->Cout = H8 * W8; d->KH = 1; d->KW = 1; d->stride = 1; d->pad_h = 0; d->pad_w = 0; d->pad_mode = 0; d->upsample = 1;
  d->out_c_total = H8 * W8; d->out_c_off = 0; d->act = 0;
  d->out_scale = 1.0f / sqrtf((float)dim) / 256.0f;       // undo the two 2^4 operand scales
  d->res_c_total = 0; d->res_c_off = 0;
}

// floats of the three pooled copies of fmap2 (levels 1..3 of the pyramid are GEMMs on them, see below)
static int64_t corr_pooled_floats(int dim, int H8, int W8) {
  int64_t n = 0;
  int h = H8, w = W8;
  for (int l = 1; l < 4; ++l) { h /= 2; w /= 2; n += (int64_t)dim * h * w + 64; }
  return n;
}

int64_t b200_corr_build_tc_workspace_bytes(int32_t dim, int32_t H8, int32_t W8) {
  if (dim <= 0 || H8 < 8 || W8 < 8) return -1;
  B200ConvDesc d; corr_desc(dim, H8, W8, &d);
  TmaGeom g;
  if (tma_geometry(&d, &g) != B200_OK) return -1;
  return 2 * g.pack_bytes + 2 * (int64_t)g.n_tiles_n * g.n_chunks * g.n_tile * 128 + corr_pooled_floats(dim, H8, W8) * 4 + 2048;
}

int b200_corr_build_tc(const float* fmap1, const float* fmap2, int32_t dim, int32_t H8, int32_t W8, float* pyramid,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(fmap1 && fmap2 && pyramid && workspace && dim > 0 && H8 >= 8 && W8 >= 8, "bad arguments");
  if (!b200_device_supports_tc()) { set_error("b200_corr_build_tc needs a compute-capability 10.x device"); return B200_ERR_UNSUPPORTED; }
  B200ConvDesc d; corr_desc(dim, H8, W8, &d);
  TmaGeom g;
  if (int rc = tma_geometry(&d, &g)) return rc;
  const int64_t need = b200_corr_build_tc_workspace_bytes(dim, H8, W8);
  B200_REQUIRE(workspace_bytes >= need, "b200_corr_build_tc: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  char* images = base + 2 * g.pack_bytes;                  // pack_bytes is a multiple of 128
  const int HW = H8 * W8;
  if (int rc = launch_weight_images(&d, g, fmap1, /*co*/ 1, /*ci*/ HW, /*tap*/ 0, 16.0f, 1, images, st)) return rc;
  if (int rc = launch_conv_tma(&d, g, fmap2, base, images, nullptr, nullptr, pyramid, 1, 16.0f, st)) return rc;
  // Levels 1..3 (corr.py:27-31: avg_pool2d of the previous level over the TARGET pixel grid).  Average pooling is linear
  // in fmap2, so level l = fmap1^T . avgpool^l(fmap2): three small GEMMs on pooled 33 MB feature maps with the same
  // fmap1 weight images, instead of three passes that re-read the 4.2 GB level-0 volume (1.7 of 3.5 ms at 1080p).
  const int64_t image_bytes = 2 * (int64_t)g.n_tiles_n * g.n_chunks * g.n_tile * 128;
  float* pooled = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(images + image_bytes) + 255) & ~(uintptr_t)255);
  const float* src = fmap2;
  float* out = pyramid + (int64_t)HW * HW;
  int h = H8, w = W8;
  for (int l = 1; l < 4; ++l) {
    if (int rc = launch_avgpool2(src, pooled, dim, h, w, st)) return rc;
    h /= 2; w /= 2;
    if (h < 1 || w < 1) break;
    B200ConvDesc dl = d;
    dl.H = h; dl.W = w;                                   // queries (output channels) stay at full resolution
    TmaGeom gl;
    if (int rc = tma_geometry(&dl, &gl)) return rc;
    if (int rc = launch_conv_tma(&dl, gl, pooled, base, images, nullptr, nullptr, out, 1, 16.0f, st)) return rc;
    out += (int64_t)HW * h * w;
    src = pooled;
    pooled += (((int64_t)dim * h * w + 63) / 64) * 64;
  }
  return B200_OK;
}

}  // extern "C"
