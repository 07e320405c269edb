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

constexpr int TM_STAGES = 4;
constexpr int TM_A_BYTES = 16384;               // 128 pixels x 64 channels x 2 B
constexpr int TM_B_BYTES = 32768;               // <= 256 rows x 128 B
constexpr int TM_STAGE_BYTES = TM_A_BYTES + TM_B_BYTES;
constexpr int TM_THREADS = 320;
constexpr int TM_SMEM = TM_STAGES * TM_STAGE_BYTES + 256;

struct ConvTmaArgs {
  B200ConvDesc d;
  const char* w_img; const float* bias; const float* res; float* y;
  int OH, OW, x_tiles, cchunks, n_chunks, n_tile, n_tiles_n, phases, shift, total_tiles;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float tm_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.0f / (1.0f + expf(-v));
    case 4: return tanhf(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(TM_THREADS, 1) conv2d_tma_kernel(const __grid_constant__ ConvTmaArgs a,
                                                                   const __grid_constant__ CUtensorMap xmap) {
  extern __shared__ __align__(1024) char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TM_STAGES * TM_STAGE_BYTES);
  uint64_t* full = bars;                     // [4] tx
  uint64_t* empty = full + TM_STAGES;        // [4] commit
  uint64_t* d_full = empty + TM_STAGES;      // [2] commit
  uint64_t* d_empty = d_full + 2;            // [2] 256 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const B200ConvDesc& d = a.d;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("b200: conv smem not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < TM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 256); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int b_bytes = a.n_tile * 128;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        const int nt = t % a.n_tiles_n;
        int r = t / a.n_tiles_n;
        const int xb = r % a.x_tiles; r /= a.x_tiles;
        const int oy = r % a.OH, n = r / a.OH;
        const char* wsrc = a.w_img + (int64_t)nt * a.n_chunks * b_bytes;
        int tap = 0, cc = 0, ky = 0, kx = 0;
        for (int j = 0; j < a.n_chunks; ++j, ++it) {
          const int s = it % TM_STAGES;
          mbar_wait(&empty[s], ((it / TM_STAGES) & 1) ^ 1);
          char* sa = smem + s * TM_STAGE_BYTES;
          mbar_expect_tx(&full[s], TM_A_BYTES + b_bytes);
          const int ph = a.phases == 4 ? ((ky & 1) * 2 + (kx & 1)) : 0;
          const int yy = oy + (ky >> a.shift), xx = xb * 128 + (kx >> a.shift);
          tma_load_4d(sa, &xmap, cc * 64, xx, yy, n * a.phases + ph, &full[s]);
          bulk_g2s(sa + TM_A_BYTES, wsrc + (int64_t)j * b_bytes, b_bytes, &full[s]);
          if (++cc == a.cchunks) { cc = 0; ++tap; if (++kx == d.KW) { kx = 0; ++ky; } }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, a.n_tile, 0, 0);
      uint32_t it = 0, tile_i = 0;
      for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tile_i) {
        const int acc = tile_i & 1;
        mbar_wait(&d_empty[acc], ((tile_i >> 1) & 1) ^ 1);
        tc_fence_after();
        int cc = 0;
        for (int j = 0; j < a.n_chunks; ++j, ++it) {
          const int s = it % TM_STAGES;
          mbar_wait(&full[s], (it / TM_STAGES) & 1);
          tc_fence_after();
          const uint32_t pa = smem_u32(smem + s * TM_STAGE_BYTES), pb = pa + TM_A_BYTES;
          const int left = d.Cin - cc * 64;
          const int ksteps = ((left < 64 ? left : 64) + 15) >> 4;
          for (int ks = 0; ks < ksteps; ++ks)
            mma_ss(tmem + acc * 256, make_desc(pa + ks * 32, 16, 1024), make_desc(pb + ks * 32, 16, 1024), idesc,
                   (j | ks) ? 1u : 0u);
          mma_commit(&empty[s]);
          if (++cc == a.cchunks) cc = 0;
        }
        mma_commit(&d_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t oplane = (int64_t)a.OH * a.OW;
    uint32_t tile_i = 0;
    for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x, ++tile_i) {
      const int acc = tile_i & 1;
      const int nt = t % a.n_tiles_n;
      int r = t / a.n_tiles_n;
      const int xb = r % a.x_tiles; r /= a.x_tiles;
      const int oy = r % a.OH, n = r / a.OH;
      const int x = xb * 128 + q * 32 + lane;
      const bool live = x < a.OW;
      const int64_t sp = (int64_t)oy * a.OW + x;
      mbar_wait(&d_full[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      for (int c0 = half * 32; c0 < a.n_tile; c0 += 64) {
        uint32_t raw[32];
        tmem_ld32(tlane + acc * 256 + c0, raw);
        tmem_ld_wait();
        if (!live) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = nt * a.n_tile + c0 + i;
          if (c0 + i < a.n_tile && j < d.Cout) {
            float val = __uint_as_float(raw[i]);
            if (a.bias) val += __ldg(a.bias + j);
            val = tm_act(val, d.act) * d.out_scale;
            if (a.res) val += __ldg(a.res + ((int64_t)n * d.res_c_total + d.res_c_off + j) * oplane + sp);
            a.y[((int64_t)n * d.out_c_total + d.out_c_off + j) * oplane + sp] = val;
          }
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
                                                              B200ConvDesc d, int HP2, int WP2, int Cp, int phases) {
  __shared__ float tile[64][33];
  const int cblocks = Cp / 64;
  const int cb = blockIdx.z % cblocks, nph = blockIdx.z / cblocks;
  const int ph = nph % phases, n = nph / phases;
  const int yy = blockIdx.y, bx = blockIdx.x * 32;
  const int s = d.stride, HU = d.H * d.upsample, WU = d.W * d.upsample;
  {
    const int lx = threadIdx.x & 31, crow = threadIdx.x >> 5;
    const int Y = yy * s + (phases == 4 ? (ph >> 1) : 0) - d.pad_h;
    const int X = (bx + lx) * s + (phases == 4 ? (ph & 1) : 0) - d.pad_w;
    int u = Y, w = X;
    bool ok;
    if (d.pad_mode == 1) {
      ok = Y >= -d.pad_h && Y < HU + d.pad_h && X >= -d.pad_w && X < WU + d.pad_w && bx + lx < WP2;
      u = tm_reflect(Y, HU); w = tm_reflect(X, WU);
    } else {
      ok = Y >= 0 && Y < HU && X >= 0 && X < WU;
    }
    if (d.upsample > 1) { u >>= 1; w >>= 1; }
    const float* src = x + ((int64_t)n * d.in_c_total + d.in_c_off) * d.H * d.W + (int64_t)u * d.W + w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cb * 64 + crow + 8 * i;
      tile[crow + 8 * i][lx] = (ok && c < d.Cin) ? __ldg(src + (int64_t)c * d.H * d.W) : 0.f;
    }
  }
  __syncthreads();
  const int px = threadIdx.x >> 3, c8 = threadIdx.x & 7;
  if (bx + px < WP2) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = tile[c8 * 8 + q][px];
    const uint4 pk = make_uint4(cvt_pack_f16x2(v[0], v[1]), cvt_pack_f16x2(v[2], v[3]), cvt_pack_f16x2(v[4], v[5]),
                                cvt_pack_f16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(xp + (((int64_t)nph * HP2 + yy) * WP2 + bx + px) * Cp + cb * 64 + c8 * 8) = pk;
  }
}

// weights [Cout][Cin][KH][KW] fp32 -> per cout tile, per (tap, channel block): [n_tile rows x 64 channels] fp16,
// K-major, 128-byte swizzle, zero padded
__global__ void conv_tma_weight_images_kernel(const float* __restrict__ w, char* __restrict__ img, int Cout, int Cin,
                                              int KHW, int n_tile, int n_tiles_n, int cchunks) {
  const int n_chunks = KHW * cchunks;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk each
  if (e >= (int64_t)n_tiles_n * n_tile * n_chunks * 8) return;
  const int c8 = (int)(e % 8), j = (int)((e / 8) % n_chunks), grow = (int)(e / (8 * n_chunks));
  const int nt = grow / n_tile, lrow = grow % n_tile, row = nt * n_tile + lrow;
  const int tap = j / cchunks, cc = j % cchunks;
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int ci = cc * 64 + c8 * 8 + q;
    v[q] = (row < Cout && ci < Cin) ? w[((int64_t)row * Cin + ci) * KHW + tap] : 0.f;
  }
  const uint4 pk = make_uint4(cvt_pack_f16x2(v[0], v[1]), cvt_pack_f16x2(v[2], v[3]), cvt_pack_f16x2(v[4], v[5]),
                              cvt_pack_f16x2(v[6], v[7]));
  const int rr = lrow & 7;
  char* dst = img + ((int64_t)nt * n_chunks + j) * n_tile * 128 + (lrow >> 3) * 1024 + rr * 128 + ((c8 ^ rr) << 4);
  *reinterpret_cast<uint4*>(dst) = pk;
}

struct TmaGeom {
  int HU, WU, OH, OW, phases, shift, HP2, WP2, Cp, cchunks, n_chunks, n_tile, n_tiles_n;
  int64_t pack_bytes;
};

static int tma_geometry(const B200ConvDesc* d, TmaGeom* g) {
  B200_REQUIRE(d, "null descriptor");
  B200_REQUIRE(d->N > 0 && d->Cin > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
               (d->stride == 1 || d->stride == 2) && (d->upsample == 1 || d->upsample == 2) &&
               (d->pad_mode == 0 || d->pad_mode == 1) && d->act >= 0 && d->act <= 4 && d->pad_h >= 0 && d->pad_w >= 0,
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

}  // namespace b200

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
  const int64_t total = (int64_t)g.n_tiles_n * g.n_tile * g.n_chunks * 8;
  conv_tma_weight_images_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, reinterpret_cast<char*>(images), d->Cout, d->Cin, d->KH * d->KW, g.n_tile, g.n_tiles_n, g.cchunks);
  B200_CHECK_LAUNCH();
  return B200_OK;
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  B200_REQUIRE(g.HP2 <= 65535 && (int64_t)d->N * g.phases * g.cchunks <= 65535, "input too large for the repack grid");
  conv_pack_input_kernel<<<dim3((g.WP2 + 31) / 32, g.HP2, d->N * g.phases * g.cchunks), 256, 0, st>>>(
      x, reinterpret_cast<__half*>(base), *d, g.HP2, g.WP2, g.Cp, g.phases);
  B200_CHECK_LAUNCH();

  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B200_ERR_UNSUPPORTED; }
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.WP2, (cuuint64_t)g.HP2, (cuuint64_t)d->N * g.phases};
  const cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 2, (cuuint64_t)g.WP2 * g.Cp * 2, (cuuint64_t)g.HP2 * g.WP2 * g.Cp * 2};
  const cuuint32_t box[4] = {64, 128, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return B200_ERR_CUDA; }

  static bool attr_done = false;
  if (!attr_done) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(conv2d_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
    attr_done = true;
  }
  ConvTmaArgs a{};
  a.d = *d; a.w_img = reinterpret_cast<const char*>(w_images); a.bias = bias; a.res = residual; a.y = y;
  a.OH = g.OH; a.OW = g.OW; a.x_tiles = (g.OW + 127) / 128; a.cchunks = g.cchunks; a.n_chunks = g.n_chunks;
  a.n_tile = g.n_tile; a.n_tiles_n = g.n_tiles_n; a.phases = g.phases; a.shift = g.shift;
  const int64_t tiles = (int64_t)d->N * g.OH * a.x_tiles * g.n_tiles_n;
  B200_REQUIRE(tiles < (1ll << 31), "too many tiles");
  a.total_tiles = (int)tiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  conv2d_tma_kernel<<<(unsigned)(tiles < sms ? tiles : sms), TM_THREADS, TM_SMEM, st>>>(a, map);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
