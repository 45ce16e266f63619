// rd_conv_direct.cu — CUDA-core convolution kernels (fp32 accumulate, fp32 or bf16 storage) and the
// rd_conv2d_{fwd,dgrad,wgrad} entry points that choose between them and the tcgen05 implicit-GEMM
// kernels of rd_conv_tc.cu.  The direct kernels serve (a) the fp32 parity mode (1e-3 vs the oracle),
// (b) the few layers whose channel counts are not tensor-core shaped (7->32, 4->C, C->4, 16->7 ...).
#include <stdlib.h>
#include "rd_common.cuh"

struct ConvGeom {
  int N, H, W, Cin;      // "input" of this pass (dy for dgrad)
  int OH, OW, Cout;      // "output" of this pass (dx for dgrad)
  int KH, KW, stride, pad;
  int ipg;               // images per group
  int mode;              // 0 forward gather, 1 transposed (dgrad) gather
};

// source coordinate of tap (kh,kw) for output pixel (oy,ox); returns false if the tap does not contribute
__device__ __forceinline__ bool tap_src(const ConvGeom& g, int oy, int ox, int kh, int kw, int& iy, int& ix) {
  if (g.mode == 0) {
    iy = oy * g.stride - g.pad + kh;
    ix = ox * g.stride - g.pad + kw;
  } else {
    int ty = oy + g.pad - kh, tx = ox + g.pad - kw;
    if (ty < 0 || tx < 0) return false;
    if (g.stride > 1) {
      if ((ty % g.stride) | (tx % g.stride)) return false;
      ty /= g.stride; tx /= g.stride;
    }
    iy = ty; ix = tx;
  }
  return iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
}

constexpr int kDirPix = 64, kDirCo = 16, kDirCi = 32;

template <typename T>
__global__ void __launch_bounds__(128) k_conv_direct(const T* __restrict__ x, const T* __restrict__ w,
                                                      const float* __restrict__ bias, T* __restrict__ y, ConvGeom g,
                                                      int tiles_pg, int act, float slope, int bias_gpr) {
  __shared__ float ws[kDirCo][16 * kDirCi + 1];
  int taps = g.KH * g.KW;
  int grp = blockIdx.x / tiles_pg;
  int tile = blockIdx.x - grp * tiles_pg;
  int64_t ppg = (int64_t)g.ipg * g.OH * g.OW;
  int64_t lp = (int64_t)tile * kDirPix + (threadIdx.x & (kDirPix - 1));
  bool valid = lp < ppg;
  int64_t p = (int64_t)grp * ppg + (valid ? lp : 0);
  int ox = (int)(p % g.OW);
  int64_t t = p / g.OW;
  int oy = (int)(t % g.OH);
  int n = (int)(t / g.OH);
  int co_l = (threadIdx.x / kDirPix) * 8;
  int co0 = blockIdx.y * kDirCo;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const T* wg = w + (int64_t)grp * g.Cout * taps * g.Cin;
  for (int ci0 = 0; ci0 < g.Cin; ci0 += kDirCi) {
    int cc = g.Cin - ci0 < kDirCi ? g.Cin - ci0 : kDirCi;
    __syncthreads();
    for (int idx = threadIdx.x; idx < kDirCo * taps * cc; idx += blockDim.x) {
      int co = idx / (taps * cc);
      int rem = idx - co * taps * cc;
      int tap = rem / cc, ci = rem - tap * cc;
      float v = 0.f;
      if (co0 + co < g.Cout) v = ldf<T>(wg + ((int64_t)(co0 + co) * taps + tap) * g.Cin + ci0 + ci);
      ws[co][tap * kDirCi + ci] = v;
    }
    __syncthreads();
    if (valid) {
      for (int tap = 0; tap < taps; ++tap) {
        int kh = tap / g.KW, kw = tap - kh * g.KW, iy, ix;
        if (!tap_src(g, oy, ox, kh, kw, iy, ix)) continue;
        const T* xp = x + (((int64_t)n * g.H + iy) * g.W + ix) * g.Cin + ci0;
        for (int ci = 0; ci < cc; ++ci) {
          float xv = ldf<T>(xp + ci);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += xv * ws[co_l + k][tap * kDirCi + ci];
        }
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int co = co0 + co_l + k;
      if (co < g.Cout) {
        float v = acc[k] + (bias ? bias[(bias_gpr ? grp / bias_gpr : 0) * g.Cout + co] : 0.f);
        if (act == RD_ACT_LRELU) v = v > 0.f ? v : v * slope;
        stf<T>(y + p * g.Cout + co, v);
      }
    }
  }
}

static int launch_direct(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias,
                         void* y, cudaStream_t st) {
  ConvGeom g;
  if (mode == 0) {
    g.N = d->n; g.H = d->h; g.W = d->w; g.Cin = d->cin; g.OH = d->oh; g.OW = d->ow; g.Cout = d->cout;
  } else {  // dgrad: input = dy (oh, ow, cout), output = dx (h, w, cin)
    g.N = d->n; g.H = d->oh; g.W = d->ow; g.Cin = d->cout; g.OH = d->h; g.OW = d->w; g.Cout = d->cin;
  }
  g.KH = d->kh; g.KW = d->kw; g.stride = d->stride; g.pad = d->pad; g.mode = mode;
  g.ipg = d->n / d->groups;
  if (g.KH * g.KW > 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "direct conv: at most 16 taps");
  int64_t ppg = (int64_t)g.ipg * g.OH * g.OW;
  int tiles_pg = rd_div_up(ppg, kDirPix);
  dim3 grid(tiles_pg * d->groups, rd_div_up(g.Cout, kDirCo));
  int act = mode == 0 ? d->act : RD_ACT_NONE;
  int bstride = (mode == 0 && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;     // groups per bias row
  if (d->dtype == RD_F32)
    k_conv_direct<float><<<grid, 128, 0, st>>>((const float*)x, (const float*)w, bias, (float*)y, g, tiles_pg, act, d->act_slope, bstride);
  else
    k_conv_direct<bf16><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)w, bias, (bf16*)y, g, tiles_pg, act, d->act_slope, bstride);
  RD_CHECK_LAUNCH(ctx, mode == 0 ? "conv_direct_fwd" : "conv_direct_dgrad");
  return RD_OK;
}

// ---------------------------------------------------------------- direct wgrad
// grid (pixel chunks, co_tiles*taps*ci_tiles, G); block 256 = 32 ci lanes x 8 pixel lanes (warps).
constexpr int kWgPix = 1024;
template <typename T>
__global__ void __launch_bounds__(256) k_wgrad_direct(const T* __restrict__ x, const T* __restrict__ dy,
                                                       float* __restrict__ dK, ConvGeom g, int co_tiles, int ci_tiles) {
  __shared__ float red[8][8][33];
  int taps = g.KH * g.KW;
  int grp = blockIdx.z;
  int yi = blockIdx.y;
  int ci_t = yi % ci_tiles; yi /= ci_tiles;
  int tap = yi % taps;
  int co_t = yi / taps;
  int kh = tap / g.KW, kw = tap - kh * g.KW;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int ci = ci_t * 32 + lane;
  int co0 = co_t * 8;
  int64_t ppg = (int64_t)g.ipg * g.OH * g.OW;
  int64_t p0 = (int64_t)blockIdx.x * kWgPix, p1 = p0 + kWgPix;
  if (p1 > ppg) p1 = ppg;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int64_t lp = p0 + wid; lp < p1; lp += 8) {
    int64_t p = (int64_t)grp * ppg + lp;
    int ox = (int)(p % g.OW);
    int64_t t = p / g.OW;
    int oy = (int)(t % g.OH);
    int n = (int)(t / g.OH);
    int iy = oy * g.stride - g.pad + kh, ix = ox * g.stride - g.pad + kw;
    if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
    float xv = (ci < g.Cin) ? ldf<T>(x + (((int64_t)n * g.H + iy) * g.W + ix) * g.Cin + ci) : 0.f;
    const T* dp = dy + p * g.Cout + co0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (co0 + k < g.Cout) acc[k] += xv * ldf<T>(dp + k);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[wid][k][lane] = acc[k];
  __syncthreads();
  if (wid == 0 && ci < g.Cin) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (co0 + k >= g.Cout) break;
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][k][lane];
      atomicAdd(dK + (((int64_t)grp * g.Cout + co0 + k) * taps + tap) * g.Cin + ci, s);
    }
  }
}

// 1x1 weight gradient of a 16 -> 16 channel layer (the CondConv 1x1 16 -> 7 that ends each decoder half, src/model.py:2612, with dY
// zero-padded to 16 channels): dK[g][co][ci] = sum_p dY[p, co] X[p, ci] is 512 FLOP per 64 bytes read — HBM-bound streaming, for which
// the tensor-core kernel's 2 KB TMA boxes are the wrong tool (0.75 ms per 7.9 M pixels; this kernel: ~0.1 ms).  One thread per
// pixel stream keeps an 8 (co half) x 8 (ci half) accumulator tile in registers (four 64-thread quadrants per block); warp-shuffle
// reduction, then one red.global.add per element and warp.  grid (pixel chunks, groups).
__global__ void __launch_bounds__(256, 2) k_wgrad_1x1_c16(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dK,
                                                           float* __restrict__ dbias, int64_t ppg, int64_t chunk, int dbias_gpr) {
  // Four 64-thread quadrants = (co half, ci half): an 8 x 8 accumulator tile per thread (the first version kept 8 x 16 and ran ONE
  // block of 8 warps per SM: two warps per sub-partition could not overlap their load latency with the other's FMAs, 0.38 ms per
  // 7.9 M pixels against 0.08 ms of traffic).  With 64 accumulators two blocks fit: four warps per sub-partition.
  const int grp = blockIdx.y;
  const int quad = threadIdx.x >> 6, coh = quad >> 1, cih = quad & 1, t = threadIdx.x & 63, lane = threadIdx.x & 31;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  int64_t p1 = p0 + chunk;
  if (p1 > ppg) p1 = ppg;
  const bf16* xg = x + (int64_t)grp * ppg * 16 + cih * 8;
  const bf16* dg = dy + (int64_t)grp * ppg * 16 + coh * 8;
  float acc[8][8], bsum[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    bsum[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  }
  // four pixels per iteration: all 8 16-byte loads are issued before the first FMA
  for (int64_t p = p0 + t; p < p1; p += 4 * 64) {
    uint4 rx[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t q = p + u * 64;
      if (q < p1) {
        rx[u] = __ldg(reinterpret_cast<const uint4*>(xg + q * 16));
        rd[u] = __ldg(reinterpret_cast<const uint4*>(dg + q * 16));
      } else {
        rx[u] = make_uint4(0, 0, 0, 0); rd[u] = rx[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float xv[8], dv[8];
      const uint32_t wx[4] = {rx[u].x, rx[u].y, rx[u].z, rx[u].w};
      const uint32_t wd[4] = {rd[u].x, rd[u].y, rd[u].z, rd[u].w};
#pragma unroll
      for (int b = 0; b < 4; ++b) { xv[2 * b] = __uint_as_float(wx[b] << 16); xv[2 * b + 1] = __uint_as_float(wx[b] & 0xffff0000u); }
#pragma unroll
      for (int b = 0; b < 4; ++b) { dv[2 * b] = __uint_as_float(wd[b] << 16); dv[2 * b + 1] = __uint_as_float(wd[b] & 0xffff0000u); }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        bsum[a] += dv[a];
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(dv[a], xv[b], acc[a][b]);
      }
    }
  }
  float* dKg = dK + ((int64_t)grp * 16 + coh * 8) * 16 + cih * 8;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float v = acc[a][b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == ((a * 8 + b) & 31)) atomicAdd(dKg + a * 16 + b, v);
    }
    if (dbias && cih == 0) {
      float v = bsum[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == a) atomicAdd(dbias + (size_t)(dbias_gpr ? grp / dbias_gpr : 0) * 16 + coh * 8 + a, v);
    }
  }
}
// dbias[c] += sum over all pixels of dy[., c]
template <typename T>
__global__ void k_bias_grad(const T* __restrict__ dy, float* __restrict__ dbias, int64_t pixels, int C) {
  __shared__ float sm[8][33];
  int c = blockIdx.y * 32 + threadIdx.x;
  int64_t p0 = (int64_t)blockIdx.x * 4096, p1 = p0 + 4096;
  if (p1 > pixels) p1 = pixels;
  float a = 0.f;
  if (c < C)
    for (int64_t p = p0 + threadIdx.y; p < p1; p += 8) a += ldf<T>(dy + p * C + c);
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][threadIdx.x];
    atomicAdd(dbias + c, s);
  }
}

// bf16, C % 8 == 0: 16-byte loads, each block covers (256 / (C/8)) pixel lanes x 64 pixels, shared-memory combine, atomics
__global__ void __launch_bounds__(256) k_bias_grad_vec(const bf16* __restrict__ dy, float* __restrict__ dbias, int64_t pixels,
                                                        int C) {
  __shared__ float acc_s[2048];
  const int cv = C / 8;
  const int cvt = cv < 256 ? cv : 256;
  const int lanes = 256 / cvt;
  for (int i = threadIdx.x; i < cvt * 8; i += 256) acc_s[i] = 0.f;
  __syncthreads();
  const int vec = threadIdx.x % cvt, pl = threadIdx.x / cvt;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (pl < lanes) {
    int64_t p0 = (int64_t)blockIdx.x * lanes * 64;
    for (int it = 0; it < 64; ++it) {
      int64_t p = p0 + (int64_t)it * lanes + pl;
      if (p < pixels) {
        float v[8];
        VecIO<bf16>::load(dy + p * C + vec * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&acc_s[vec * 8 + k], a[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cvt * 8; i += 256) atomicAdd(dbias + i, acc_s[i]);
}

static int check_desc(rd_ctx* ctx, const rd_conv_desc* d) {
  if (!d) RD_FAIL(ctx, RD_ERR_ARG, "conv: null desc");
  if (d->groups < 1 || d->n % d->groups) RD_FAIL(ctx, RD_ERR_ARG, "conv: n (%d) must be a multiple of groups (%d)", d->n, d->groups);
  if (d->dtype != RD_F32 && d->dtype != RD_BF16) RD_FAIL(ctx, RD_ERR_ARG, "conv: bad dtype");
  if (d->bias_groups > 1 && d->groups % d->bias_groups) RD_FAIL(ctx, RD_ERR_ARG, "conv: bias_groups must divide groups");
  int eoh = (d->h + 2 * d->pad - d->kh) / d->stride + 1, eow = (d->w + 2 * d->pad - d->kw) / d->stride + 1;
  if (eoh != d->oh || eow != d->ow) RD_FAIL(ctx, RD_ERR_ARG, "conv: output size %dx%d does not match %dx%d", d->oh, d->ow, eoh, eow);
  return RD_OK;
}


// 1x1 convolution with 16 input channels and <= 16 output channels (the CondConv 1x1 16 -> 7 that ends each decoder half and its
// input gradient 16 (zero-padded 7) -> 16): 32 FLOP per byte moved, i.e. an HBM-bound stream — on the tensor-core kernel one
// 128-pixel tile is ONE K = 16 MMA, so the kernel runs at its per-tile overhead (0.49 ms for 7.9 M pixels, 0.06 ms of traffic).
//   y[p, co] = act(bias[co] + sum_ci x[p, ci] w[g][co][ci])        (forward: w = packed;  dgrad: w = packedT, x = dY)
// One thread per pixel: 32-byte coalesced row load, weights of the block's group broadcast from shared memory (fp32), output tile
// staged in shared memory and written as one contiguous run of 16-byte vectors.  grid (pixel chunks of 1024, groups).
template <int COUT>
__global__ void __launch_bounds__(256) k_conv1x1_c16(const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ bias,
                                                      bf16* __restrict__ y, int64_t ppg, int act, float slope, int bias_gpr) {
  __shared__ __align__(16) float ws[COUT][16];
  __shared__ float bs[COUT];
  __shared__ __align__(16) bf16 stage[256 * COUT];
  const int grp = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < COUT * 16; i += 256) ws[i >> 4][i & 15] = __bfloat162float(w[(int64_t)grp * COUT * 16 + i]);
  if (t < COUT) bs[t] = bias ? bias[(size_t)(bias_gpr ? grp / bias_gpr : 0) * COUT + t] : 0.f;
  __syncthreads();
  const bf16* xg = x + (int64_t)grp * ppg * 16;
  bf16* yg = y + (int64_t)grp * ppg * COUT;
  const int64_t c0 = (int64_t)blockIdx.x * 1024;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int64_t p0 = c0 + r * 256;
    if (p0 >= ppg) break;
    const int64_t p = p0 + t;
    if (p < ppg) {
      float x0[8], x1[8];
      VecIO<bf16>::load(xg + p * 16, x0);
      VecIO<bf16>::load(xg + p * 16 + 8, x1);
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        float a = bs[co];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[co][q * 4]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[co][8 + q * 4]);
          a = fmaf(x0[q * 4 + 0], w0.x, a); a = fmaf(x0[q * 4 + 1], w0.y, a); a = fmaf(x0[q * 4 + 2], w0.z, a); a = fmaf(x0[q * 4 + 3], w0.w, a);
          a = fmaf(x1[q * 4 + 0], w1.x, a); a = fmaf(x1[q * 4 + 1], w1.y, a); a = fmaf(x1[q * 4 + 2], w1.z, a); a = fmaf(x1[q * 4 + 3], w1.w, a);
        }
        if (act == RD_ACT_LRELU) a = a > 0.f ? a : a * slope;
        stage[t * COUT + co] = __float2bfloat16_rn(a);
      }
    }
    __syncthreads();
    const int64_t np = (ppg - p0 < 256) ? (ppg - p0) : 256;            // pixels of this round
    bf16* dst = yg + p0 * COUT;
    const int n16 = (int)(np * COUT / 8);                               // whole 16-byte vectors (the run starts 16-byte aligned: p0 % 256 == 0)
    for (int i = t; i < n16; i += 256) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(stage)[i];
    for (int i = n16 * 8 + t; i < (int)(np * COUT); i += 256) dst[i] = stage[i];
    __syncthreads();
  }
}
static inline bool conv1x1_c16_ok(const rd_conv_desc* d, int cin, int cout, const void* x, const void* y) {
  static const bool off = getenv("RD_B200_NO_CONV_1X1") != nullptr;
  if (off || d->dtype != RD_BF16 || d->algo != RD_ALGO_AUTO) return false;
  if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad != 0 || cin != 16) return false;
  if (cout != 7 && cout != 16) return false;
  const int64_t ppg = (int64_t)(d->n / d->groups) * d->oh * d->ow;
  if ((ppg * cout * 2) % 16) return false;                               // every group's output run starts 16-byte aligned
  return ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
}
static int launch_conv1x1_c16(rd_ctx* ctx, const rd_conv_desc* d, int cout, const void* x, const void* w, const float* bias, void* y, int act,
                              cudaStream_t s) {
  const int64_t ppg = (int64_t)(d->n / d->groups) * d->oh * d->ow;
  dim3 grid((unsigned)rd_div_up(ppg, 1024), d->groups);
  const int bgpr = (bias && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;
  if (cout == 7) k_conv1x1_c16<7><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)w, bias, (bf16*)y, ppg, act, d->act_slope, bgpr);
  else k_conv1x1_c16<16><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)w, bias, (bf16*)y, ppg, act, d->act_slope, bgpr);
  RD_CHECK_LAUNCH(ctx, "conv1x1_c16");
  ctx->last_conv_algo = RD_ALGO_DIRECT;
  return RD_OK;
}

extern "C" int rd_conv2d_fwd(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* packed, const float* bias,
                             void* y, rd_stream st) {
  int rc = check_desc(ctx, d);
  if (rc) return rc;
  if (conv1x1_c16_ok(d, d->cin, d->cout, x, y)) return launch_conv1x1_c16(ctx, d, d->cout, x, packed, bias, y, d->act, (cudaStream_t)st);
  bool tc_ok = rd_conv_tc_supported(d, 0);
  if (d->algo == RD_ALGO_TCGEN05 && !tc_ok) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv fwd: shape not supported by the tcgen05 kernel");
  if (d->algo == RD_ALGO_HALO && !rd_conv_halo_supported(d, 0, ctx->sm_count, 1)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv fwd: shape not supported by the halo kernel");
  if (tc_ok && d->algo != RD_ALGO_DIRECT) {
    ctx->last_conv_algo = RD_ALGO_TCGEN05;
    return rd_conv_tc_launch(ctx, d, 0, x, packed, bias, y, (cudaStream_t)st);
  }
  ctx->last_conv_algo = RD_ALGO_DIRECT;
  return launch_direct(ctx, d, 0, x, packed, bias, y, (cudaStream_t)st);
}

// gamma|beta convolution of a SPADE block with the modulation fused into its epilogue (halo kernel only; the caller falls back to
// rd_conv2d_fwd + rd_spade_modulate_fwd where this returns 0)
extern "C" int rd_conv2d_fwd_spade_supported(rd_ctx* ctx, const rd_conv_desc* d) {
  static const bool off = getenv("RD_B200_NO_SPADE_FUSE") != nullptr;
  if (off || check_desc(ctx, d)) return 0;
  if (d->dtype != RD_BF16 || (d->algo != RD_ALGO_AUTO && d->algo != RD_ALGO_HALO)) return 0;
  return rd_conv_halo_spade_supported(d, ctx->sm_count);
}
extern "C" int rd_conv2d_fwd_spade(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* packed, const float* bias,
                                   const void* z, const float* mean, const float* invstd, void* gamma, void* mix, rd_stream st) {
  int rc = check_desc(ctx, d);
  if (rc) return rc;
  if (!rd_conv2d_fwd_spade_supported(ctx, d)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv fwd spade: shape not supported");
  ctx->last_conv_algo = RD_ALGO_HALO;
  rd_trace_conv("fwd", "halo_spade", d);
  return rd_conv_halo_spade_launch(ctx, d, x, packed, bias, z, mean, invstd, gamma, mix, (cudaStream_t)st);
}

extern "C" int rd_conv2d_dgrad(rd_ctx* ctx, const rd_conv_desc* d, const void* dy, const void* packedT, void* dx,
                               rd_stream st) {
  int rc = check_desc(ctx, d);
  if (rc) return rc;
  if (conv1x1_c16_ok(d, d->cout, d->cin, dy, dx)) return launch_conv1x1_c16(ctx, d, d->cin, dy, packedT, nullptr, dx, RD_ACT_NONE, (cudaStream_t)st);
  bool tc_ok = rd_conv_tc_supported(d, 1);
  if (d->algo == RD_ALGO_TCGEN05 && !tc_ok) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv dgrad: shape not supported by the tcgen05 kernel");
  if (d->algo == RD_ALGO_HALO && !rd_conv_halo_supported(d, 1, ctx->sm_count, 1)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv dgrad: shape not supported by the halo kernel");
  if (tc_ok && d->algo != RD_ALGO_DIRECT) {
    ctx->last_conv_algo = RD_ALGO_TCGEN05;
    return rd_conv_tc_launch(ctx, d, 1, dy, packedT, nullptr, dx, (cudaStream_t)st);
  }
  ctx->last_conv_algo = RD_ALGO_DIRECT;
  return launch_direct(ctx, d, 1, dy, packedT, nullptr, dx, (cudaStream_t)st);
}

extern "C" int rd_conv2d_wgrad(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                               rd_stream st) {
  int rc = check_desc(ctx, d);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)st;
  int taps = d->kh * d->kw;
  RD_CUDA(ctx, cudaMemsetAsync(dK, 0, sizeof(float) * (size_t)d->groups * d->cout * taps * d->cin, s));
  {
    static const bool no_small = getenv("RD_B200_NO_WGRAD_1X1") != nullptr;
    if (!no_small && d->dtype == RD_BF16 && d->algo == RD_ALGO_AUTO && d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0 &&
        d->cin == 16 && d->cout == 16) {
      const int64_t ppg = (int64_t)(d->n / d->groups) * d->oh * d->ow;
      int64_t chunks = rd_div_up((int64_t)ctx->sm_count * 4, (int64_t)d->groups);      // two resident blocks per SM, two waves
      if (chunks < 1) chunks = 1;
      int64_t chunk = rd_div_up(ppg, chunks);
      chunk = rd_div_up(chunk, (int64_t)256) * 256;
      dim3 grid((unsigned)rd_div_up(ppg, chunk), d->groups);
      k_wgrad_1x1_c16<<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)dy, dK, dbias, ppg, chunk,
                                           d->bias_groups > 1 ? d->groups / d->bias_groups : 0);
      RD_CHECK_LAUNCH(ctx, "wgrad_1x1_c16");
      ctx->last_conv_algo = RD_ALGO_DIRECT;
      return RD_OK;
    }
  }
  bool tc_ok = rd_wgrad_tc_supported(d);
  if (d->algo == RD_ALGO_TCGEN05 && !tc_ok) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv wgrad: shape not supported by the tcgen05 kernel");
  if (d->algo == RD_ALGO_HALO && !rd_wgrad_halo_supported(d, ctx->sm_count)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv wgrad: shape not supported by the halo kernel");
  if (tc_ok && d->algo != RD_ALGO_DIRECT) {
    ctx->last_conv_algo = RD_ALGO_TCGEN05;
    static const bool no_tma = getenv("RD_B200_NO_TMA") != nullptr;
    if (!no_tma && rd_wgrad_halo_supported(d, ctx->sm_count)) {
      rd_trace_conv("wgrad", "halo", d);
      return rd_wgrad_halo_launch(ctx, d, x, dy, dK, dbias, s);  // halo-resident X tile: C <= 64 3x3 layers
    } else if (!no_tma && rd_wgrad_tma_supported(d)) {
      rd_trace_conv("wgrad", "tma", d);
      return rd_wgrad_tma_launch(ctx, d, x, dy, dK, dbias, s);  // bias gradient = one extra N=16 MMA against a ones block
    } else {
      rd_trace_conv("wgrad", "tc_gather", d);
      return rd_wgrad_tc_launch(ctx, d, x, dy, dK, dbias, s);  // the bias gradient rides along as a "ones" im2col column
    }
  } else {
    ctx->last_conv_algo = RD_ALGO_DIRECT;
    ConvGeom g;
    g.N = d->n; g.H = d->h; g.W = d->w; g.Cin = d->cin; g.OH = d->oh; g.OW = d->ow; g.Cout = d->cout;
    g.KH = d->kh; g.KW = d->kw; g.stride = d->stride; g.pad = d->pad; g.mode = 0; g.ipg = d->n / d->groups;
    int64_t ppg = (int64_t)g.ipg * g.OH * g.OW;
    int co_tiles = rd_div_up(g.Cout, 8), ci_tiles = rd_div_up(g.Cin, 32);
    dim3 grid(rd_div_up(ppg, kWgPix), co_tiles * taps * ci_tiles, d->groups);
    if (d->dtype == RD_F32)
      k_wgrad_direct<float><<<grid, 256, 0, s>>>((const float*)x, (const float*)dy, dK, g, co_tiles, ci_tiles);
    else
      k_wgrad_direct<bf16><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)dy, dK, g, co_tiles, ci_tiles);
    RD_CHECK_LAUNCH(ctx, "wgrad_direct");
  }
  // bias gradient of the CUDA-core path: one reduction per bias row (all pixels, or one weight group's pixels)
  const int brows = d->bias_groups > 1 ? d->bias_groups : 1;
  const int64_t pixels = (int64_t)d->n * d->oh * d->ow / brows;
  const size_t esz = d->dtype == RD_F32 ? 4 : 2;
  for (int b = 0; b < brows && dbias; ++b) {
    const char* dyb = (const char*)dy + (size_t)b * pixels * d->cout * esz;
    float* dbb = dbias + (size_t)b * d->cout;
    if (d->dtype == RD_BF16 && d->cout % 8 == 0) {
      int cv = d->cout / 8;
      int lanes = 256 / (cv < 256 ? cv : 256);
      int64_t per_block = (int64_t)lanes * 64;
      k_bias_grad_vec<<<rd_div_up(pixels, per_block), 256, 0, s>>>((const bf16*)dyb, dbb, pixels, d->cout);
      RD_CHECK_LAUNCH(ctx, "bias_grad_vec");
    } else {
      dim3 grid(rd_div_up(pixels, 4096), rd_div_up(d->cout, 32)), block(32, 8);
      if (d->dtype == RD_F32) k_bias_grad<float><<<grid, block, 0, s>>>((const float*)dyb, dbb, pixels, d->cout);
      else k_bias_grad<bf16><<<grid, block, 0, s>>>((const bf16*)dyb, dbb, pixels, d->cout);
      RD_CHECK_LAUNCH(ctx, "bias_grad");
    }
  }
  return RD_OK;
}
