// rd_metrics.cu — the callers / data formats either side of the hot path:
//   * evaluation metrics on the device: PSNR / SSIM / MSE of `compute_reconstruction_metrics_single` (reference src/util.py:956-978,
//     skimage.metrics with data_range = max(target - min(target))) and Dice / IoU of `compute_segmentation_metrics_single`
//     (src/util.py:980-992), one result row per image, no D2H of the images;
//   * slab assembly from a device-resident volume store: `ZeroDoseDataset.__getitem__` (src/util.py:471-566) for a whole batch —
//     7-slice window, missing / dropped contrasts zeroed, BraTS label 4 -> 3, optional skull strip, `mask_img = (inputs[0] == 0)`.
// All HBM-bound streaming kernels: plane copies in 16-byte vectors, per-image block reductions by warp shuffles.
#include "rd_common.cuh"

namespace {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float block_min(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 3.4e38f;
  if (wid == 0) r = warp_min(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) { return -block_min(-v, red); }
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// element (n, y, x) of channel c0 of an NHWC tensor with C channels
template <typename T>
__device__ __forceinline__ float px(const T* base, int64_t img_off, int y, int x, int W, int C) {
  return ldf<T>(base + img_off + ((int64_t)y * W + x) * C);
}

// ---------------------------------------------------------------- pass 1: per-image extrema
// stats[n] = {min(target), max(target), min(pred)}; one block per image
template <typename TT, typename TP>
__global__ void k_metrics_extrema(const TT* __restrict__ target, const int32_t* __restrict__ t_index, const TP* __restrict__ pred,
                                  float* __restrict__ stats, int H, int W, int Ct, int Cp, int ct0, int cp0) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const int tn = t_index ? t_index[n] : n;
  const int64_t toff = (int64_t)tn * H * W * Ct + ct0, poff = (int64_t)n * H * W * Cp + cp0;
  float tmin = 3.4e38f, tmax = -3.4e38f, pmin = 3.4e38f;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const float t = ldf<TT>(target + toff + (int64_t)i * Ct), p = ldf<TP>(pred + poff + (int64_t)i * Cp);
    tmin = fminf(tmin, t); tmax = fmaxf(tmax, t); pmin = fminf(pmin, p);
  }
  tmin = block_min(tmin, red);
  tmax = block_max(tmax, red);
  pmin = block_min(pmin, red);
  if (threadIdx.x == 0) { stats[3 * n] = tmin; stats[3 * n + 1] = tmax; stats[3 * n + 2] = pmin; }
}

// ---------------------------------------------------------------- pass 2: MSE + SSIM partial sums
// skimage.metrics.structural_similarity defaults: 7x7 uniform window, sample covariance (cov_norm = 49 / 48), K1 = 0.01, K2 = 0.03,
// the SSIM map averaged over the interior [3, H-3) x [3, W-3).  a = target - min(target), b = pred - min(pred), R = max(a).
// grid (tiles_x, tiles_y, N), 16 x 16 output pixels per block, (16+6)^2 halo in shared memory; partial[n][tile] = {sum sq err, sum ssim}
constexpr int kMT = 16, kMH = kMT + 6;
template <typename TT, typename TP>
__global__ void __launch_bounds__(256) k_metrics_partial(const TT* __restrict__ target, const int32_t* __restrict__ t_index, const TP* __restrict__ pred,
                                                         const float* __restrict__ stats, double* __restrict__ partial, int H, int W, int Ct, int Cp,
                                                         int ct0, int cp0) {
  __shared__ float sa[kMH][kMH + 1], sb[kMH][kMH + 1];
  __shared__ double redd[2][8];
  const int n = blockIdx.z;
  const int tn = t_index ? t_index[n] : n;
  const int64_t toff = (int64_t)tn * H * W * Ct + ct0, poff = (int64_t)n * H * W * Cp + cp0;
  const float tmin = stats[3 * n], tmax = stats[3 * n + 1], pmin = stats[3 * n + 2];
  const float R = tmax - tmin;
  const float C1 = (0.01f * R) * (0.01f * R), C2 = (0.03f * R) * (0.03f * R);
  const int y0 = blockIdx.y * kMT, x0 = blockIdx.x * kMT;
  for (int i = threadIdx.x; i < kMH * kMH; i += blockDim.x) {
    const int hy = i / kMH, hx = i - hy * kMH;
    const int y = y0 + hy - 3, x = x0 + hx - 3;
    float a = 0.f, b = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      a = ldf<TT>(target + toff + ((int64_t)y * W + x) * Ct) - tmin;
      b = ldf<TP>(pred + poff + ((int64_t)y * W + x) * Cp) - pmin;
    }
    sa[hy][hx] = a; sb[hy][hx] = b;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int y = y0 + ty, x = x0 + tx;
  double se = 0.0, ss = 0.0;
  if (y < H && x < W) {
    const float d = sa[ty + 3][tx + 3] - sb[ty + 3][tx + 3];
    se = (double)d * (double)d;
    if (y >= 3 && y < H - 3 && x >= 3 && x < W - 3) {
      float s_a = 0.f, s_b = 0.f, s_aa = 0.f, s_bb = 0.f, s_ab = 0.f;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy)
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const float a = sa[ty + dy][tx + dx], b = sb[ty + dy][tx + dx];
          s_a += a; s_b += b; s_aa = fmaf(a, a, s_aa); s_bb = fmaf(b, b, s_bb); s_ab = fmaf(a, b, s_ab);
        }
      const float inv = 1.f / 49.f, cov = 49.f / 48.f;
      const float ux = s_a * inv, uy = s_b * inv, uxx = s_aa * inv, uyy = s_bb * inv, uxy = s_ab * inv;
      const float vx = cov * (uxx - ux * ux), vy = cov * (uyy - uy * uy), vxy = cov * (uxy - ux * uy);
      const float A1 = 2.f * ux * uy + C1, A2 = 2.f * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
      ss = (double)((A1 * A2) / (B1 * B2));
    }
  }
  se = warp_sum_d(se); ss = warp_sum_d(ss);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { redd[0][wid] = se; redd[1][wid] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) { a += redd[0][k]; b += redd[1][k]; }
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tiles = gridDim.x * gridDim.y;
    partial[((int64_t)n * tiles + tile) * 2] = a;
    partial[((int64_t)n * tiles + tile) * 2 + 1] = b;
  }
}
// out[n] = {ssim, psnr, mse ("rmse" key of the reference: skimage mean_squared_error)}
__global__ void k_metrics_finalize(const double* __restrict__ partial, const float* __restrict__ stats, float* __restrict__ out, int N,
                                   int tiles, int H, int W) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double se = 0.0, ss = 0.0;
  for (int t = 0; t < tiles; ++t) { se += partial[((int64_t)n * tiles + t) * 2]; ss += partial[((int64_t)n * tiles + t) * 2 + 1]; }
  const double mse = se / ((double)H * W);
  const double R = (double)(stats[3 * n + 1] - stats[3 * n]);
  const double interior = (double)(H - 6) * (double)(W - 6);
  out[3 * n] = (float)(ss / interior);
  out[3 * n + 1] = (float)(10.0 * log10((R * R) / mse));          // inf when mse == 0, like skimage (which warns)
  out[3 * n + 2] = (float)mse;
}

// ---------------------------------------------------------------- Dice / IoU (src/util.py:980-992)
// class i in {0, 1, 2}: target == i + 1 against pred[:, i] > 0.5 (the reference indexes the prediction channel by i, not i + 1)
template <typename TP>
__global__ void k_metrics_seg(const float* __restrict__ target, const TP* __restrict__ pred, float* __restrict__ out, int HW, int Cp) {
  __shared__ int red[4][3][8];
  const int n = blockIdx.x;
  int inter[3] = {0, 0, 0}, uni[3] = {0, 0, 0}, ts[3] = {0, 0, 0}, ps[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float t = target[(int64_t)n * HW + i];
    const TP* p = pred + ((int64_t)n * HW + i) * Cp;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const bool tb = (t == (float)(c + 1)), pb = ldf<TP>(p + c) > 0.5f;
      inter[c] += (tb && pb); uni[c] += (tb || pb); ts[c] += tb; ps[c] += pb;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v[4] = {inter[c], uni[c], ts[c], ps[c]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if (lane == 0) red[k][c][wid] = v[k];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    double dice = 0.0, iou = 0.0;
    for (int c = 0; c < 3; ++c) {
      long long I = 0, U = 0, T = 0, P = 0;
      for (int w = 0; w < nw; ++w) { I += red[0][c][w]; U += red[1][c][w]; T += red[2][c][w]; P += red[3][c][w]; }
      dice += (2.0 * (double)I + 1.0) / ((double)T + (double)P + 1.0);
      iou += ((double)I + 1.0) / ((double)U + 1.0);
    }
    out[2 * n] = (float)(dice / 3.0);
    out[2 * n + 1] = (float)(iou / 3.0);
  }
}

// ---------------------------------------------------------------- slab assembly (src/util.py:471-566)
// grid (planes = M * C + 1, B): plane p < M*C of sample b = slice (slice_idx - block + k) of contrast m of subject subj[b] (zeros when
// the contrast is absent or dropped); plane M*C = the target slice.  H*W is a multiple of 4 (float4 copies).
__global__ void __launch_bounds__(256) k_assemble_slabs(const float* __restrict__ vols, const uint8_t* __restrict__ present,
                                                        const float* __restrict__ tvols, const uint8_t* __restrict__ has_target,
                                                        const float* __restrict__ brain_mask, const int32_t* __restrict__ subj,
                                                        const int32_t* __restrict__ slice_idx, const int32_t* __restrict__ drop,
                                                        float* __restrict__ inputs, float* __restrict__ targets, float* __restrict__ mask,
                                                        int M, int C, int D, int HW, int block, int remap4, int clamp_hi) {
  const int b = blockIdx.y, p = blockIdx.x;
  const int s = subj[b];
  int sl = slice_idx[b];
  if (sl < block) sl = block;                           // src/util.py:476-483: slice_idx clamped to [block, 155 - block] (89 for 'Tau')
  if (sl > clamp_hi - block) sl = clamp_hi - block;
  if (sl + block > D - 1) sl = D - 1 - block;           // never read past the stored volume (the host filters such samples like the reference drops them)
  const int64_t hw4 = HW >> 2;
  if (p < M * C) {
    const int m = p / C, k = p - m * C;
    const bool on = present[(int64_t)s * M + m] != 0 && drop[b] != m;
    if (k == 0 && threadIdx.x == 0) mask[(int64_t)b * M + m] = on ? 1.f : 0.f;
    const int z = sl - block + k;
    float4* dst = reinterpret_cast<float4*>(inputs + ((int64_t)b * M * C + p) * HW);
    const float4* bm = brain_mask ? reinterpret_cast<const float4*>(brain_mask + (int64_t)z * HW) : nullptr;
    if (on) {
      const float4* src = reinterpret_cast<const float4*>(vols + (((int64_t)s * M + m) * D + z) * HW);
      for (int64_t i = threadIdx.x; i < hw4; i += blockDim.x) {
        float4 v = src[i];
        if (bm) { const float4 w = bm[i]; v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w; }
        dst[i] = v;
      }
    } else {
      for (int64_t i = threadIdx.x; i < hw4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(targets + (int64_t)b * HW);
    const float4* bm = brain_mask ? reinterpret_cast<const float4*>(brain_mask + (int64_t)sl * HW) : nullptr;
    if (tvols != nullptr && has_target[s]) {
      const float4* src = reinterpret_cast<const float4*>(tvols + ((int64_t)s * D + sl) * HW);
      for (int64_t i = threadIdx.x; i < hw4; i += blockDim.x) {
        float4 v = src[i];
        if (remap4) { v.x = v.x == 4.f ? 3.f : v.x; v.y = v.y == 4.f ? 3.f : v.y; v.z = v.z == 4.f ? 3.f : v.z; v.w = v.w == 4.f ? 3.f : v.w; }
        if (bm) { const float4 w = bm[i]; v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w; }
        dst[i] = v;
      }
    } else {
      for (int64_t i = threadIdx.x; i < hw4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// mask_img[b] = (inputs[b, 0] == 0) after the assembly (src/util.py:563-564)
__global__ void k_mask_img(const float* __restrict__ inputs, float* __restrict__ mask_img, int64_t HW, int64_t img_stride, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, e = i - b * HW;
    mask_img[i] = inputs[b * img_stride + e] == 0.f ? 1.f : 0.f;
  }
}

}  // namespace

extern "C" int rd_metrics_recon_tiles(int H, int W) { return rd_div_up(W, kMT) * rd_div_up(H, kMT); }

extern "C" int rd_metrics_recon(rd_ctx* ctx, const void* target, int t_dtype, int Ct, int ct0, const int32_t* t_index, const void* pred,
                                int p_dtype, int Cp, int cp0, int N, int H, int W, float* stats, double* partial, float* out, rd_stream st) {
  if (N < 1) return RD_OK;
  if (H < 7 || W < 7) RD_FAIL(ctx, RD_ERR_ARG, "metrics_recon: images must be at least 7 x 7 (SSIM window)");
  cudaStream_t s = (cudaStream_t)st;
  dim3 grid(rd_div_up(W, kMT), rd_div_up(H, kMT), N);
  const int tiles = grid.x * grid.y;
#define RD_METRICS_LAUNCH(TT, TP)                                                                                                   \
  do {                                                                                                                               \
    k_metrics_extrema<TT, TP><<<N, 512, 0, s>>>((const TT*)target, t_index, (const TP*)pred, stats, H, W, Ct, Cp, ct0, cp0);          \
    RD_CHECK_LAUNCH(ctx, "metrics_extrema");                                                                                         \
    k_metrics_partial<TT, TP><<<grid, 256, 0, s>>>((const TT*)target, t_index, (const TP*)pred, stats, partial, H, W, Ct, Cp, ct0, cp0); \
    RD_CHECK_LAUNCH(ctx, "metrics_partial");                                                                                         \
  } while (0)
  if (t_dtype == RD_F32 && p_dtype == RD_F32) RD_METRICS_LAUNCH(float, float);
  else if (t_dtype == RD_F32 && p_dtype == RD_BF16) RD_METRICS_LAUNCH(float, bf16);
  else if (t_dtype == RD_BF16 && p_dtype == RD_BF16) RD_METRICS_LAUNCH(bf16, bf16);
  else RD_FAIL(ctx, RD_ERR_ARG, "metrics_recon: unsupported dtype combination");
#undef RD_METRICS_LAUNCH
  k_metrics_finalize<<<rd_div_up(N, 128), 128, 0, s>>>(partial, stats, out, N, tiles, H, W);
  RD_CHECK_LAUNCH(ctx, "metrics_finalize");
  return RD_OK;
}

extern "C" int rd_metrics_seg(rd_ctx* ctx, const float* target, const void* pred, int p_dtype, int Cp, int N, int64_t hw, float* out,
                              rd_stream st) {
  if (N < 1) return RD_OK;
  if (Cp < 3) RD_FAIL(ctx, RD_ERR_ARG, "metrics_seg: the prediction needs at least 3 channels");
  cudaStream_t s = (cudaStream_t)st;
  if (p_dtype == RD_F32) k_metrics_seg<float><<<N, 256, 0, s>>>(target, (const float*)pred, out, (int)hw, Cp);
  else if (p_dtype == RD_BF16) k_metrics_seg<bf16><<<N, 256, 0, s>>>(target, (const bf16*)pred, out, (int)hw, Cp);
  else RD_FAIL(ctx, RD_ERR_ARG, "metrics_seg: bad dtype");
  RD_CHECK_LAUNCH(ctx, "metrics_seg");
  return RD_OK;
}

extern "C" int rd_assemble_slabs(rd_ctx* ctx, const float* vols, const uint8_t* present, const float* tvols, const uint8_t* has_target,
                                 const float* brain_mask, const int32_t* subj, const int32_t* slice_idx, const int32_t* drop, float* inputs,
                                 float* targets, float* mask, float* mask_img, int B, int M, int block, int D, int H, int W, int remap4,
                                 int clamp_hi, rd_stream st) {
  if (B < 1) return RD_OK;
  const int C = 2 * block + 1;
  const int64_t HW = (int64_t)H * W;
  if (HW % 4) RD_FAIL(ctx, RD_ERR_ARG, "assemble_slabs: H * W must be a multiple of 4");
  if (D < C) RD_FAIL(ctx, RD_ERR_ARG, "assemble_slabs: the volumes have fewer slices than one window");
  cudaStream_t s = (cudaStream_t)st;
  dim3 grid(M * C + 1, B);
  k_assemble_slabs<<<grid, 256, 0, s>>>(vols, present, tvols, has_target, brain_mask, subj, slice_idx, drop, inputs, targets, mask, M, C, D,
                                        (int)HW, block, remap4, clamp_hi);
  RD_CHECK_LAUNCH(ctx, "assemble_slabs");
  const int64_t total = (int64_t)B * HW;
  k_mask_img<<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, s>>>(inputs, mask_img, HW, (int64_t)M * C * HW, total);
  RD_CHECK_LAUNCH(ctx, "mask_img");
  return RD_OK;
}
