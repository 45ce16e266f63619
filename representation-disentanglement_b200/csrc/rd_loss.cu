// rd_loss.cu — missing-modality fusion gather, reconstruction / latent / similarity / KL / segmentation losses
// with their gradients, and the clip_grad_norm_ + Adam(amsgrad) optimizer.  The masked skip / index logic
// of the reference's Python loops (src/model.py:3268-3341, 3384-3394, 3478-3557) runs on the device, so
// there is no `mask.sum() == 0` host synchronisation and the step is CUDA-graph capturable.
#include "rd_common.cuh"

// ============================================================================ fusion gather (Q3)
// grid (B*M, chunks); one block = one chunk of one candidate row.  Flags staged in shared memory,
// exclusive prefix count over the (b, m) row-major order gives the destination row.
template <typename T>
__global__ void k_fuse_gather(const T* __restrict__ si, const float* __restrict__ mask, T* __restrict__ out,
                              int32_t* __restrict__ idx_out, int32_t* __restrict__ count_out, int B, int M,
                              int64_t row_elems, int64_t chunk) {
  __shared__ int flags[1024];
  int R = B * M;
  for (int i = threadIdx.x; i < R; i += blockDim.x) flags[i] = (mask[i] == 1.0f) ? 1 : 0;
  __syncthreads();
  int row = blockIdx.x;   // = b*M + m
  int dst = 0;
  for (int i = 0; i < row; ++i) dst += flags[i];
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    int k = 0;
    for (int i = 0; i < R; ++i) if (flags[i]) { if (idx_out) idx_out[k] = i; ++k; }
    if (count_out) *count_out = k;
    if (idx_out) for (int i = k; i < R; ++i) idx_out[i] = -1;
  }
  if (!flags[row]) return;
  int b = row / M, m = row - b * M;
  const T* src = si + ((int64_t)m * B + b) * row_elems;
  T* d = out + (int64_t)dst * row_elems;
  int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = e0 + chunk;
  if (e1 > row_elems) e1 = row_elems;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) d[e] = src[e];
}
template <typename T>
__global__ void k_fuse_scatter(const T* __restrict__ dout, const float* __restrict__ mask, T* __restrict__ dsi, int B,
                               int M, int64_t row_elems, int64_t chunk) {
  __shared__ int flags[1024];
  int R = B * M;
  for (int i = threadIdx.x; i < R; i += blockDim.x) flags[i] = (mask[i] == 1.0f) ? 1 : 0;
  __syncthreads();
  int row = blockIdx.x;
  int src_row = 0;
  for (int i = 0; i < row; ++i) src_row += flags[i];
  int b = row / M, m = row - b * M;
  T* d = dsi + ((int64_t)m * B + b) * row_elems;
  const T* s = dout + (int64_t)src_row * row_elems;
  bool sel = flags[row];
  int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = e0 + chunk;
  if (e1 > row_elems) e1 = row_elems;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    if (sel) d[e] = s[e]; else stf<T>(d + e, 0.f);
  }
}
extern "C" int rd_fuse_gather_fwd(rd_ctx* ctx, const void* si, const float* mask, void* out, int32_t* idx_out,
                                  int32_t* count_out, int B, int M, int64_t row_elems, int dtype, rd_stream st) {
  if (B * M > 1024) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "fuse_gather: B*M <= 1024");
  int64_t chunk = 16384;
  dim3 grid(B * M, rd_div_up(row_elems, chunk));
  RD_DISPATCH_DTYPE(dtype, k_fuse_gather<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)si, mask, (T*)out, idx_out, count_out, B, M, row_elems, chunk));
  RD_CHECK_LAUNCH(ctx, "fuse_gather_fwd");
  return RD_OK;
}
extern "C" int rd_fuse_gather_bwd(rd_ctx* ctx, const void* dout, const float* mask, void* dsi, int B, int M,
                                  int64_t row_elems, int dtype, rd_stream st) {
  if (B * M > 1024) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "fuse_gather: B*M <= 1024");
  int64_t chunk = 16384;
  dim3 grid(B * M, rd_div_up(row_elems, chunk));
  RD_DISPATCH_DTYPE(dtype, k_fuse_scatter<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)dout, mask, (T*)dsi, B, M, row_elems, chunk));
  RD_CHECK_LAUNCH(ctx, "fuse_gather_bwd");
  return RD_OK;
}

// ============================================================================ reconstruction rows
constexpr int64_t kReconChunk = 8192;
// 8-element vector access for the row losses (rows whose length and chunk size are multiples of 8, 16-byte aligned)
template <typename T> __device__ __forceinline__ void recon_load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void recon_load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void recon_load8<bf16>(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
  v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}
template <typename T> __device__ __forceinline__ void recon_store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void recon_store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void recon_store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
template <typename T, typename GT>
__device__ __forceinline__ bool recon_vec_ok(const T* xr, const GT* gr, int64_t row_elems) {
  return (row_elems & 7) == 0 && (kReconChunk & 7) == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(gr)) & 15u) == 0;
}
template <typename T, typename GT>
__global__ void k_recon_partial(const T* __restrict__ x, const GT* __restrict__ gt, const int32_t* __restrict__ gt_index,
                                float* __restrict__ partial, int64_t row_elems, int chunks, int p) {
  __shared__ float red[32];
  int r = blockIdx.y;
  int gi = gt_index ? gt_index[r] : r;
  float acc = 0.f;
  if (gi >= 0) {
    const T* xr = x + (int64_t)r * row_elems;
    const GT* gr = gt + (int64_t)gi * row_elems;
    int64_t e0 = (int64_t)blockIdx.x * kReconChunk, e1 = e0 + kReconChunk;
    if (e1 > row_elems) e1 = row_elems;
    if (recon_vec_ok<T, GT>(xr, gr, row_elems)) {
      // 8 elements per thread and iteration: one 16-byte load of x-hat, 16 / 32 bytes of the target (the element loop ran at 1.9 TB/s)
      for (int64_t e = e0 + 8 * (int64_t)threadIdx.x; e < e1; e += 8 * (int64_t)blockDim.x) {
        float xv[8], gv[8];
        recon_load8<T>(xr + e, xv);
        recon_load8<GT>(gr + e, gv);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = gv[k] - xv[k]; acc += (p == 1) ? fabsf(d) : d * d; }
      }
    } else
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
      float d = ldf<GT>(gr + e) - ldf<T>(xr + e);
      acc += (p == 1) ? fabsf(d) : d * d;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)r * chunks + blockIdx.x] = acc;
}
__global__ void k_recon_finalize(const float* __restrict__ partial, float* __restrict__ row_loss, int R, int chunks,
                                 float inv_n) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[(int64_t)r * chunks + k];
  row_loss[r] = s * inv_n;
}
template <typename T, typename GT>
__global__ void k_recon_bwd(const T* __restrict__ x, const GT* __restrict__ gt, const int32_t* __restrict__ gt_index,
                            const float* __restrict__ coef, T* __restrict__ dx, int64_t row_elems, int p, float inv_n) {
  int r = blockIdx.y;
  int gi = gt_index ? gt_index[r] : r;
  float cf = (gi >= 0) ? coef[r] * inv_n : 0.f;
  const T* xr = x + (int64_t)r * row_elems;
  const GT* gr = gt + (int64_t)(gi >= 0 ? gi : 0) * row_elems;
  T* dr = dx + (int64_t)r * row_elems;
  int64_t e0 = (int64_t)blockIdx.x * kReconChunk, e1 = e0 + kReconChunk;
  if (e1 > row_elems) e1 = row_elems;
  if (recon_vec_ok<T, GT>(xr, gr, row_elems) && (reinterpret_cast<uintptr_t>(dr) & 15u) == 0) {
    for (int64_t e = e0 + 8 * (int64_t)threadIdx.x; e < e1; e += 8 * (int64_t)blockDim.x) {
      float xv[8], gv[8], o[8];
      if (gi >= 0) { recon_load8<T>(xr + e, xv); recon_load8<GT>(gr + e, gv); }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float g = 0.f;
        if (gi >= 0) {
          const float d = xv[k] - gv[k];
          g = (p == 1) ? ((d > 0.f) ? cf : ((d < 0.f) ? -cf : 0.f)) : 2.f * d * cf;
        }
        o[k] = g;
      }
      recon_store8<T>(dr + e, o);
    }
    return;
  }
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    float g = 0.f;
    if (gi >= 0) {
      float d = ldf<T>(xr + e) - ldf<GT>(gr + e);   // d/dx |gt-x|^p
      g = (p == 1) ? ((d > 0.f) ? cf : ((d < 0.f) ? -cf : 0.f)) : 2.f * d * cf;
    }
    stf<T>(dr + e, g);
  }
}
extern "C" int rd_recon_rows_fwd(rd_ctx* ctx, const void* x, const void* gt, int gt_dtype, const int32_t* gt_index,
                                 float* row_loss, float* partial, int R, int64_t row_elems, int p, int dtype, rd_stream st) {
  int chunks = rd_div_up(row_elems, kReconChunk);
  dim3 grid(chunks, R);
  cudaStream_t s = (cudaStream_t)st;
  if (dtype == RD_F32 && gt_dtype == RD_F32) k_recon_partial<float, float><<<grid, 256, 0, s>>>((const float*)x, (const float*)gt, gt_index, partial, row_elems, chunks, p);
  else if (dtype == RD_BF16 && gt_dtype == RD_F32) k_recon_partial<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (const float*)gt, gt_index, partial, row_elems, chunks, p);
  else if (dtype == RD_BF16 && gt_dtype == RD_BF16) k_recon_partial<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)gt, gt_index, partial, row_elems, chunks, p);
  else RD_FAIL(ctx, RD_ERR_ARG, "recon_rows: unsupported dtype combination");
  RD_CHECK_LAUNCH(ctx, "recon_partial");
  k_recon_finalize<<<rd_div_up(R, 128), 128, 0, s>>>(partial, row_loss, R, chunks, 1.f / (float)row_elems);
  RD_CHECK_LAUNCH(ctx, "recon_finalize");
  return RD_OK;
}
extern "C" int rd_recon_rows_bwd(rd_ctx* ctx, const void* x, const void* gt, int gt_dtype, const int32_t* gt_index,
                                 const float* coef, void* dx, int R, int64_t row_elems, int p, int dtype, rd_stream st) {
  int chunks = rd_div_up(row_elems, kReconChunk);
  dim3 grid(chunks, R);
  cudaStream_t s = (cudaStream_t)st;
  float inv_n = 1.f / (float)row_elems;
  if (dtype == RD_F32 && gt_dtype == RD_F32) k_recon_bwd<float, float><<<grid, 256, 0, s>>>((const float*)x, (const float*)gt, gt_index, coef, (float*)dx, row_elems, p, inv_n);
  else if (dtype == RD_BF16 && gt_dtype == RD_F32) k_recon_bwd<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (const float*)gt, gt_index, coef, (bf16*)dx, row_elems, p, inv_n);
  else if (dtype == RD_BF16 && gt_dtype == RD_BF16) k_recon_bwd<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (const bf16*)gt, gt_index, coef, (bf16*)dx, row_elems, p, inv_n);
  else RD_FAIL(ctx, RD_ERR_ARG, "recon_rows: unsupported dtype combination");
  RD_CHECK_LAUNCH(ctx, "recon_bwd");
  return RD_OK;
}

// ---------------------------------------------------------------- masked combination (device-side python loops)
// plan for compute_recon_loss_x_mix_list: row block t of the x_mix stack is paired with gt modality j of the
// t-th NON-skipped (i, j) pair (the reference's index lag).
__global__ void k_xmix_plan(const float* __restrict__ mask, int32_t* __restrict__ gt_index, int B, int M) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int t = 0, P = M * (M - 1);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      if (i == j) continue;
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += mask[b * M + i] * mask[b * M + j];
      if (s == 0.f) continue;
      for (int b = 0; b < B; ++b) gt_index[t * B + b] = j * B + b;
      ++t;
    }
  for (; t < P; ++t)
    for (int b = 0; b < B; ++b) gt_index[t * B + b] = -1;
}
extern "C" int rd_xmix_plan(rd_ctx* ctx, const float* mask, int32_t* gt_index, int B, int M, rd_stream st) {
  k_xmix_plan<<<1, 32, 0, (cudaStream_t)st>>>(mask, gt_index, B, M);
  RD_CHECK_LAUNCH(ctx, "xmix_plan");
  return RD_OK;
}
__global__ void k_masked_combine(const float* __restrict__ row_loss, const float* __restrict__ mask, float* __restrict__ loss,
                                 float* __restrict__ coef, int B, int M, int kind) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  int cnt = 0;
  if (kind == 0) {
    for (int i = 0; i < M; ++i) {
      float ms = 0.f;
      for (int b = 0; b < B; ++b) ms += mask[b * M + i];
      if (ms == 0.f) { for (int b = 0; b < B; ++b) coef[i * B + b] = 0.f; continue; }
      ++cnt;
      float s = 0.f;
      for (int b = 0; b < B; ++b) { s += mask[b * M + i] * row_loss[i * B + b]; coef[i * B + b] = mask[b * M + i] / ms; }
      total += s / ms;
    }
    if (cnt > 0) {
      total /= (float)cnt;
      for (int k = 0; k < M * B; ++k) coef[k] /= (float)cnt;
    }
  } else {
    int P = M * (M - 1);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < M; ++j) {
        if (i == j) continue;
        float ms = 0.f;
        for (int b = 0; b < B; ++b) ms += mask[b * M + i] * mask[b * M + j];
        if (ms == 0.f) continue;
        float s = 0.f;
        for (int b = 0; b < B; ++b) {
          float mm = mask[b * M + i] * mask[b * M + j];
          s += mm * row_loss[cnt * B + b];
          coef[cnt * B + b] = mm / ms;
        }
        total += s / ms;
        ++cnt;
      }
    for (int k = cnt * B; k < P * B; ++k) coef[k] = 0.f;
    if (cnt > 0) {
      total /= (float)cnt;
      for (int k = 0; k < cnt * B; ++k) coef[k] /= (float)cnt;
    }
  }
  loss[0] = total;
}
// w[i] = [contrast i present in some row] / #present contrasts (all 0 when none is): the weights of compute_segmentation_loss_y_list's
// mean over the non-skipped contrasts (src/model.py:3299-3313, `if mask[:, i].sum() == 0: continue`) evaluated on the device, so that
// the stage-2 iteration has no host read of the mask and can be captured in the CUDA graph.
__global__ void k_modality_weights(const float* __restrict__ mask, float* __restrict__ w, int B, int M) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < M; ++i) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += mask[b * M + i];
      w[i] = s != 0.f ? 1.f : 0.f;
      n += s != 0.f ? 1 : 0;
    }
    for (int i = 0; i < M; ++i) w[i] = n ? w[i] / (float)n : 0.f;
  }
}
extern "C" int rd_modality_weights(rd_ctx* ctx, const float* mask, float* w, int B, int M, rd_stream st) {
  k_modality_weights<<<1, 32, 0, (cudaStream_t)st>>>(mask, w, B, M);
  RD_CHECK_LAUNCH(ctx, "modality_weights");
  return RD_OK;
}
extern "C" int rd_masked_combine(rd_ctx* ctx, const float* row_loss, const float* mask, float* loss, float* coef, int B,
                                 int M, int kind, rd_stream st) {
  k_masked_combine<<<1, 32, 0, (cudaStream_t)st>>>(row_loss, mask, loss, coef, B, M, kind);
  RD_CHECK_LAUNCH(ctx, "masked_combine");
  return RD_OK;
}

// ============================================================================ latent z / similarity z / KL
__global__ void k_latent_z(const float* __restrict__ mu, const float* __restrict__ mu_new, const float* __restrict__ mask,
                           float* __restrict__ loss, float* __restrict__ dmu, float* __restrict__ dmu_new, int B, int M, int Z) {
  __shared__ float red[32];
  __shared__ float msum[16];
  __shared__ int cnt_s;
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int i = 0; i < M; ++i) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += mask[b * M + i];
      msum[i] = s;
      if (s != 0.f) ++cnt;
    }
    cnt_s = cnt;
  }
  __syncthreads();
  int cnt = cnt_s;
  float acc = 0.f;
  int total = M * B * Z;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e / (B * Z);
    int b = (e / Z) % B;
    float g = 0.f;
    if (msum[i] != 0.f) {
      float d = mu[e] - mu_new[e];
      float wgt = mask[b * M + i] / msum[i] / (float)cnt;
      acc += wgt * fabsf(d);
      g = (d > 0.f) ? wgt : ((d < 0.f) ? -wgt : 0.f);
    }
    dmu[e] = g;
    dmu_new[e] = -g;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc;
}
extern "C" int rd_latent_z_loss(rd_ctx* ctx, const float* mu, const float* mu_new, const float* mask, float* loss, float* dmu,
                                float* dmu_new, int B, int M, int Z, rd_stream st) {
  if (M > 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "latent_z: M <= 16");
  k_latent_z<<<1, 256, 0, (cudaStream_t)st>>>(mu, mu_new, mask, loss, dmu, dmu_new, B, M, Z);
  RD_CHECK_LAUNCH(ctx, "latent_z_loss");
  return RD_OK;
}

// cosine with the reference's epsilons (compute_cosine, src/model.py:3407-3415) and its gradient
struct CosOut { float c, nx, ny, dot; };
__device__ __forceinline__ CosOut cos_fwd(const float* x, const float* y, int Z) {
  float sx = 0.f, sy = 0.f, d = 0.f;
  for (int k = 0; k < Z; ++k) { sx += x[k] * x[k]; sy += y[k] * y[k]; d += x[k] * y[k]; }
  CosOut o;
  o.nx = fmaxf(sqrtf(sx + 1e-8f), 1e-8f);
  o.ny = fmaxf(sqrtf(sy + 1e-8f), 1e-8f);
  o.dot = d;
  o.c = d / (o.nx * o.ny);
  return o;
}
// adds w * dcos/dx to gx and w * dcos/dy to gy
__device__ __forceinline__ void cos_bwd(const float* x, const float* y, int Z, const CosOut& o, float w, float* gx, float* gy) {
  float inv = 1.f / (o.nx * o.ny);
  for (int k = 0; k < Z; ++k) {
    atomicAdd(gx + k, w * (y[k] * inv - o.c * x[k] / (o.nx * o.nx)));
    atomicAdd(gy + k, w * (x[k] * inv - o.c * y[k] / (o.ny * o.ny)));
  }
}
__global__ void k_sim_z(const float* __restrict__ z, const float* __restrict__ mask, float margin, float* __restrict__ loss,
                        float* __restrict__ dz, int B, int M, int Z) {
  __shared__ float red[32];
  for (int e = threadIdx.x; e < M * B * Z; e += blockDim.x) dz[e] = 0.f;
  __syncthreads();
  // number of counted pairs (identical in every thread)
  int cnt = 0;
  for (int i = 0; i < M - 1; ++i)
    for (int j = i + 1; j < M; ++j) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += mask[b * M + i] * mask[b * M + j] * mask[((b + 1) % B) * M + i];
      if (s != 0.f) ++cnt;
    }
  float acc = 0.f;
  if (cnt > 0) {
    for (int i = 0; i < M - 1; ++i)
      for (int j = i + 1; j < M; ++j) {
        float ms = 0.f;
        for (int b = 0; b < B; ++b) ms += mask[b * M + i] * mask[b * M + j] * mask[((b + 1) % B) * M + i];
        if (ms == 0.f) continue;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
          int bp = (b + 1) % B;
          float mm = mask[b * M + i] * mask[b * M + j] * mask[bp * M + i];
          const float* zi = z + ((int64_t)i * B + b) * Z;
          const float* zj = z + ((int64_t)j * B + b) * Z;
          const float* zp = z + ((int64_t)i * B + bp) * Z;
          CosOut c = cos_fwd(zi, zj, Z), cm = cos_fwd(zi, zp, Z);
          float h = margin - cm.c + c.c;
          if (h > 0.f && mm != 0.f) {
            float wgt = mm / ms / (float)cnt;
            acc += wgt * h;
            cos_bwd(zi, zj, Z, c, wgt, dz + ((int64_t)i * B + b) * Z, dz + ((int64_t)j * B + b) * Z);
            cos_bwd(zi, zp, Z, cm, -wgt, dz + ((int64_t)i * B + b) * Z, dz + ((int64_t)i * B + bp) * Z);
          }
        }
      }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc;
}
extern "C" int rd_sim_z_loss(rd_ctx* ctx, const float* z, const float* mask, float margin, float* loss, float* dz, int B, int M,
                             int Z, rd_stream st) {
  k_sim_z<<<1, 128, 0, (cudaStream_t)st>>>(z, mask, margin, loss, dz, B, M, Z);
  RD_CHECK_LAUNCH(ctx, "sim_z_loss");
  return RD_OK;
}
__global__ void k_kl(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ mask,
                     float* __restrict__ loss, float* __restrict__ dmu, float* __restrict__ dlv, int B, int M, int Z) {
  __shared__ float red[32];
  float ms = 0.f;
  for (int k = 0; k < B * M; ++k) ms += mask[k];
  float scale = 1.f / ms / (float)M;   // 0/0 -> NaN exactly like the reference when every contrast is missing
  float acc = 0.f;
  for (int e = threadIdx.x; e < M * B * Z; e += blockDim.x) {
    int i = e / (B * Z), b = (e / Z) % B;
    float m = mask[b * M + i];
    float ev = expf(lv[e]);
    acc += m * 0.5f * (ev + mu[e] * mu[e] - 1.f - lv[e]);
    dmu[e] = m * mu[e] * scale;
    dlv[e] = m * 0.5f * (ev - 1.f) * scale;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc * scale;
}
extern "C" int rd_kl_loss(rd_ctx* ctx, const float* mu, const float* lv, const float* mask, float* loss, float* dmu, float* dlv,
                          int B, int M, int Z, rd_stream st) {
  k_kl<<<1, 256, 0, (cudaStream_t)st>>>(mu, lv, mask, loss, dmu, dlv, B, M, Z);
  RD_CHECK_LAUNCH(ctx, "kl_loss");
  return RD_OK;
}

// ============================================================================ anatomy similarity (max-pool 16x16 + cosine hinge)
template <typename T>
__global__ void k_maxpool16(const T* __restrict__ s, float* __restrict__ pooled, int32_t* __restrict__ argmax, int N, int H,
                            int W, int C) {
  int PH = H / 16, PW = W / 16;
  int64_t total = (int64_t)N * PH * PW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int pw = (int)(t % PW); t /= PW;
    int ph = (int)(t % PH);
    int n = (int)(t / PH);
    const T* base = s + (int64_t)n * H * W * C + c;
    float best = -INFINITY;
    int bi = (ph * 16) * W + pw * 16;
    for (int dy = 0; dy < 16; ++dy)
      for (int dx = 0; dx < 16; ++dx) {
        int pix = (ph * 16 + dy) * W + pw * 16 + dx;
        float v = ldf<T>(base + (int64_t)pix * C);
        if (v > best || v != v) { best = v; bi = pix; }   // first maximum wins (torch max_pool2d CPU/CUDA scan order)
      }
    int64_t o = (int64_t)n * C * PH * PW + (int64_t)c * PH * PW + ph * PW + pw;   // x_pool.view(B,-1) order of NCHW
    pooled[o] = best;
    argmax[o] = bi;
  }
}
template <typename T>
__global__ void k_maxpool16_bwd(const float* __restrict__ dpooled, const int32_t* __restrict__ argmax, T* __restrict__ ds,
                                int N, int H, int W, int C) {
  int PH = H / 16, PW = W / 16;
  int64_t total = (int64_t)N * C * PH * PW;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    int n = (int)(o / ((int64_t)C * PH * PW));
    int c = (int)((o / (PH * PW)) % C);
    stf<T>(ds + ((int64_t)n * H * W + argmax[o]) * C + c, dpooled[o]);
  }
}
extern "C" int rd_maxpool16_fwd(rd_ctx* ctx, const void* s, float* pooled, int32_t* argmax, int N, int H, int W, int C,
                                int dtype, rd_stream st) {
  if (H % 16 || W % 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "maxpool16: H, W multiples of 16");
  int64_t total = (int64_t)N * (H / 16) * (W / 16) * C;
  RD_DISPATCH_DTYPE(dtype, k_maxpool16<T><<<rd_grid_1d(total, 128, ctx->sm_count), 128, 0, (cudaStream_t)st>>>((const T*)s, pooled, argmax, N, H, W, C));
  RD_CHECK_LAUNCH(ctx, "maxpool16_fwd");
  return RD_OK;
}
extern "C" int rd_maxpool16_bwd(rd_ctx* ctx, const float* dpooled, const int32_t* argmax, void* ds, int N, int H, int W, int C,
                                int dtype, rd_stream st) {
  int64_t total = (int64_t)N * (H / 16) * (W / 16) * C;
  size_t esz = dtype == RD_F32 ? 4 : 2;
  RD_CUDA(ctx, cudaMemsetAsync(ds, 0, esz * (size_t)N * H * W * C, (cudaStream_t)st));
  RD_DISPATCH_DTYPE(dtype, k_maxpool16_bwd<T><<<rd_grid_1d(total, 128, ctx->sm_count), 128, 0, (cudaStream_t)st>>>(dpooled, argmax, (T*)ds, N, H, W, C));
  RD_CHECK_LAUNCH(ctx, "maxpool16_bwd");
  return RD_OK;
}

// compute_compact_s_mean (src/model.py:3453-3456): 16x16 average pool, same output order as the max variant
template <typename T>
__global__ void k_avgpool16(const T* __restrict__ s, float* __restrict__ pooled, int N, int H, int W, int C) {
  int PH = H / 16, PW = W / 16;
  int64_t total = (int64_t)N * PH * PW * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int pw = (int)(t % PW); t /= PW;
    int ph = (int)(t % PH);
    int n = (int)(t / PH);
    const T* base = s + (int64_t)n * H * W * C + c;
    float acc = 0.f;
    for (int dy = 0; dy < 16; ++dy)
      for (int dx = 0; dx < 16; ++dx) acc += ldf<T>(base + (int64_t)((ph * 16 + dy) * W + pw * 16 + dx) * C);
    pooled[(int64_t)n * C * PH * PW + (int64_t)c * PH * PW + ph * PW + pw] = acc * (1.f / 256.f);
  }
}
template <typename T>
__global__ void k_avgpool16_bwd(const float* __restrict__ dpooled, T* __restrict__ ds, int N, int H, int W, int C) {
  int PH = H / 16, PW = W / 16;
  int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int x = (int)(t % W); t /= W;
    int y = (int)(t % H);
    int n = (int)(t / H);
    stf<T>(ds + i, dpooled[(int64_t)n * C * PH * PW + (int64_t)c * PH * PW + (y / 16) * PW + x / 16] * (1.f / 256.f));
  }
}
extern "C" int rd_avgpool16_fwd(rd_ctx* ctx, const void* s, float* pooled, int N, int H, int W, int C, int dtype, rd_stream st) {
  if (H % 16 || W % 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "avgpool16: H, W multiples of 16");
  int64_t total = (int64_t)N * (H / 16) * (W / 16) * C;
  RD_DISPATCH_DTYPE(dtype, k_avgpool16<T><<<rd_grid_1d(total, 128, ctx->sm_count), 128, 0, (cudaStream_t)st>>>((const T*)s, pooled, N, H, W, C));
  RD_CHECK_LAUNCH(ctx, "avgpool16_fwd");
  return RD_OK;
}
extern "C" int rd_avgpool16_bwd(rd_ctx* ctx, const float* dpooled, void* ds, int N, int H, int W, int C, int dtype, rd_stream st) {
  int64_t total = (int64_t)N * H * W * C;
  RD_DISPATCH_DTYPE(dtype, k_avgpool16_bwd<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(dpooled, (T*)ds, N, H, W, C));
  RD_CHECK_LAUNCH(ctx, "avgpool16_bwd");
  return RD_OK;
}

// pooled [M][B][D].  pair = device int32[2] (i, j) chosen on the host (np.random.choice).
__global__ void k_sim_s(const float* __restrict__ pooled, const float* __restrict__ mask, const int32_t* __restrict__ pair,
                        float margin, float* __restrict__ loss, float* __restrict__ dpooled, int B, int M, int D) {
  __shared__ float red[32];
  for (int e = threadIdx.x; e < M * B * D; e += blockDim.x) dpooled[e] = 0.f;
  __syncthreads();
  int i = pair[0], j = pair[1];
  float ms = 0.f;
  for (int b = 0; b < B; ++b) ms += mask[b * M + i] * mask[b * M + j] * mask[((b + 1) % B) * M + i];
  float total = 0.f;
  if (M > 1 && ms > 0.f) {
    for (int b = 0; b < B; ++b) {
      int bp = (b + 1) % B;
      float mm = mask[b * M + i] * mask[b * M + j] * mask[bp * M + i];
      const float* a = pooled + ((int64_t)i * B + b) * D;      // si_c[b]
      const float* q = pooled + ((int64_t)j * B + b) * D;      // sj_c[b]
      const float* c = pooled + ((int64_t)i * B + bp) * D;     // si_perm_c[b]
      float saa = 0.f, sqq = 0.f, scc = 0.f, saq = 0.f, sca = 0.f;
      for (int k = threadIdx.x; k < D; k += blockDim.x) {
        float av = a[k], qv = q[k], cv = c[k];
        saa += av * av; sqq += qv * qv; scc += cv * cv; saq += av * qv; sca += cv * av;
      }
      saa = block_sum(saa, red); sqq = block_sum(sqq, red); scc = block_sum(scc, red);
      saq = block_sum(saq, red); sca = block_sum(sca, red);
      float na = fmaxf(sqrtf(saa + 1e-8f), 1e-8f), nq = fmaxf(sqrtf(sqq + 1e-8f), 1e-8f), nc = fmaxf(sqrtf(scc + 1e-8f), 1e-8f);
      float sim = saq / (na * nq), sim_mix = sca / (nc * na);
      float h = margin - sim + sim_mix;
      if (h > 0.f && mm != 0.f) {
        float wgt = mm / ms;
        total += wgt * h;
        float* ga = dpooled + ((int64_t)i * B + b) * D;
        float* gq = dpooled + ((int64_t)j * B + b) * D;
        float* gc = dpooled + ((int64_t)i * B + bp) * D;
        for (int k = threadIdx.x; k < D; k += blockDim.x) {
          float av = a[k], qv = q[k], cv = c[k];
          // -d sim
          ga[k] += -wgt * (qv / (na * nq) - sim * av / (na * na));
          gq[k] += -wgt * (av / (na * nq) - sim * qv / (nq * nq));
          // + d sim_mix  (cos(c, a))
          gc[k] += wgt * (av / (nc * na) - sim_mix * cv / (nc * nc));
          ga[k] += wgt * (cv / (nc * na) - sim_mix * av / (na * na));
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) loss[0] = total;
}
extern "C" int rd_sim_s_loss(rd_ctx* ctx, const float* pooled, const float* mask, const int32_t* pair, float margin, float* loss,
                             float* dpooled, int B, int M, int D, rd_stream st) {
  k_sim_s<<<1, 256, 0, (cudaStream_t)st>>>(pooled, mask, pair, margin, loss, dpooled, B, M, D);
  RD_CHECK_LAUNCH(ctx, "sim_s_loss");
  return RD_OK;
}

// ============================================================================ segmentation loss (weighted CE + soft Dice)
// 11 global sums: [0] sum w[t], [1] sum w[t]*nll, [2+c] sum p_c g_c, [5+c] sum p_c^2, [8+c] sum g_c  (c = class 1..3 -> 0..2)
__device__ __forceinline__ float seg_w(int t) { return t == 0 ? 1.f : 5.f; }
template <typename T>
__global__ void k_seg_partial(const T* __restrict__ y, const float* __restrict__ target, float* __restrict__ partial,
                              int64_t pixels) {
  __shared__ float red[32];
  float acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    float mx = -INFINITY;
    for (int c = 0; c < 4; ++c) { v[c] = ldf<T>(y + i * 4 + c); mx = fmaxf(mx, v[c]); }
    float sum = 0.f;
    for (int c = 0; c < 4; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    int t = (int)target[i];
    float pt = v[t] / sum;
    float w = seg_w(t);
    acc[0] += w;
    acc[1] += -w * logf(pt);
    for (int c = 1; c < 4; ++c) {
      float p = v[c] / sum, g = (t == c) ? 1.f : 0.f;
      acc[2 + c - 1] += p * g;
      acc[5 + c - 1] += p * p;
      acc[8 + c - 1] += g;
    }
  }
  for (int k = 0; k < 11; ++k) {
    float s = block_sum(acc[k], red);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * 11 + k] = s;
  }
}
__global__ void k_seg_finalize(float* __restrict__ partial, int blocks, float* __restrict__ loss) {
  if (threadIdx.x != 0) return;
  float tot[11];
  for (int k = 0; k < 11; ++k) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * 11 + k];
    tot[k] = s;
  }
  float ce = tot[1] / tot[0];
  float dice = 0.f;
  for (int c = 0; c < 3; ++c) dice += 1.f - 2.f * tot[2 + c] / (tot[5 + c] + tot[8 + c] + 1e-6f);
  loss[0] = ce + dice / 3.f;
  float* totals = partial + (int64_t)blocks * 11;   // kept for the backward pass
  for (int k = 0; k < 11; ++k) totals[k] = tot[k];
}
template <typename T>
__global__ void k_seg_bwd(const T* __restrict__ y, const float* __restrict__ target, const float* __restrict__ totals,
                          const float* __restrict__ upstream, T* __restrict__ dy, int64_t pixels) {
  float up = upstream[0];
  float inv_w = 1.f / totals[0];
  float Dc[4];
  Dc[0] = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    float mx = -INFINITY;
    for (int c = 0; c < 4; ++c) { v[c] = ldf<T>(y + i * 4 + c); mx = fmaxf(mx, v[c]); }
    float sum = 0.f;
    for (int c = 0; c < 4; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    for (int c = 0; c < 4; ++c) v[c] /= sum;
    int t = (int)target[i];
    float w = seg_w(t);
    // dDice/dp_c = -(1/3) * (2 g_c (den_c) - num_c * 2 p_c) / den_c^2 with num_c = 2*sum(p g), den_c = sum(p^2+g)+eps
    float sdp = 0.f;
    for (int c = 1; c < 4; ++c) {
      float num = 2.f * totals[2 + c - 1];
      float den = totals[5 + c - 1] + totals[8 + c - 1] + 1e-6f;
      float g = (t == c) ? 1.f : 0.f;
      Dc[c] = -(1.f / 3.f) * (2.f * g * den - num * 2.f * v[c]) / (den * den);
      sdp += Dc[c] * v[c];
    }
    for (int k = 0; k < 4; ++k) {
      float gce = w * inv_w * (v[k] - ((k == t) ? 1.f : 0.f));
      float gd = v[k] * (Dc[k] - sdp);
      stf<T>(dy + i * 4 + k, up * (gce + gd));
    }
  }
}
extern "C" int rd_seg_loss_fwd(rd_ctx* ctx, const void* y, const float* target, float* loss, float* partial, int N, int64_t hw,
                               int dtype, rd_stream st) {
  int64_t pixels = (int64_t)N * hw;
  int blocks = 256;
  cudaStream_t s = (cudaStream_t)st;
  RD_DISPATCH_DTYPE(dtype, k_seg_partial<T><<<blocks, 256, 0, s>>>((const T*)y, target, partial, pixels));
  RD_CHECK_LAUNCH(ctx, "seg_partial");
  k_seg_finalize<<<1, 32, 0, s>>>(partial, blocks, loss);
  RD_CHECK_LAUNCH(ctx, "seg_finalize");
  return RD_OK;
}
extern "C" int rd_seg_loss_bwd(rd_ctx* ctx, const void* y, const float* target, const float* partial, const float* upstream,
                               void* dy, int N, int64_t hw, int dtype, rd_stream st) {
  int64_t pixels = (int64_t)N * hw;
  const float* totals = partial + (int64_t)256 * 11;
  RD_DISPATCH_DTYPE(dtype, k_seg_bwd<T><<<rd_grid_1d(pixels, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)y, target, totals, upstream, (T*)dy, pixels));
  RD_CHECK_LAUNCH(ctx, "seg_bwd");
  return RD_OK;
}

// ============================================================================ optimizer
// segments: int64 [nseg][2] (offset, length); one block per segment (the host splits parameters into
// segments of at most 64 Ki elements so the grid covers all SMs several times).
__global__ void k_grad_sqsum(const float* __restrict__ grad, const int64_t* __restrict__ seg, float* __restrict__ partial) {
  __shared__ float red[32];
  int64_t off = seg[2 * blockIdx.x], len = seg[2 * blockIdx.x + 1];
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < len; e += blockDim.x) { float g = grad[off + e]; acc += g * g; }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void k_grad_norm_finalize(const float* __restrict__ partial, int nseg, float* __restrict__ scalars, float max_norm) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int k = threadIdx.x; k < nseg; k += blockDim.x) acc += partial[k];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    float total = sqrtf(acc);
    float coef = max_norm / (total + 1e-6f);
    scalars[0] = total;
    scalars[1] = coef < 1.f ? coef : 1.f;          // torch.clamp(max_norm / (total + 1e-6), max=1.0)
    scalars[2] = isfinite(total) ? 1.f : 0.f;      // replaces the per-parameter isfinite scan (main_missing.py:273-278)
  }
}
extern "C" int rd_grad_norm(rd_ctx* ctx, const float* grad, const int64_t* segments, int nseg, float* partial, float* scalars,
                            float max_norm, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  k_grad_sqsum<<<nseg, 256, 0, s>>>(grad, segments, partial);
  RD_CHECK_LAUNCH(ctx, "grad_sqsum");
  k_grad_norm_finalize<<<1, 1024, 0, s>>>(partial, nseg, scalars, max_norm);
  RD_CHECK_LAUNCH(ctx, "grad_norm_finalize");
  return RD_OK;
}
__global__ void k_grad_scale(float* __restrict__ grad, const int64_t* __restrict__ seg, const float* __restrict__ scalars) {
  float coef = scalars[1];
  if (coef >= 1.f) return;
  int64_t off = seg[2 * blockIdx.x], len = seg[2 * blockIdx.x + 1];
  for (int64_t e = threadIdx.x; e < len; e += blockDim.x) grad[off + e] *= coef;
}
extern "C" int rd_grad_scale(rd_ctx* ctx, float* grad, const int64_t* segments, int nseg, const float* scalars, rd_stream st) {
  k_grad_scale<<<nseg, 256, 0, (cudaStream_t)st>>>(grad, segments, scalars);
  RD_CHECK_LAUNCH(ctx, "grad_scale");
  return RD_OK;
}
// hyper[6], hyper[7] = (1 - beta1), (1 - beta2) rounded from DOUBLE like torch.optim.Adam computes them (`value=1 - beta2` is a Python
// float: fl32(0.001), whereas 1.f - fl32(0.999) = 0.00100004673 would bias exp_avg_sq by 4.7e-5 relative); 0 = derive in fp32
struct AdamBetas { float b1, b2, omb1, omb2; };
__device__ __forceinline__ AdamBetas adam_betas(const float* hyper) {
  AdamBetas a;
  a.b1 = hyper[1]; a.b2 = hyper[2];
  a.omb1 = hyper[6] != 0.f ? hyper[6] : 1.f - a.b1;
  a.omb2 = hyper[7] != 0.f ? hyper[7] : 1.f - a.b2;
  return a;
}
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float& vm, float coef, const AdamBetas& ab, float eps, float wd,
                                         float step_size, float inv_sqrt_bc2) {
  // explicit roundings: the fused and the three-launch paths must agree bit for bit whatever the compiler would contract
  const float b1 = ab.b1, b2 = ab.b2;
  const float gs = coef < 1.f ? __fmul_rn(g, coef) : g;
  const float gg = __fmaf_rn(wd, p, gs);
  const float mi = __fmaf_rn(b1, m, __fmul_rn(ab.omb1, gg));
  const float vi = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(ab.omb2, gg), gg));
  const float vx = fmaxf(vm, vi);
  m = mi; v = vi; vm = vx;
  const float denom = __fadd_rn(__fmul_rn(sqrtf(vx), inv_sqrt_bc2), eps);
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(step_size, mi), denom));
  g = 0.f;
}
__global__ void k_adam(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ vmax, const int64_t* __restrict__ seg, const float* __restrict__ hyper) {
  float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const AdamBetas ab = adam_betas(hyper);
  float step = hyper[5] + 1.f;
  float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  int64_t off = seg[2 * blockIdx.x], len = seg[2 * blockIdx.x + 1];
  for (int64_t e = threadIdx.x; e < len; e += blockDim.x) {
    int64_t i = off + e;
    float p = param[i], g = grad[i], mm = m[i], vv = v[i], xx = vmax[i];
    adam_one(p, g, mm, vv, xx, 1.f, ab, eps, wd, step_size, inv_sqrt_bc2);
    param[i] = p; m[i] = mm; v[i] = vv; vmax[i] = xx;
  }
}
// Fused clip-scale + Adam + gradient reset: g = grad * coef (the multiply k_grad_scale would have stored, bit for bit), the Adam
// update, and grad = 0 for the next iteration — one pass of 40 B per parameter instead of three launches / 52 B.  float4 when the
// segment is 16-byte aligned (the large tensors), scalar tail otherwise.
__global__ void __launch_bounds__(256) k_clip_adam(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                                   float* __restrict__ vmax, const int64_t* __restrict__ seg, const float* __restrict__ hyper,
                                                   const float* __restrict__ scalars, int zero_grad) {
  float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const AdamBetas ab = adam_betas(hyper);
  float step = hyper[5] + 1.f;
  float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float coef = scalars ? scalars[1] : 1.f;
  int64_t off = seg[2 * blockIdx.x], len = seg[2 * blockIdx.x + 1];
  int64_t e0 = 0;
  if ((off & 3) == 0) {
    const int64_t nv = len >> 2;
    float4* p4 = reinterpret_cast<float4*>(param + off); float4* g4 = reinterpret_cast<float4*>(grad + off);
    float4* m4 = reinterpret_cast<float4*>(m + off); float4* v4 = reinterpret_cast<float4*>(v + off); float4* x4 = reinterpret_cast<float4*>(vmax + off);
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
      float4 p = p4[i], g = g4[i], mm = m4[i], vv = v4[i], xx = x4[i];
      adam_one(p.x, g.x, mm.x, vv.x, xx.x, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.y, g.y, mm.y, vv.y, xx.y, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.z, g.z, mm.z, vv.z, xx.z, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.w, g.w, mm.w, vv.w, xx.w, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      p4[i] = p; m4[i] = mm; v4[i] = vv; x4[i] = xx;
      if (zero_grad) g4[i] = g;
    }
    e0 = nv << 2;
  }
  for (int64_t e = e0 + threadIdx.x; e < len; e += blockDim.x) {
    int64_t i = off + e;
    float p = param[i], g = grad[i], mm = m[i], vv = v[i], xx = vmax[i];
    adam_one(p, g, mm, vv, xx, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
    param[i] = p; m[i] = mm; v[i] = vv; vmax[i] = xx;
    if (zero_grad) grad[i] = 0.f;
  }
}
__global__ void k_adam_tick(float* hyper) { hyper[5] += 1.f; }

// ---- "grad is None" semantics of torch.optim.Adam (src/main_missing.py:118, 282-284).  torch skips a parameter whose .grad is None —
// no update, no moment decay, no weight decay, and its own `state['step']` does not advance.  In the reference that happens per
// ITERATION WINDOW for modules the masked loss terms never reach (e.g. the private decoder half of a contrast that is missing in every
// row and whose x_mix slots are not read).  Here a parameter is skipped when every one of its gradient segments is
// exactly zero (kernels accumulate exact zeros into unreached modules), and every parameter carries its own step counter.
__global__ void k_param_flags(const float* __restrict__ partial, const int32_t* __restrict__ seg_param, int nseg, int32_t* __restrict__ flags,
                              int nparams) {
  for (int i = threadIdx.x; i < nparams; i += blockDim.x) flags[i] = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < nseg; k += blockDim.x)
    if (partial[k] != 0.f) flags[seg_param[k]] = 1;
}
__global__ void __launch_bounds__(256) k_clip_adam_gated(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m,
                                                         float* __restrict__ v, float* __restrict__ vmax, const int64_t* __restrict__ seg,
                                                         const int32_t* __restrict__ seg_param, const int32_t* __restrict__ flags,
                                                         const float* __restrict__ param_steps, const float* __restrict__ hyper,
                                                         const float* __restrict__ scalars, int zero_grad) {
  const int pid = seg_param[blockIdx.x];
  if (!flags[pid]) return;                                  // grad None: the optimizer does not touch this parameter
  float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const AdamBetas ab = adam_betas(hyper);
  float step = param_steps[pid] + 1.f;
  float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const float coef = scalars ? scalars[1] : 1.f;
  int64_t off = seg[2 * blockIdx.x], len = seg[2 * blockIdx.x + 1];
  int64_t e0 = 0;
  if ((off & 3) == 0) {
    const int64_t nv = len >> 2;
    float4* p4 = reinterpret_cast<float4*>(param + off); float4* g4 = reinterpret_cast<float4*>(grad + off);
    float4* m4 = reinterpret_cast<float4*>(m + off); float4* v4 = reinterpret_cast<float4*>(v + off); float4* x4 = reinterpret_cast<float4*>(vmax + off);
    for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
      float4 p = p4[i], g = g4[i], mm = m4[i], vv = v4[i], xx = x4[i];
      adam_one(p.x, g.x, mm.x, vv.x, xx.x, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.y, g.y, mm.y, vv.y, xx.y, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.z, g.z, mm.z, vv.z, xx.z, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      adam_one(p.w, g.w, mm.w, vv.w, xx.w, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
      p4[i] = p; m4[i] = mm; v4[i] = vv; x4[i] = xx;
      if (zero_grad) g4[i] = g;
    }
    e0 = nv << 2;
  }
  for (int64_t e = e0 + threadIdx.x; e < len; e += blockDim.x) {
    int64_t i = off + e;
    float p = param[i], g = grad[i], mm = m[i], vv = v[i], xx = vmax[i];
    adam_one(p, g, mm, vv, xx, coef, ab, eps, wd, step_size, inv_sqrt_bc2);
    param[i] = p; m[i] = mm; v[i] = vv; vmax[i] = xx;
    if (zero_grad) grad[i] = 0.f;
  }
}
__global__ void k_adam_tick_gated(float* hyper, const int32_t* __restrict__ flags, float* __restrict__ param_steps, int nparams) {
  for (int i = threadIdx.x; i < nparams; i += blockDim.x)
    if (flags[i]) param_steps[i] += 1.f;
  if (threadIdx.x == 0) hyper[5] += 1.f;
}
extern "C" int rd_clip_adam_amsgrad_gated(rd_ctx* ctx, float* param, float* grad, float* m, float* v, float* vmax, const int64_t* segments,
                                          const int32_t* seg_param, int nseg, const float* partial, int32_t* param_flags, float* param_steps,
                                          int nparams, float* hyper, const float* scalars, int zero_grad, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  if (nseg < 1) return RD_OK;
  k_param_flags<<<1, 1024, 0, s>>>(partial, seg_param, nseg, param_flags, nparams);
  RD_CHECK_LAUNCH(ctx, "param_flags");
  k_clip_adam_gated<<<nseg, 256, 0, s>>>(param, grad, m, v, vmax, segments, seg_param, param_flags, param_steps, hyper, scalars, zero_grad);
  RD_CHECK_LAUNCH(ctx, "clip_adam_amsgrad_gated");
  k_adam_tick_gated<<<1, 1024, 0, s>>>(hyper, param_flags, param_steps, nparams);
  RD_CHECK_LAUNCH(ctx, "adam_tick_gated");
  return RD_OK;
}
extern "C" int rd_clip_adam_amsgrad(rd_ctx* ctx, float* param, float* grad, float* m, float* v, float* vmax, const int64_t* segments,
                                    int nseg, float* hyper, const float* scalars, int zero_grad, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  if (nseg < 1) return RD_OK;
  k_clip_adam<<<nseg, 256, 0, s>>>(param, grad, m, v, vmax, segments, hyper, scalars, zero_grad);
  RD_CHECK_LAUNCH(ctx, "clip_adam_amsgrad");
  k_adam_tick<<<1, 1, 0, s>>>(hyper);
  RD_CHECK_LAUNCH(ctx, "adam_tick");
  return RD_OK;
}
extern "C" int rd_adam_amsgrad(rd_ctx* ctx, float* param, const float* grad, float* m, float* v, float* vmax,
                               const int64_t* segments, int nseg, float* hyper, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  k_adam<<<nseg, 256, 0, s>>>(param, grad, m, v, vmax, segments, hyper);
  RD_CHECK_LAUNCH(ctx, "adam_amsgrad");
  k_adam_tick<<<1, 1, 0, s>>>(hyper);
  RD_CHECK_LAUNCH(ctx, "adam_tick");
  return RD_OK;
}
