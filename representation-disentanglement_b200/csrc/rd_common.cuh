// Shared device / host helpers for the rd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/rd_b200.h"

struct rd_ctx {
  int device;
  int sm_count;
  int max_smem_optin;
  std::atomic<int64_t> launches;
  int last_conv_algo;
  char err[512];
  bool tc_attr_set;
  void* nccl_comm;          // rd_ddp_init (rd_runtime.cu)
  int ddp_world, ddp_rank;
  int* spf_sync;            // kSpfSlots zero-initialised counter slots of k_spade_bwd_fused (device memory, rd_ctx_create)
  unsigned spf_next;        // next slot (round-robin per launch)
};
constexpr int kSpfMaxN = 4096;                     // images per launch the counter slots are sized for
constexpr int kSpfSlotInts = 2 + 2 * kSpfMaxN;     // ticket, done, arrive[N], ready[N]
constexpr int kSpfSlots = 8;

#define RD_FAIL(ctx, code, ...)                                   \
  do {                                                            \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return (code);                                                \
  } while (0)

#define RD_CHECK_LAUNCH(ctx, name)                                              \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess)                                                     \
      RD_FAIL(ctx, RD_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                    \
  } while (0)

#define RD_CUDA(ctx, call)                                                      \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess)                                                     \
      RD_FAIL(ctx, RD_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- dtype-generic load / store
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4-wide vector access (16 B for fp32, 8 B for bf16); pointers must be aligned accordingly.
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// widest 16-byte vector per dtype: 8 bf16 or 4 fp32 channels
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { Vec4<float>::load(p, v); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { Vec4<float>::store(p, v); }
};
template <> struct VecIO<bf16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __low2float(h[k]); v[2 * k + 1] = __high2float(h[k]); }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; `red` >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}

static inline int rd_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// RD_B200_TRACE_CONV=1: one stderr line per convolution launch (operation, kernel chosen, shape) — how the launch lists
// in profiles/ are mapped back to layers
static inline void rd_trace_conv(const char* op, const char* kernel, const rd_conv_desc* d) {
  static const bool on = getenv("RD_B200_TRACE_CONV") != nullptr;
  if (on)
    fprintf(stderr, "rd_conv %s %s n=%d g=%d %dx%d cin=%d cout=%d k=%d s=%d p=%d bg=%d\n", op, kernel, d->n, d->groups, d->h, d->w, d->cin,
            d->cout, d->kh, d->stride, d->pad, d->bias_groups);
}
static inline int rd_grid_1d(int64_t n, int block, int sm_count) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count * 16;   // grid-stride loops; a few waves of resident CTAs per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

#define RD_DISPATCH_DTYPE(dtype, ...)                       \
  if ((dtype) == RD_F32) { typedef float T; __VA_ARGS__; }  \
  else if ((dtype) == RD_BF16) { typedef bf16 T; __VA_ARGS__; } \
  else RD_FAIL(ctx, RD_ERR_ARG, "bad dtype %d", (int)(dtype));

// implemented in rd_conv_tc.cu
int rd_conv_tc_supported(const rd_conv_desc* d, int mode);
int rd_conv_tc_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias,
                      void* y, cudaStream_t st);
int rd_wgrad_tc_supported(const rd_conv_desc* d);
int rd_wgrad_tc_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                       cudaStream_t st);
// implemented in rd_conv_tma.cu (persistent TMA-fed kernel for the stride-1 "same" convolutions)
int rd_conv_tma_supported(const rd_conv_desc* d, int mode);
int rd_conv_tma_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias,
                       void* y, cudaStream_t st);
int rd_wgrad_tma_supported(const rd_conv_desc* d);
int rd_wgrad_tma_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                        cudaStream_t st);
// implemented in rd_conv_halo.cu (halo-tile kernel with shared-memory-resident weights for the 3x3 stride-1 layers)
int rd_conv_halo_supported(const rd_conv_desc* d, int mode, int sm_count, int forced);
int rd_conv_halo_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias,
                        void* y, cudaStream_t st);
int rd_conv_halo_spade_supported(const rd_conv_desc* d, int sm_count);
int rd_conv_halo_spade_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* w, const float* bias, const void* z,
                              const float* mean, const float* invstd, void* gamma, void* mix, cudaStream_t st);
int rd_wgrad_halo_supported(const rd_conv_desc* d, int sm_count);
int rd_wgrad_halo_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                         cudaStream_t st);
// cuTensorMapEncodeTiled through the runtime's driver entry point (NULL when unavailable); rd_conv_tma.cu
void* rd_tensormap_encode_fn();
