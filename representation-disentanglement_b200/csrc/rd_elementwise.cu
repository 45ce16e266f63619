// rd_elementwise.cu — context, layout/cast, CondConv expert mixing, normalisation (BatchNorm /
// InstanceNorm / SPADE modulation), bilinear resize, activations, masked softmax, small linears.
// All activation tensors are NHWC; these kernels are HBM-bound: coalesced channel-innermost access,
// 16-byte vectors where the channel count allows, warp-shuffle + shared-memory reductions for the
// per-(group, channel) statistics, deterministic two-level reductions (partials -> finalize).
#include "rd_common.cuh"

// ============================================================================ context
extern "C" int rd_abi_version(void) { return 1; }

extern "C" int rd_ctx_create(rd_ctx** out, int device) {
  if (!out) return RD_ERR_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return RD_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return RD_ERR_ARG;
  rd_ctx* c = new rd_ctx();
  c->device = device;
  c->launches = 0;
  c->last_conv_algo = 0;
  c->err[0] = 0;
  c->tc_attr_set = false;
  c->nccl_comm = nullptr; c->ddp_world = 1; c->ddp_rank = 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { delete c; return RD_ERR_CUDA; }
  c->sm_count = p.multiProcessorCount;
  // RD_B200_SM_COUNT: size the persistent kernels' grids for fewer SMs than the device has.  With data parallelism the NCCL kernels of the
  // overlapped bucket all-reduces hold a few SMs; a persistent kernel launched with one CTA per SM then cannot make all its CTAs resident,
  // the left-over CTAs run as a second wave and the kernel takes twice as long while the collective is in flight.
  if (const char* e = getenv("RD_B200_SM_COUNT")) {
    int v = atoi(e);
    if (v >= 8 && v < c->sm_count) c->sm_count = v;
  }
  c->max_smem_optin = (int)p.sharedMemPerBlockOptin;
  if (p.major != 10) {
    snprintf(c->err, sizeof(c->err), "rd_b200 needs an sm_100 device, found sm_%d%d", p.major, p.minor);
    *out = c;
    return RD_ERR_UNSUPPORTED;
  }
  // zeroed counter slots of the single-pass SPADE backward (k_spade_bwd_fused resets its slot itself at the end of every launch)
  c->spf_sync = nullptr; c->spf_next = 0;
  {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    if (cudaMalloc(&c->spf_sync, sizeof(int) * (size_t)kSpfSlots * kSpfSlotInts) == cudaSuccess)
      cudaMemset(c->spf_sync, 0, sizeof(int) * (size_t)kSpfSlots * kSpfSlotInts);
    else { c->spf_sync = nullptr; cudaGetLastError(); }
    cudaSetDevice(cur);
  }
  *out = c;
  return RD_OK;
}
extern "C" int rd_ctx_destroy(rd_ctx* ctx) {
  if (ctx && ctx->spf_sync) cudaFree(ctx->spf_sync);
  delete ctx;
  return RD_OK;
}
extern "C" const char* rd_last_error(rd_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" int64_t rd_launch_count(rd_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
extern "C" int rd_last_conv_algo(rd_ctx* ctx) { return ctx ? ctx->last_conv_algo : 0; }

// ============================================================================ layout / cast
template <typename T>
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, T* __restrict__ dst, int n, int c_total, int c0, int c,
                               int64_t hw) {
  int64_t total = (int64_t)n * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t img = i / hw, p = i - img * hw;
    const float* s = src + (img * c_total + c0) * hw + p;
    T* d = dst + i * c;
    for (int k = 0; k < c; ++k) stf<T>(d + k, s[(int64_t)k * hw]);
  }
}
// many channels, few pixels (the (B, 128, 5, 6) code out of zi_scaler): one thread per ELEMENT — a thread per pixel walking 128
// channels left 30 blocks looping for 41 us
template <typename T>
__global__ void k_nchw_to_nhwc_elem(const float* __restrict__ src, T* __restrict__ dst, int c_total, int c0, int c, int64_t hw, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / c;
    const int k = (int)(i - pix * c);
    const int64_t img = pix / hw, p = pix - img * hw;
    stf<T>(dst + i, src[(img * c_total + c0 + k) * hw + p]);
  }
}
template <typename T>
__global__ void k_nhwc_to_nchw(const T* __restrict__ src, float* __restrict__ dst, int n, int c, int64_t hw) {
  int64_t total = (int64_t)n * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t img = i / hw, p = i - img * hw;
    const T* s = src + i * c;
    float* d = dst + img * c * hw + p;
    for (int k = 0; k < c; ++k) d[(int64_t)k * hw] = ldf<T>(s + k);
  }
}
extern "C" int rd_nchw_to_nhwc(rd_ctx* ctx, const float* src, void* dst, int n, int c_total, int c0, int c, int h,
                               int w, int dtype, rd_stream st) {
  int64_t hw = (int64_t)h * w;
  if (c >= 16 && (int64_t)n * hw < 256 * (int64_t)ctx->sm_count * 4) {
    const int64_t total = (int64_t)n * hw * c;
    RD_DISPATCH_DTYPE(dtype, (k_nchw_to_nhwc_elem<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(src, (T*)dst, c_total, c0, c, hw, total)));
    RD_CHECK_LAUNCH(ctx, "nchw_to_nhwc_elem");
    return RD_OK;
  }
  int grid = rd_grid_1d((int64_t)n * hw, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_nchw_to_nhwc<T><<<grid, 256, 0, (cudaStream_t)st>>>(src, (T*)dst, n, c_total, c0, c, hw)));
  RD_CHECK_LAUNCH(ctx, "nchw_to_nhwc");
  return RD_OK;
}
// dst[(m * n + b), p, 0:c] = src[b, m * c : (m + 1) * c, p]: the modality-major NHWC stack of a (n, mods * c, h, w) batch in ONE launch
template <typename T>
__global__ void k_stack_modalities(const float* __restrict__ src, T* __restrict__ dst, int n, int mods, int c, int c_pad, int64_t hw) {
  const int64_t total = (int64_t)mods * n * hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / hw, p = i - j * hw;
    const int m = (int)(j / n), b = (int)(j - (int64_t)m * n);
    const float* s = src + ((int64_t)b * mods * c + (int64_t)m * c) * hw + p;
    T* d = dst + i * c_pad;
    if (sizeof(T) == 2 && c <= 8 && c_pad == 16) {            // 7-channel slabs straight into the zero-padded 16-channel tensor-core layout
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < c) {
          const bf16 v = __float2bfloat16_rn(s[(int64_t)k * hw]);
          w[k >> 1] |= (uint32_t)(*reinterpret_cast<const uint16_t*>(&v)) << ((k & 1) * 16);
        }
      uint4* o4 = reinterpret_cast<uint4*>(d);
      o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o4[1] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int k = 0; k < c_pad; ++k) stf<T>(d + k, k < c ? s[(int64_t)k * hw] : 0.f);
    }
  }
}
extern "C" int rd_stack_modalities(rd_ctx* ctx, const float* src, void* dst, int n, int mods, int c, int c_pad, int h, int w, int dtype, rd_stream st) {
  const int64_t hw = (int64_t)h * w;
  if (c_pad < c) RD_FAIL(ctx, RD_ERR_ARG, "stack_modalities: c_pad < c");
  RD_DISPATCH_DTYPE(dtype, (k_stack_modalities<T><<<rd_grid_1d((int64_t)mods * n * hw, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(src, (T*)dst, n, mods, c, c_pad, hw)));
  RD_CHECK_LAUNCH(ctx, "stack_modalities");
  return RD_OK;
}
extern "C" int rd_nhwc_to_nchw(rd_ctx* ctx, const void* src, float* dst, int n, int c, int h, int w, int dtype,
                               rd_stream st) {
  int64_t hw = (int64_t)h * w;
  int grid = rd_grid_1d((int64_t)n * hw, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_nhwc_to_nchw<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)src, dst, n, c, hw)));
  RD_CHECK_LAUNCH(ctx, "nhwc_to_nchw");
  return RD_OK;
}

template <typename S, typename D>
__global__ void k_cast(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf<D>(d + i, ldf<S>(s + i));
}
extern "C" int rd_cast(rd_ctx* ctx, const void* src, int sd, void* dst, int dd, int64_t n, rd_stream st) {
  int grid = rd_grid_1d(n, 256, ctx->sm_count);
  cudaStream_t s = (cudaStream_t)st;
  if (sd == RD_F32 && dd == RD_BF16) k_cast<float, bf16><<<grid, 256, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (sd == RD_BF16 && dd == RD_F32) k_cast<bf16, float><<<grid, 256, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (sd == RD_F32 && dd == RD_F32) k_cast<float, float><<<grid, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else if (sd == RD_BF16 && dd == RD_BF16) k_cast<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  else RD_FAIL(ctx, RD_ERR_ARG, "rd_cast: bad dtypes");
  RD_CHECK_LAUNCH(ctx, "cast");
  return RD_OK;
}

template <typename T>
__global__ void k_concat(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t pixels, int ca,
                         int cb) {
  int ct = ca + cb;
  int64_t total = pixels * ct;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / ct;
    int c = (int)(i - p * ct);
    out[i] = (c < ca) ? a[p * ca + c] : b[p * cb + (c - ca)];
  }
}
template <typename T>
__global__ void k_split(const T* __restrict__ in, T* __restrict__ a, T* __restrict__ b, int64_t pixels, int ca, int cb) {
  int ct = ca + cb;
  int64_t total = pixels * ct;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / ct;
    int c = (int)(i - p * ct);
    if (c < ca) { if (a) a[p * ca + c] = in[i]; }
    else { if (b) b[p * cb + (c - ca)] = in[i]; }
  }
}
// 16-byte vector versions (channel counts that are multiples of 8 bf16 / 4 fp32 — every concat of the anatomy decoder): plain uint4
// moves, one 32-bit division per vector instead of a 64-bit one per 2-byte element (the scalar kernels ran at 1/4 of the HBM rate)
template <typename T>
__global__ void __launch_bounds__(256) k_concat_vec(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t pixels,
                                                    int ca, int cb) {
  constexpr int V = 16 / (int)sizeof(T);
  const int va = ca / V, vt = (ca + cb) / V;
  const int64_t total = pixels * vt;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / vt;
    const int c = (int)(i - p * vt);
    const uint4 v = (c < va) ? __ldg(reinterpret_cast<const uint4*>(a + p * ca) + c) : __ldg(reinterpret_cast<const uint4*>(b + p * cb) + (c - va));
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) k_split_vec(const T* __restrict__ in, T* __restrict__ a, T* __restrict__ b, int64_t pixels, int ca, int cb) {
  constexpr int V = 16 / (int)sizeof(T);
  const int va = ca / V, vt = (ca + cb) / V;
  const int64_t total = pixels * vt;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / vt;
    const int c = (int)(i - p * vt);
    if (c < va) { if (a) reinterpret_cast<uint4*>(a + p * ca)[c] = __ldg(reinterpret_cast<const uint4*>(in) + i); }
    else { if (b) reinterpret_cast<uint4*>(b + p * cb)[c - va] = __ldg(reinterpret_cast<const uint4*>(in) + i); }
  }
}
static inline bool rd_vec16_ok(const void* a, const void* b, const void* c, int ca, int cb, int dtype) {
  const int V = dtype == RD_F32 ? 4 : 8;
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return ca % V == 0 && cb % V == 0 && al(a) && al(b) && al(c);
}
extern "C" int rd_concat_channels(rd_ctx* ctx, const void* a, const void* b, void* out, int64_t pixels, int ca, int cb,
                                  int dtype, rd_stream st) {
  if (rd_vec16_ok(a, b, out, ca, cb, dtype)) {
    const int V = dtype == RD_F32 ? 4 : 8;
    int gridv = rd_grid_1d(pixels * ((ca + cb) / V), 256, ctx->sm_count);
    RD_DISPATCH_DTYPE(dtype, (k_concat_vec<T><<<gridv, 256, 0, (cudaStream_t)st>>>((const T*)a, (const T*)b, (T*)out, pixels, ca, cb)));
    RD_CHECK_LAUNCH(ctx, "concat_vec");
    return RD_OK;
  }
  int grid = rd_grid_1d(pixels * (ca + cb), 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_concat<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)a, (const T*)b, (T*)out, pixels, ca, cb)));
  RD_CHECK_LAUNCH(ctx, "concat");
  return RD_OK;
}
extern "C" int rd_split_channels(rd_ctx* ctx, const void* in, void* a, void* b, int64_t pixels, int ca, int cb, int dtype,
                                 rd_stream st) {
  if (rd_vec16_ok(a, b, in, ca, cb, dtype)) {
    const int V = dtype == RD_F32 ? 4 : 8;
    int gridv = rd_grid_1d(pixels * ((ca + cb) / V), 256, ctx->sm_count);
    RD_DISPATCH_DTYPE(dtype, (k_split_vec<T><<<gridv, 256, 0, (cudaStream_t)st>>>((const T*)in, (T*)a, (T*)b, pixels, ca, cb)));
    RD_CHECK_LAUNCH(ctx, "split_vec");
    return RD_OK;
  }
  int grid = rd_grid_1d(pixels * (ca + cb), 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_split<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)in, (T*)a, (T*)b, pixels, ca, cb)));
  RD_CHECK_LAUNCH(ctx, "split");
  return RD_OK;
}
template <typename T>
__global__ void k_add(const T* __restrict__ x, const T* __restrict__ a, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(y + i, ldf<T>(x + i) + ldf<T>(a + i));
}
// y = sum of up to 8 tensors (fp32 accumulation): the gradient of a tensor with several consumers in ONE pass (ops.fanout) instead of the
// chain of at::add launches autograd's input buffer would issue
struct AddNPtrs { const void* p[8]; };
template <typename T>
__global__ void k_add_n(AddNPtrs in, int k, T* __restrict__ y, int64_t n) {
  constexpr int V = VecIO<T>::V;
  const int64_t nv = n / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) acc[q] = 0.f;
    for (int j = 0; j < k; ++j) {
      float v[V];
      VecIO<T>::load((const T*)in.p[j] + i * V, v);
#pragma unroll
      for (int q = 0; q < V; ++q) acc[q] += v[q];
    }
    VecIO<T>::store(y + i * V, acc);
  }
  for (int64_t i = nv * V + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < k; ++j) a += ldf<T>((const T*)in.p[j] + i);
    stf<T>(y + i, a);
  }
}
extern "C" int rd_add_n(rd_ctx* ctx, const void* const* xs, int k, void* y, int64_t n, int dtype, rd_stream st) {
  if (k < 1 || k > 8) RD_FAIL(ctx, RD_ERR_ARG, "add_n: 1..8 inputs");
  AddNPtrs in;
  for (int j = 0; j < 8; ++j) in.p[j] = j < k ? xs[j] : nullptr;
  for (int j = 0; j < k; ++j)
    if (((uintptr_t)in.p[j] & 15u) != 0) RD_FAIL(ctx, RD_ERR_ARG, "add_n: inputs must be 16-byte aligned");
  RD_DISPATCH_DTYPE(dtype, (k_add_n<T><<<rd_grid_1d(n / 4 + 1, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(in, k, (T*)y, n)));
  RD_CHECK_LAUNCH(ctx, "add_n");
  return RD_OK;
}
extern "C" int rd_add(rd_ctx* ctx, const void* x, const void* a, void* y, int64_t n, int dtype, rd_stream st) {
  int grid = rd_grid_1d(n, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_add<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)x, (const T*)a, (T*)y, n)));
  RD_CHECK_LAUNCH(ctx, "add");
  return RD_OK;
}

// ============================================================================ CondConv expert mixing
struct MixTypes { float t[16]; };

// Two index spaces per launch so that BOTH outputs are written with the channel index fastest (coalesced 2-byte stores;
// the outputs are G x larger than the fp32 expert weights that are read): first `per` work items in packed order
// (o, tap, i), then `per` items in packedT order (i, tap, o).  One item mixes all G groups from 3 weight reads; the G x E
// routing sigmoids are computed once per block.
template <typename T>
__global__ void __launch_bounds__(256) k_mix_fwd(const float* __restrict__ W, const float* __restrict__ fcw, const float* __restrict__ fcb,
                          MixTypes types, int G, int E, int O, int I, int i_pad, int taps, int o_total, int oT_total,
                          int o_off, T* __restrict__ packed, T* __restrict__ packedT, float* __restrict__ r_out) {
  __shared__ float rs[16 * 3];
  if (threadIdx.x < G * 3) {
    int g = threadIdx.x / 3, e = threadIdx.x % 3;
    rs[threadIdx.x] = e < E ? (fcw ? 1.f / (1.f + expf(-(fcw[e] * types.t[g] + fcb[e]))) : 1.f) : 0.f;
  }
  __syncthreads();
  const int per = O * taps * i_pad;
  const int64_t wexp = (int64_t)O * I * taps;                 // elements per expert
  const int n_items = (packed ? per : 0) + (packedT ? per : 0);
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += gridDim.x * blockDim.x) {
    const bool to_packed = packed && it < per;
    const int idx = to_packed ? it : it - (packed ? per : 0);
    int o, tap, i;
    if (to_packed) { o = idx / (taps * i_pad); int r2 = idx - o * taps * i_pad; tap = r2 / i_pad; i = r2 - tap * i_pad; }
    else { i = idx / (taps * O); int r2 = idx - i * taps * O; tap = r2 / O; o = r2 - tap * O; }
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if (i < I) {
      const int64_t wi = ((int64_t)o * I + i) * taps + tap;   // W layout (E, O, I, kh, kw): tap is the fastest index
      w0 = W[wi];
      if (E > 1) w1 = W[wexp + wi];
      if (E > 2) w2 = W[2 * wexp + wi];
    }
    if (to_packed) {
      T* dst = packed + ((int64_t)(o_off + o) * taps + tap) * i_pad + i;
      const int64_t gs = (int64_t)o_total * taps * i_pad;
      for (int g = 0; g < G; ++g) stf<T>(dst + g * gs, rs[g * 3] * w0 + rs[g * 3 + 1] * w1 + rs[g * 3 + 2] * w2);
    } else {
      T* dst = packedT + ((int64_t)i * taps + tap) * oT_total + o_off + o;
      const int64_t gs = (int64_t)i_pad * taps * oT_total;
      for (int g = 0; g < G; ++g) stf<T>(dst + g * gs, rs[g * 3] * w0 + rs[g * 3 + 1] * w1 + rs[g * 3 + 2] * w2);
    }
  }
  if (r_out && blockIdx.x == 0 && threadIdx.x < G * E) {
    int g = threadIdx.x / E, e = threadIdx.x % E;
    r_out[threadIdx.x] = rs[g * 3 + e];
  }
}
extern "C" int rd_condconv_mix_fwd(rd_ctx* ctx, const float* W, const float* fc_w, const float* fc_b, const float* types,
                                   int G, int E, int O, int I, int i_pad, int kh, int kw, int o_total, int oT_total,
                                   int o_off, void* packed, void* packedT, float* r_out, int dtype, rd_stream st) {
  if (G < 1 || G > 16 || E < 1 || E > 3) RD_FAIL(ctx, RD_ERR_ARG, "mix_fwd: G in [1,16], E in [1,3]");
  if (!fc_w && E != 1) RD_FAIL(ctx, RD_ERR_ARG, "mix_fwd: plain conv weights need E == 1");
  if (i_pad < I) RD_FAIL(ctx, RD_ERR_ARG, "mix_fwd: i_pad < I");
  MixTypes mt;
  for (int g = 0; g < 16; ++g) mt.t[g] = (types && g < G) ? types[g] : 0.f;
  int64_t total = (int64_t)2 * O * i_pad * kh * kw;            // one work item per (output kind, o, tap, i): all G groups
  int grid = rd_grid_1d(total, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_mix_fwd<T><<<grid, 256, 0, (cudaStream_t)st>>>(W, fc_w, fc_b, mt, G, E, O, I, i_pad, kh * kw,
                                                                              o_total, oT_total, o_off, (T*)packed,
                                                                              (T*)packedT, r_out)));
  RD_CHECK_LAUNCH(ctx, "condconv_mix_fwd");
  return RD_OK;
}

// The same mixing (+ the bias row of the fused launch) for EVERY CondConv head of an iteration in one launch: the per-head launches
// are ~6 us each and ~90 per step, and the weights only change at the optimizer step.  Job j owns blocks [block_begin, +blocks).
// Round 2: a block owns UNITS of 16 output channels x 16 input channels x all taps.  The expert weights are (E, O, I, kh, kw) — tap
// fastest — while both packed layouts are channel fastest ((o, tap, i) and (i, tap, o)): with one thread per output element either the W
// reads or the stores touched 32 sectors per warp instruction (the launch ran at 1.2 TB/s).  A unit's W slice is `nol` contiguous runs of
// ic x taps floats: it is read coalesced into shared memory once, and both outputs are written as 16-byte runs of 8 channels
// (8 input channels for `packed`, 8 output channels for `packedT`), one thread mixing all G groups of its run.
constexpr int kMixFNO = 16, kMixFIC = 16;
constexpr int kMixFSlice = kMixFNO * kMixFIC * 17;            // floats per expert: taps <= 16 -> row pitch (taps | 1) <= 17
constexpr int kMixFDynBytes = 3 * kMixFSlice * 4;
extern "C" int rd_mixf_job_blocks(int O, int i_pad, int taps) {
  (void)taps;
  int b = ((O + kMixFNO - 1) / kMixFNO) * ((i_pad + kMixFIC - 1) / kMixFIC);
  return b < 1 ? 1 : b;
}
template <typename T> __device__ __forceinline__ void mixf_store8(T* dst, const float (&v)[8]);
template <> __device__ __forceinline__ void mixf_store8<float>(float* dst, const float (&v)[8]) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void mixf_store8<bf16>(bf16* dst, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(dst) = t;
}
template <typename T>
__global__ void __launch_bounds__(256) k_mix_fwd_batched(const rd_mixf_job* __restrict__ jobs, int njobs) {
  extern __shared__ float mixf_ws[];      // [3 experts][kMixFSlice]: slot (ol * 16 + ii) * tp + tap
  __shared__ float rs[16 * 3];
  __shared__ int job_s;
  if (threadIdx.x == 0) {               // binary search: last job with block_begin <= blockIdx.x
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    job_s = lo;
  }
  __syncthreads();
  const rd_mixf_job& J = jobs[job_s];
  const int G = J.G, E = J.E, O = J.O, I = J.I, i_pad = J.i_pad, taps = J.taps, o_total = J.o_total, oT_total = J.oT_total, o_off = J.o_off;
  const float* __restrict__ W = J.W; const float* __restrict__ fcw = J.fc_w; const float* __restrict__ fcb = J.fc_b;
  T* __restrict__ packed = (T*)J.packed; T* __restrict__ packedT = (T*)J.packedT;
  if (threadIdx.x < G * 3) {
    int g = threadIdx.x / 3, e = threadIdx.x % 3;
    rs[threadIdx.x] = e < E ? (fcw ? 1.f / (1.f + expf(-(fcw[e] * J.types[g] + fcb[e]))) : 1.f) : 0.f;
  }
  __syncthreads();
  const int lb = (int)blockIdx.x - J.block_begin;
  if (lb == 0 && J.bias_dst)
    for (int i = threadIdx.x; i < J.bias_n; i += 256) J.bias_dst[i] = J.bias_src[i];
  const int64_t wexp = (int64_t)O * I * taps;                 // elements per expert
  const int tp = taps | 1;
  const int ichunks = (i_pad + kMixFIC - 1) / kMixFIC;
  const int units = ((O + kMixFNO - 1) / kMixFNO) * ichunks;
  // 16-byte stores need 8-element aligned runs in both packed layouts (true for every layer of the model; otherwise element stores)
  const bool vec_p = (i_pad & 7) == 0 && (reinterpret_cast<uintptr_t>(packed) & 15u) == 0;
  const bool vec_t = (oT_total & 7) == 0 && (o_off & 7) == 0 && (reinterpret_cast<uintptr_t>(packedT) & 15u) == 0;
  const int64_t gs_p = (int64_t)o_total * taps * i_pad, gs_t = (int64_t)i_pad * taps * oT_total;
  for (int u = lb; u < units; u += J.blocks) {
    const int ob = u / ichunks, ib = u - ob * ichunks;
    const int o0 = ob * kMixFNO, i0 = ib * kMixFIC;
    const int nol = O - o0 < kMixFNO ? O - o0 : kMixFNO;
    const int icv = I - i0 < kMixFIC ? (I - i0 > 0 ? I - i0 : 0) : kMixFIC;        // input channels that exist in W (the rest is zero padding)
    const int icp = i_pad - i0 < kMixFIC ? i_pad - i0 : kMixFIC;                    // input channels of the packed layouts in this chunk
    const int run = icv * taps;
    __syncthreads();                                           // the previous unit's reads of mixf_ws are done
    for (int t = threadIdx.x; t < nol * run; t += 256) {
      const int ol = t / run, rem = t - ol * run;
      const int ii = rem / taps, tap = rem - ii * taps;
      const int64_t wi = ((int64_t)(o0 + ol) * I + i0) * taps + rem;
      const int slot = (ol * kMixFIC + ii) * tp + tap;
      for (int e = 0; e < E; ++e) mixf_ws[e * kMixFSlice + slot] = __ldg(W + e * wexp + wi);
    }
    __syncthreads();
    // packed[g][o_off + o][tap][i0 + 8 half ..]: 8 input channels per item
    for (int it = threadIdx.x; it < nol * taps * 2; it += 256) {
      const int half = it & 1, r2 = it >> 1;
      const int ol = r2 / taps, tap = r2 - ol * taps;
      const int ib8 = half * 8;
      if (ib8 >= icp) continue;
      float w[3][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool v = ib8 + k < icv;
        const int slot = (ol * kMixFIC + ib8 + k) * tp + tap;
#pragma unroll
        for (int e = 0; e < 3; ++e) w[e][k] = (v && e < E) ? mixf_ws[e * kMixFSlice + slot] : 0.f;
      }
      T* dst = packed + ((int64_t)(o_off + o0 + ol) * taps + tap) * i_pad + i0 + ib8;
      const bool full = vec_p && ib8 + 8 <= icp;
      for (int g = 0; g < G; ++g) {
        const float r0 = rs[g * 3], r1 = rs[g * 3 + 1], r2w = rs[g * 3 + 2];
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = r0 * w[0][k] + r1 * w[1][k] + r2w * w[2][k];
        if (full) mixf_store8<T>(dst + g * gs_p, m);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) if (ib8 + k < icp) stf<T>(dst + g * gs_p + k, m[k]);
        }
      }
    }
    // packedT[g][i0 + ii][tap][o_off + o0 + 8 half ..]: 8 output channels per item
    for (int it = threadIdx.x; it < icp * taps * 2; it += 256) {
      const int half = it & 1, r2 = it >> 1;
      const int ii = r2 / taps, tap = r2 - ii * taps;
      const int ob8 = half * 8;
      if (ob8 >= nol) continue;
      const bool iv = ii < icv;
      float w[3][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool v = iv && ob8 + k < nol;
        const int slot = ((ob8 + k) * kMixFIC + ii) * tp + tap;
#pragma unroll
        for (int e = 0; e < 3; ++e) w[e][k] = (v && e < E) ? mixf_ws[e * kMixFSlice + slot] : 0.f;
      }
      T* dst = packedT + ((int64_t)(i0 + ii) * taps + tap) * oT_total + o_off + o0 + ob8;
      const bool full = vec_t && ob8 + 8 <= nol;
      for (int g = 0; g < G; ++g) {
        const float r0 = rs[g * 3], r1 = rs[g * 3 + 1], r2w = rs[g * 3 + 2];
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = r0 * w[0][k] + r1 * w[1][k] + r2w * w[2][k];
        if (full) mixf_store8<T>(dst + g * gs_t, m);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) if (ob8 + k < nol) stf<T>(dst + g * gs_t + k, m[k]);
        }
      }
    }
  }
}
extern "C" int rd_condconv_mix_fwd_batched(rd_ctx* ctx, const rd_mixf_job* jobs_dev, int njobs, int total_blocks, int dtype, rd_stream st) {
  if (njobs < 1 || total_blocks < 1) return RD_OK;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_mix_fwd_batched<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixFDynBytes));
    RD_CUDA(ctx, cudaFuncSetAttribute(k_mix_fwd_batched<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixFDynBytes));
    attr_set = true;
  }
  RD_DISPATCH_DTYPE(dtype, (k_mix_fwd_batched<T><<<total_blocks, 256, kMixFDynBytes, (cudaStream_t)st>>>(jobs_dev, njobs)));
  RD_CHECK_LAUNCH(ctx, "condconv_mix_fwd_batched");
  return RD_OK;
}

// One pass over dK: dW[e][o][i][tap] += sum_g r[g,e] dK[g][o][tap][i] and the routing gradients
// dr[g,e] = <dK[g], W[e]>  ->  dfc_w[e] += dr r(1-r) t_g ; dfc_b[e] += dr r(1-r)   (block reduction + atomics).
template <int GM>   // GM = compile-time bound on the number of groups (register-resident accumulators); E <= 3
__global__ void __launch_bounds__(256) k_mix_bwd(const float* __restrict__ dK, const float* __restrict__ W,
                                                  const float* __restrict__ fcw, const float* __restrict__ fcb,
                                                  MixTypes types, int G, int E, int O, int I, int i_pad, int taps, int o_total,
                                                  int o_off, float* __restrict__ dW, float* __restrict__ dfcw,
                                                  float* __restrict__ dfcb, int route) {
  __shared__ float rs[16 * 3];
  __shared__ float drs[16 * 3];
  if (threadIdx.x < 48) drs[threadIdx.x] = 0.f;
  if (threadIdx.x < G * E) {
    int g = threadIdx.x / E, e = threadIdx.x % E;
    rs[g * 3 + e] = fcw ? 1.f / (1.f + expf(-(fcw[e] * types.t[g] + fcb[e]))) : 1.f;
  }
  __syncthreads();
  float dr[GM * 3];
#pragma unroll
  for (int k = 0; k < GM * 3; ++k) dr[k] = 0.f;
  int64_t per = (int64_t)O * I * taps;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < per; idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates (o, tap, i) so the dK read is coalesced over i
    int o = (int)(idx / ((int64_t)taps * I));
    int rem = (int)(idx - (int64_t)o * taps * I);
    int tap = rem / I, i = rem - tap * I;
    float w[3], acc[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) { acc[e] = 0.f; w[e] = (e < E) ? W[(((int64_t)e * O + o) * I + i) * taps + tap] : 0.f; }
#pragma unroll
    for (int g = 0; g < GM; ++g) {
      if (g < G) {
        float d = dK[(((int64_t)g * o_total + o_off + o) * taps + tap) * i_pad + i];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
          if (e < E) {
            acc[e] += rs[g * 3 + e] * d;
            dr[g * 3 + e] += d * w[e];
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 3; ++e)
      if (e < E) dW[(((int64_t)e * O + o) * I + i) * taps + tap] += acc[e];
  }
  if (route) {
    // warp shuffles, one shared-memory atomic per warp and (g, e), then G*E threads issue the global atomics
#pragma unroll
    for (int g = 0; g < GM; ++g) {
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        if (g < G && e < E) {      // uniform across the block
          float v = warp_sum(dr[g * 3 + e]);
          if ((threadIdx.x & 31) == 0) atomicAdd(&drs[g * 3 + e], v);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < G * E) {
      int g = threadIdx.x / E, e = threadIdx.x % E;
      float r = rs[g * 3 + e];
      float sgrad = drs[g * 3 + e] * r * (1.f - r);
      atomicAdd(dfcw + e, sgrad * types.t[g]);
      atomicAdd(dfcb + e, sgrad);
    }
  }
}
extern "C" int rd_condconv_mix_bwd(rd_ctx* ctx, const float* dK, const float* W, const float* fc_w, const float* fc_b,
                                   const float* types, int G, int E, int O, int I, int i_pad, int kh, int kw, int o_total,
                                   int o_off, float* dW, float* dfc_w, float* dfc_b, rd_stream st) {
  if (G < 1 || G > 16 || E < 1 || E > 3) RD_FAIL(ctx, RD_ERR_ARG, "mix_bwd: G in [1,16], E in [1,3]");
  MixTypes mt;
  for (int g = 0; g < 16; ++g) mt.t[g] = (types && g < G) ? types[g] : 0.f;
  int taps = kh * kw;
  int64_t per = (int64_t)O * I * taps;
  int grid = (int)((per + 256 * 4 - 1) / (256 * 4));
  if (grid < 1) grid = 1;
  if (grid > ctx->sm_count * 2) grid = ctx->sm_count * 2;
  cudaStream_t s = (cudaStream_t)st;
  int route = (fc_w && dfc_w && dfc_b) ? 1 : 0;
  if (G <= 4) k_mix_bwd<4><<<grid, 256, 0, s>>>(dK, W, fc_w, fc_b, mt, G, E, O, I, i_pad, taps, o_total, o_off, dW, dfc_w, dfc_b, route);
  else k_mix_bwd<16><<<grid, 256, 0, s>>>(dK, W, fc_w, fc_b, mt, G, E, O, I, i_pad, taps, o_total, o_off, dW, dfc_w, dfc_b, route);
  RD_CHECK_LAUNCH(ctx, "condconv_mix_bwd");
  return RD_OK;
}

// ---- batched variant: one launch for many heads (job table in device memory) -------------------------------------
constexpr int kMixJobElemsPerBlock = 256 * 8;       // few, fat blocks: the routing gradients end in same-address atomics
extern "C" int rd_mix_job_blocks(int O, int I, int taps) {
  int64_t per = (int64_t)O * I * taps;
  int b = (int)((per + kMixJobElemsPerBlock - 1) / kMixJobElemsPerBlock);
  return b < 1 ? 1 : b;
}
constexpr int kMixIC = 64;                      // input channels per shared-memory slice of the batched mixing backward
constexpr int kMixSlice = kMixIC * 17;          // floats per (expert, tensor) slice: taps <= 16 -> row pitch (taps | 1) <= 17
constexpr int kMixDynBytes = 2 * 2 * 3 * kMixSlice * 4;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// GMAX = 4: every job has at most 4 weight groups (one module's 4 contrast types — every launch of the 4-contrast model): the routing-
// gradient accumulators shrink from 48 to 12 registers per thread, 3 blocks per SM instead of 2 (ncu: the kernel is latency bound, 50 %
// of the stall samples wait on global loads and 28 % at the barriers between the three phases of a unit, at 25 % occupancy).
template <int GMAX>
__global__ void __launch_bounds__(256, GMAX <= 4 ? 3 : 2) k_mix_bwd_batched(const rd_mix_job* __restrict__ jobs, int njobs) {
  extern __shared__ float mix_dyn[];             // [2 buffers][W | dW][3 experts][kMixSlice]: see the unit pipeline below
  __shared__ float rs[16 * 3];
  __shared__ float drs[16 * 3];
  __shared__ int job_s;
  if (threadIdx.x == 0) {               // binary search: last job with block_begin <= blockIdx.x
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    job_s = lo;
  }
  if (threadIdx.x < 48) drs[threadIdx.x] = 0.f;
  __syncthreads();
  const rd_mix_job& J = jobs[job_s];
  const int G = J.G, E = J.E, O = J.O, I = J.I, i_pad = J.i_pad, taps = J.taps, o_total = J.o_total, o_off = J.o_off;
  const float* __restrict__ dK = J.dK; const float* __restrict__ W = J.W;
  float* __restrict__ dW = J.dW;
  if (threadIdx.x < G * E) {
    int g = threadIdx.x / E, e = threadIdx.x % E;
    rs[g * 3 + e] = J.fc_w ? 1.f / (1.f + expf(-(J.fc_w[e] * J.types[g] + J.fc_b[e]))) : 1.f;
  }
  __syncthreads();
  if ((int)blockIdx.x == J.block_begin && J.bias_dst)          // the head's bias-gradient slice rides along (first block of the job)
    for (int k = threadIdx.x; k < J.bias_n; k += blockDim.x) atomicAdd(J.bias_dst + k, J.bias_src[k]);    // several jobs may share a bias
  float dr[GMAX * 3];
#pragma unroll
  for (int k = 0; k < GMAX * 3; ++k) dr[k] = 0.f;
  const int per = O * I * taps;
  const int64_t wexp = (int64_t)per;
  const int lb = (int)blockIdx.x - J.block_begin;
  const int64_t gs = (int64_t)o_total * taps * i_pad;
  if (taps > 1) {
    // The expert weights / gradients are (E, O, I, kh, kw), the packed dK is (o, tap, i): with either order as the thread index
    // the other side's accesses touch 32 sectors per warp instruction (the first version ran at 0.6 TB/s).  A unit of work is
    // (o, a chunk of <= kMixIC input channels): its W / dW slice is CONTIGUOUS (ic * taps floats), its dK slice is `taps` runs of
    // ic floats.  The slice goes through shared memory: coalesced W load, dK walked in packed order (the weights read and the
    // mixed gradients written back transposed, odd row pitch = no bank conflicts), coalesced dW += .
    const int tp = taps | 1;
    // I <= kMixIC: a unit is `no` consecutive output channels with all their input channels (still one contiguous W slice and
    // ~1 K elements of work per unit whatever I is); wider layers: one output channel, kMixIC input channels
    const int chunks_i = (I + kMixIC - 1) / kMixIC;
    int no = 1;
    if (chunks_i == 1) { no = (kMixIC * 17) / (I * tp); if (no < 1) no = 1; }
    const int units = chunks_i == 1 ? (O + no - 1) / no : O * chunks_i;
    // Software pipeline over the block's units (round 2): the W slice AND the dW slice (the old gradient values) of unit k + 1 are
    // requested with cp.async into the second buffer pair while unit k is processed, so a unit pays ONE exposed memory latency (its dK
    // loads) instead of three (W load -> barrier -> dK -> barrier -> dW read-modify-write): ncu had this kernel at 0.75 TB/s, latency bound.
    auto unit_geom = [&](int u, int& o0, int& i0, int& ic, int& nol) {
      if (chunks_i == 1) { o0 = u * no; i0 = 0; ic = I; nol = O - o0 < no ? O - o0 : no; }
      else { o0 = u / chunks_i; i0 = (u - o0 * chunks_i) * kMixIC; ic = I - i0 < kMixIC ? I - i0 : kMixIC; nol = 1; }
    };
    auto prefetch = [&](int u, int b) {
      int o0, i0, ic, nol;
      unit_geom(u, o0, i0, ic, nol);
      const int n = nol * ic * taps;
      const int64_t wbase = ((int64_t)o0 * I + i0) * taps;
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int row = t / taps, tap = t - row * taps;
        const int slot = row * tp + tap;
        for (int e = 0; e < E; ++e) {
          cp_async4(smem_u32(mix_dyn + ((b * 2 + 0) * 3 + e) * kMixSlice + slot), W + e * wexp + wbase + t);
          cp_async4(smem_u32(mix_dyn + ((b * 2 + 1) * 3 + e) * kMixSlice + slot), dW + e * wexp + wbase + t);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int kbuf = 0;
    if (lb < units) prefetch(lb, 0);
    for (int u = lb; u < units; u += J.blocks, kbuf ^= 1) {
      int o0, i0, ic, nol;
      unit_geom(u, o0, i0, ic, nol);
      const int n = nol * ic * taps;
      const int64_t wbase = ((int64_t)o0 * I + i0) * taps;
      const bool has_next = u + J.blocks < units;
      if (has_next) prefetch(u + J.blocks, kbuf ^ 1);
      if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      float* wsb = mix_dyn + ((kbuf * 2 + 0) * 3) * kMixSlice;      // [e][slot]: expert weights of the unit
      float* dsb = mix_dyn + ((kbuf * 2 + 1) * 3) * kMixSlice;      // [e][slot]: old dW values, += the mixed gradients
      const float* dk0 = dK + (int64_t)(o_off + o0) * taps * i_pad + i0;
      const int per_o = taps * ic;
      if ((ic & 3) == 0) {
        // four input channels per thread: 16 groups x 16 bytes of dK in flight per thread
        const int ic4 = ic >> 2, n4 = n >> 2, per_o4 = per_o >> 2;
        for (int q = threadIdx.x; q < n4; q += blockDim.x) {
          const int ol = q / per_o4, rem = q - ol * per_o4;
          const int tap = rem / ic4, ii = (rem - tap * ic4) << 2;
          const int slot = (ol * ic + ii) * tp + tap;
          float w[3][4], acc[3][4];
#pragma unroll
          for (int e = 0; e < 3; ++e)
#pragma unroll
            for (int l = 0; l < 4; ++l) { acc[e][l] = 0.f; w[e][l] = (e < E) ? wsb[e * kMixSlice + slot + l * tp] : 0.f; }
          const float* dk = dk0 + ((int64_t)ol * taps + tap) * i_pad + ii;
#pragma unroll
          for (int g = 0; g < GMAX; ++g) {
            if (g < G) {
              const float4 d4 = __ldg(reinterpret_cast<const float4*>(dk + g * gs));
              const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int e = 0; e < 3; ++e)
#pragma unroll
                for (int l = 0; l < 4; ++l) { acc[e][l] += rs[g * 3 + e] * d[l]; dr[g * 3 + e] += d[l] * w[e][l]; }
            }
          }
#pragma unroll
          for (int e = 0; e < 3; ++e)
            if (e < E) {
#pragma unroll
              for (int l = 0; l < 4; ++l) dsb[e * kMixSlice + slot + l * tp] += acc[e][l];
            }
        }
      } else
      for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int ol = q / per_o, rem = q - ol * per_o;
        const int tap = rem / ic, ii = rem - tap * ic;
        const int slot = (ol * ic + ii) * tp + tap;
        float w[3], acc[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) { acc[e] = 0.f; w[e] = (e < E) ? wsb[e * kMixSlice + slot] : 0.f; }
        const float* dk = dk0 + ((int64_t)ol * taps + tap) * i_pad + ii;
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
          if (g < G) {
            const float d = dk[g * gs];
#pragma unroll
            for (int e = 0; e < 3; ++e) { acc[e] += rs[g * 3 + e] * d; dr[g * 3 + e] += d * w[e]; }
          }
        }
#pragma unroll
        for (int e = 0; e < 3; ++e)
          if (e < E) dsb[e * kMixSlice + slot] += acc[e];
      }
      __syncthreads();
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int row = t / taps, tap = t - row * taps;
        for (int e = 0; e < E; ++e) dW[e * wexp + wbase + t] = dsb[e * kMixSlice + row * tp + tap];
      }
      __syncthreads();            // this buffer pair is the prefetch target of the next iteration
    }
  } else {
    for (int idx = lb * blockDim.x + threadIdx.x; idx < per; idx += J.blocks * blockDim.x) {
      const int o = idx / (taps * I);
      const int rem = idx - o * taps * I;
      const int tap = rem / I, i = rem - tap * I;
      const int64_t wi = ((int64_t)o * I + i) * taps + tap;
      float w[3], acc[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) { acc[e] = 0.f; w[e] = (e < E) ? W[e * wexp + wi] : 0.f; }
      const float* dk = dK + ((int64_t)(o_off + o) * taps + tap) * i_pad + i;
#pragma unroll
      for (int g = 0; g < GMAX; ++g) {
        if (g < G) {
          const float d = dk[g * gs];
#pragma unroll
          for (int e = 0; e < 3; ++e) { acc[e] += rs[g * 3 + e] * d; dr[g * 3 + e] += d * w[e]; }
        }
      }
#pragma unroll
      for (int e = 0; e < 3; ++e)
        if (e < E) dW[e * wexp + wi] += acc[e];
    }
  }
  if (J.fc_w && J.dfc_w && J.dfc_b) {
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        if (g < G && e < E) {
          float v = warp_sum(dr[g * 3 + e]);
          if ((threadIdx.x & 31) == 0) atomicAdd(&drs[g * 3 + e], v);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < E) {             // fold the groups in the block: 2 E atomics per block (all blocks of a job hit the same
      const int e = threadIdx.x;       // 2 E addresses; 2 G E atomics per block serialised for ~0.4 ms per launch)
      float sw = 0.f, sb = 0.f;
      for (int g = 0; g < G; ++g) {
        const float r = rs[g * 3 + e];
        const float sgrad = drs[g * 3 + e] * r * (1.f - r);
        sw += sgrad * J.types[g];
        sb += sgrad;
      }
      atomicAdd(J.dfc_w + e, sw);
      atomicAdd(J.dfc_b + e, sb);
    }
  }
}
extern "C" int rd_condconv_mix_bwd_batched(rd_ctx* ctx, const rd_mix_job* jobs_dev, int njobs, int total_blocks, int max_groups, rd_stream st) {
  if (njobs < 1 || total_blocks < 1) return RD_OK;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_mix_bwd_batched<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixDynBytes));
    RD_CUDA(ctx, cudaFuncSetAttribute(k_mix_bwd_batched<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMixDynBytes));
    attr_set = true;
  }
  if (max_groups >= 1 && max_groups <= 4) k_mix_bwd_batched<4><<<total_blocks, 256, kMixDynBytes, (cudaStream_t)st>>>(jobs_dev, njobs);
  else k_mix_bwd_batched<16><<<total_blocks, 256, kMixDynBytes, (cudaStream_t)st>>>(jobs_dev, njobs);
  RD_CHECK_LAUNCH(ctx, "condconv_mix_bwd_batched");
  return RD_OK;
}

template <typename T>
__global__ void k_pad_channels(const T* __restrict__ in, T* __restrict__ out, int64_t pixels, int c, int c_pad) {
  constexpr int V = VecIO<T>::V;
  if (c_pad % V == 0) {               // one 16-byte store per thread
    const int vpp = c_pad / V;
    const int64_t total = pixels * vpp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t p = i / vpp;
      const int vv = (int)(i - p * vpp);
      float v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { const int ch = vv * V + j; v[j] = ch < c ? ldf<T>(in + p * c + ch) : 0.f; }
      VecIO<T>::store(out + p * c_pad + (int64_t)vv * V, v);
    }
    return;
  }
  int64_t total = pixels * c_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i / c_pad;
    int ch = (int)(i - p * c_pad);
    if (ch < c) out[i] = in[p * c + ch]; else stf<T>(out + i, 0.f);
  }
}
extern "C" int rd_pad_channels(rd_ctx* ctx, const void* in, void* out, int64_t pixels, int c, int c_pad, int dtype, rd_stream st) {
  int grid = rd_grid_1d(pixels * c_pad / 4, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_pad_channels<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)in, (T*)out, pixels, c, c_pad)));
  RD_CHECK_LAUNCH(ctx, "pad_channels");
  return RD_OK;
}

// ============================================================================ block gather (+ channel padding)
// out block k = src block index[k] (blocks = `block_pixels` consecutive pixels, i.e. B images), channels [c, c_pad) zero.
// This is the fan-out of s_i / z_j over the (i, j) decodes (src/model.py:3187-3224) in ONE launch, with the zero padding
// the tensor-core kernels need (4 anatomy channels -> 16) folded in; the backward sums the fan-out per source block in
// fp32 and drops the padding channels.  The (<= 32 entry) index list travels in the kernel parameters.
struct GatherIdx { int idx[32]; };

template <typename T>
__global__ void k_gather_blocks(const T* __restrict__ src, T* __restrict__ dst, GatherIdx gi, int nb, int64_t block_pixels,
                                int c, int c_pad) {
  constexpr int V = VecIO<T>::V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c == c_pad && (block_pixels * c) % V == 0) {                   // plain block copy, 16-byte vectors
    const int64_t vec_per_block = block_pixels * c / V, total = vec_per_block * nb;
    for (int64_t i = t0; i < total; i += stride) {
      const int k = (int)(i / vec_per_block);
      const int64_t e = (i - (int64_t)k * vec_per_block) * V;
      float v[V];
      VecIO<T>::load(src + (int64_t)gi.idx[k] * block_pixels * c + e, v);
      VecIO<T>::store(dst + (int64_t)k * block_pixels * c + e, v);
    }
  } else if (c_pad % V == 0) {                                       // one thread per destination vector
    const int vpp = c_pad / V;
    const int64_t total = block_pixels * vpp * nb;
    for (int64_t i = t0; i < total; i += stride) {
      const int vv = (int)(i % vpp);
      const int64_t pk = i / vpp;
      const int k = (int)(pk / block_pixels);
      const int64_t p = pk - (int64_t)k * block_pixels;
      const T* sp = src + ((int64_t)gi.idx[k] * block_pixels + p) * c;
      float v[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { const int ch = vv * V + j; v[j] = ch < c ? ldf<T>(sp + ch) : 0.f; }
      VecIO<T>::store(dst + pk * c_pad + (int64_t)vv * V, v);
    }
  } else {
    const int64_t total = block_pixels * c_pad * nb;
    for (int64_t i = t0; i < total; i += stride) {
      const int ch = (int)(i % c_pad);
      const int64_t pk = i / c_pad;
      const int k = (int)(pk / block_pixels);
      const int64_t p = pk - (int64_t)k * block_pixels;
      stf<T>(dst + i, ch < c ? ldf<T>(src + ((int64_t)gi.idx[k] * block_pixels + p) * c + ch) : 0.f);
    }
  }
}

template <typename T>
__global__ void k_gather_blocks_bwd(const T* __restrict__ dout, T* __restrict__ dsrc, GatherIdx gi, int nb, int nsrc,
                                    int64_t block_pixels, int c, int c_pad) {
  constexpr int V = VecIO<T>::V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c == c_pad && (block_pixels * c) % V == 0) {
    const int64_t vec_per_block = block_pixels * c / V, total = vec_per_block * nsrc;
    for (int64_t i = t0; i < total; i += stride) {
      const int sblk = (int)(i / vec_per_block);
      const int64_t e = (i - (int64_t)sblk * vec_per_block) * V;
      float acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = 0.f;
      for (int k = 0; k < nb; ++k) {
        if (gi.idx[k] != sblk) continue;
        float v[V];
        VecIO<T>::load(dout + (int64_t)k * block_pixels * c + e, v);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[j];
      }
      VecIO<T>::store(dsrc + (int64_t)sblk * block_pixels * c + e, acc);
    }
  } else {
    const int64_t total = block_pixels * c * nsrc;
    for (int64_t i = t0; i < total; i += stride) {
      const int ch = (int)(i % c);
      const int64_t ps = i / c;
      const int sblk = (int)(ps / block_pixels);
      const int64_t p = ps - (int64_t)sblk * block_pixels;
      float acc = 0.f;
      for (int k = 0; k < nb; ++k)
        if (gi.idx[k] == sblk) acc += ldf<T>(dout + ((int64_t)k * block_pixels + p) * c_pad + ch);
      stf<T>(dsrc + i, acc);
    }
  }
}

// The anatomy code's fan-out (4 bf16 channels per pixel, zero-padded to the 16-channel vectors the tensor-core kernels read): one thread
// per pixel, an 8-byte load and c_pad / 8 16-byte stores, the block index from blockIdx.y — the generic kernels above spend their time in
// 64-bit divisions and 2-byte loads (sp6: 103 us forward, 98 us backward for 315 MB each).
__global__ void __launch_bounds__(256) k_gather_pad4_fwd(const bf16* __restrict__ src, bf16* __restrict__ dst, GatherIdx gi, int block_pixels, int c_pad) {
  const int k = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= block_pixels) return;
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(src + ((int64_t)gi.idx[k] * block_pixels + p) * 4));
  uint4* d = reinterpret_cast<uint4*>(dst + ((int64_t)k * block_pixels + p) * c_pad);
  d[0] = make_uint4(v.x, v.y, 0u, 0u);
  for (int j = 1; j < (c_pad >> 3); ++j) d[j] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(256) k_gather_pad4_bwd(const bf16* __restrict__ dout, bf16* __restrict__ dsrc, GatherIdx gi, int nb, int block_pixels, int c_pad) {
  const int sblk = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= block_pixels) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
  for (int k = 0; k < nb; ++k) {
    if (gi.idx[k] != sblk) continue;                       // block-uniform
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(dout + ((int64_t)k * block_pixels + p) * c_pad));
    a0 += __uint_as_float(v.x << 16); a1 += __uint_as_float(v.x & 0xffff0000u);
    a2 += __uint_as_float(v.y << 16); a3 += __uint_as_float(v.y & 0xffff0000u);
  }
  __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
  *reinterpret_cast<uint2*>(dsrc + ((int64_t)sblk * block_pixels + p) * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

extern "C" int rd_gather_blocks_fwd(rd_ctx* ctx, const void* src, void* dst, const int32_t* index_host, int nb, int64_t block_pixels,
                                    int c, int c_pad, int dtype, rd_stream st) {
  if (nb < 1 || nb > 32 || c_pad < c) RD_FAIL(ctx, RD_ERR_ARG, "gather_blocks: 1 <= nb <= 32 and c_pad >= c required");
  GatherIdx gi;
  for (int k = 0; k < 32; ++k) gi.idx[k] = k < nb ? index_host[k] : -1;
  if (dtype == RD_BF16 && c == 4 && c_pad % 8 == 0 && block_pixels < (1 << 30) && ((reinterpret_cast<uintptr_t>(src) & 7u) | (reinterpret_cast<uintptr_t>(dst) & 15u)) == 0) {
    dim3 g2((unsigned)rd_div_up(block_pixels, 256), (unsigned)nb);
    k_gather_pad4_fwd<<<g2, 256, 0, (cudaStream_t)st>>>((const bf16*)src, (bf16*)dst, gi, (int)block_pixels, c_pad);
    RD_CHECK_LAUNCH(ctx, "gather_blocks_fwd");
    return RD_OK;
  }
  int grid = rd_grid_1d(block_pixels * c_pad * nb / 8 + 1, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_gather_blocks<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)src, (T*)dst, gi, nb, block_pixels, c, c_pad)));
  RD_CHECK_LAUNCH(ctx, "gather_blocks_fwd");
  return RD_OK;
}
extern "C" int rd_gather_blocks_bwd(rd_ctx* ctx, const void* dout, void* dsrc, const int32_t* index_host, int nb, int nsrc,
                                    int64_t block_pixels, int c, int c_pad, int dtype, rd_stream st) {
  if (nb < 1 || nb > 32 || c_pad < c) RD_FAIL(ctx, RD_ERR_ARG, "gather_blocks: 1 <= nb <= 32 and c_pad >= c required");
  GatherIdx gi;
  for (int k = 0; k < 32; ++k) gi.idx[k] = k < nb ? index_host[k] : -1;
  if (dtype == RD_BF16 && c == 4 && c_pad % 4 == 0 && block_pixels < (1 << 30) && ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dsrc)) & 7u) == 0) {
    dim3 g2((unsigned)rd_div_up(block_pixels, 256), (unsigned)nsrc);
    k_gather_pad4_bwd<<<g2, 256, 0, (cudaStream_t)st>>>((const bf16*)dout, (bf16*)dsrc, gi, nb, (int)block_pixels, c_pad);
    RD_CHECK_LAUNCH(ctx, "gather_blocks_bwd");
    return RD_OK;
  }
  int grid = rd_grid_1d(block_pixels * c * nsrc / 4 + 1, 256, ctx->sm_count);
  RD_DISPATCH_DTYPE(dtype, (k_gather_blocks_bwd<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)dout, (T*)dsrc, gi, nb, nsrc, block_pixels, c, c_pad)));
  RD_CHECK_LAUNCH(ctx, "gather_blocks_bwd");
  return RD_OK;
}

// dst block d (B images, c_pad channels, channels >= c zero) = block sblk[d] of source sel[d] (two sources with c channels): the
// gradients of the self- and cross-reconstruction stacks written straight into the zero-padded dY the decoder's last convolution
// reads (one pass instead of two gather-backward launches, an add and a channel pad).
struct ScatterMap { int sel[32]; int sblk[32]; };
template <typename T>
__global__ void k_scatter_blocks2(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ dst, ScatterMap mp, int nb,
                                  int64_t block_pixels, int c, int c_pad) {
  const int64_t total = block_pixels * nb;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i / block_pixels);
    const int64_t p = i - (int64_t)d * block_pixels;
    const T* src = (mp.sel[d] ? b : a) + ((int64_t)mp.sblk[d] * block_pixels + p) * c;
    T* o = dst + i * c_pad;
    if (sizeof(T) == 2 && c <= 8 && c_pad == 16) {            // bf16 images (7 channels): 14 source bytes -> two 16-byte stores
      const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < c) w[k >> 1] |= (uint32_t)s16[k] << ((k & 1) * 16);
      uint4* o4 = reinterpret_cast<uint4*>(o);
      o4[0] = make_uint4(w[0], w[1], w[2], w[3]);
      o4[1] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int k = 0; k < c_pad; ++k) o[k] = k < c ? src[k] : T(0.f);
    }
  }
}
extern "C" int rd_scatter_blocks2(rd_ctx* ctx, const void* a, const void* b, void* dst, const int32_t* sel_host, const int32_t* sblk_host,
                                  int nb, int64_t block_pixels, int c, int c_pad, int dtype, rd_stream st) {
  if (nb < 1 || nb > 32 || c_pad < c) RD_FAIL(ctx, RD_ERR_ARG, "scatter_blocks2: 1 <= nb <= 32 and c_pad >= c required");
  ScatterMap mp;
  for (int k = 0; k < 32; ++k) { mp.sel[k] = k < nb ? sel_host[k] : 0; mp.sblk[k] = k < nb ? sblk_host[k] : 0; }
  RD_DISPATCH_DTYPE(dtype, (k_scatter_blocks2<T><<<rd_grid_1d(block_pixels * nb, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(
                               (const T*)a, (const T*)b, (T*)dst, mp, nb, block_pixels, c, c_pad)));
  RD_CHECK_LAUNCH(ctx, "scatter_blocks2");
  return RD_OK;
}

// ============================================================================ per-(group, channel) reductions
// Generic two-level column reduction over NHWC data: for every (group g, channel c) accumulate two
// sums over the group's pixels.  grid (chunks, ctiles, G), block (32 channels, 8 pixel lanes).
// partial layout: [G][chunks][2][C].
// Pixels per CTA: ~32 K elements (a 256-thread CTA then runs ~16 iterations of 16-byte loads per thread), so that the
// low-resolution 128-channel SPADE blocks still spread over a few hundred CTAs instead of one CTA per image.
static inline int red_pixels_per_chunk(int C) {
  int p = 32768 / (C < 1 ? 1 : C);
  int r = 64;
  while (r * 2 <= p && r < 2048) r *= 2;
  return r;
}

extern "C" int rd_norm_partial_chunks(int64_t ppg, int C) {
  int ppc = red_pixels_per_chunk(C);
  int c = (int)((ppg + ppc - 1) / ppc);
  return c < 1 ? 1 : c;
}

template <typename T> struct OpStats {   // sums of (x-K), (x-K)^2 with K = first pixel of the group
  const T* x;
  __device__ __forceinline__ void operator()(int64_t gpix0, int64_t pix, int c, int C, int g, float& a, float& b) const {
    float k = ldf<T>(x + gpix0 * C + c);
    float v = ldf<T>(x + pix * C + c) - k;
    a = v; b = v * v;
  }
};
template <typename T> struct OpNormBwd {  // sums of dy, dy*xhat
  const T* x; const T* dy; const float* mean; const float* invstd;
  __device__ __forceinline__ void operator()(int64_t gpix0, int64_t pix, int c, int C, int g, float& a, float& b) const {
    float d = ldf<T>(dy + pix * C + c);
    float xh = (ldf<T>(x + pix * C + c) - mean[g * C + c]) * invstd[g * C + c];
    a = d; b = d * xh;
  }
};
template <typename T> struct OpSpadeBwd {  // dxhat = dmix*(1+gamma); sums of dxhat, dxhat*zhat; writes dgb.  gs = pixel stride of gamma (2C: inside gb, C: own tensor)
  const T* z; const T* gb; const T* dmix; T* dgb; const float* mean; const float* invstd; int gs;
  __device__ __forceinline__ void operator()(int64_t gpix0, int64_t pix, int c, int C, int g, float& a, float& b) const {
    float dm = ldf<T>(dmix + pix * C + c);
    float zh = (ldf<T>(z + pix * C + c) - mean[g * C + c]) * invstd[g * C + c];
    float gam = ldf<T>(gb + pix * gs + c);
    stf<T>(dgb + pix * 2 * C + c, dm * zh);
    stf<T>(dgb + pix * 2 * C + C + c, dm);
    float dxh = dm * (1.f + gam);
    a = dxh; b = dxh * zh;
  }
};

template <typename Op>
__global__ void k_colreduce_partial(Op op, int64_t ppg, int C, int chunks, int ppc, float* __restrict__ partial) {
  __shared__ float sa[8][33], sb[8][33];
  int g = blockIdx.z, chunk = blockIdx.x;
  int c = blockIdx.y * 32 + threadIdx.x;
  int64_t gpix0 = (int64_t)g * ppg;
  int64_t p0 = (int64_t)chunk * ppc;
  int64_t p1 = p0 + ppc;
  if (p1 > ppg) p1 = ppg;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int64_t p = p0 + threadIdx.y; p < p1; p += 8) {
      float va, vb;
      op(gpix0, gpix0 + p, c, C, g, va, vb);
      a += va; b += vb;
    }
  }
  sa[threadIdx.y][threadIdx.x] = a;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ta += sa[k][threadIdx.x]; tb += sb[k][threadIdx.x]; }
    float* dst = partial + (((int64_t)g * chunks + chunk) * 2) * C;
    dst[c] = ta;
    dst[C + c] = tb;
  }
}

// Vectorised variant (C a multiple of the 16-byte vector width): each thread owns V consecutive channels, a warp reads
// 512 contiguous bytes, per-thread accumulators are combined through shared memory in a fixed order (deterministic).
// Each functor has a per-thread context prepared once per (group, channel vector): the shift / mean / invstd vectors.
template <typename T> struct VOpStats {
  const T* x;
  static constexpr int V = VecIO<T>::V;
  struct Ctx { float k[VecIO<T>::V]; };
  __device__ __forceinline__ void prep(Ctx& cx, int64_t gpix0, int c0, int C, int g) const { VecIO<T>::load(x + gpix0 * C + c0, cx.k); }
  __device__ __forceinline__ void operator()(const Ctx& cx, int64_t pix, int c0, int C, float (&a)[VecIO<T>::V], float (&b)[VecIO<T>::V]) const {
    float v[V];
    VecIO<T>::load(x + pix * C + c0, v);
#pragma unroll
    for (int i = 0; i < V; ++i) { float d = v[i] - cx.k[i]; a[i] += d; b[i] += d * d; }
  }
};
template <typename T> struct VOpNormBwd {
  const T* x; const T* dy; const float* mean; const float* invstd;
  static constexpr int V = VecIO<T>::V;
  struct Ctx { float m[VecIO<T>::V], is[VecIO<T>::V]; };
  __device__ __forceinline__ void prep(Ctx& cx, int64_t gpix0, int c0, int C, int g) const {
#pragma unroll
    for (int i = 0; i < V; ++i) { cx.m[i] = mean[g * C + c0 + i]; cx.is[i] = invstd[g * C + c0 + i]; }
  }
  __device__ __forceinline__ void operator()(const Ctx& cx, int64_t pix, int c0, int C, float (&a)[VecIO<T>::V], float (&b)[VecIO<T>::V]) const {
    float xv[V], dv[V];
    VecIO<T>::load(x + pix * C + c0, xv);
    VecIO<T>::load(dy + pix * C + c0, dv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float xh = (xv[i] - cx.m[i]) * cx.is[i];
      a[i] += dv[i]; b[i] += dv[i] * xh;
    }
  }
};
template <typename T> struct VOpSpadeBwd {
  const T* z; const T* gb; const T* dmix; T* dgb; const float* mean; const float* invstd; int gs;
  static constexpr int V = VecIO<T>::V;
  struct Ctx { float m[VecIO<T>::V], is[VecIO<T>::V]; };
  __device__ __forceinline__ void prep(Ctx& cx, int64_t gpix0, int c0, int C, int g) const {
#pragma unroll
    for (int i = 0; i < V; ++i) { cx.m[i] = mean[g * C + c0 + i]; cx.is[i] = invstd[g * C + c0 + i]; }
  }
  __device__ __forceinline__ void operator()(const Ctx& cx, int64_t pix, int c0, int C, float (&a)[VecIO<T>::V], float (&b)[VecIO<T>::V]) const {
    float zv[V], gv[V], dm[V], o1[V];
    VecIO<T>::load(z + pix * C + c0, zv);
    VecIO<T>::load(gb + pix * gs + c0, gv);
    VecIO<T>::load(dmix + pix * C + c0, dm);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float zh = (zv[i] - cx.m[i]) * cx.is[i];
      o1[i] = dm[i] * zh;
      float dxh = dm[i] * (1.f + gv[i]);
      a[i] += dxh; b[i] += dxh * zh;
    }
    VecIO<T>::store(dgb + pix * 2 * C + c0, o1);
    VecIO<T>::store(dgb + pix * 2 * C + C + c0, dm);
  }
};
// dbias: a = sum dy, b unused
template <typename T> struct VOpSum {
  const T* dy;
  static constexpr int V = VecIO<T>::V;
  struct Ctx {};
  __device__ __forceinline__ void prep(Ctx&, int64_t, int, int, int) const {}
  __device__ __forceinline__ void operator()(const Ctx&, int64_t pix, int c0, int C, float (&a)[VecIO<T>::V], float (&b)[VecIO<T>::V]) const {
    float dv[V];
    VecIO<T>::load(dy + pix * C + c0, dv);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] += dv[i];
  }
};

// grid (chunks, 1, G), block 256.  cv = C / V channel vectors per pixel; thread t handles vector t % cvt of pixel lane t / cvt.
template <typename Op>
__global__ void __launch_bounds__(256) k_colreduce_vec(Op op, int64_t ppg, int C, int chunks, int ppc, float* __restrict__ partial) {
  constexpr int V = Op::V;
  extern __shared__ float sred[];            // [256][2V]
  const int cv = C / V;
  const int g = blockIdx.z, chunk = blockIdx.x;
  const int64_t gpix0 = (int64_t)g * ppg;
  const int64_t p0 = (int64_t)chunk * ppc;
  int64_t p1 = p0 + ppc;
  if (p1 > ppg) p1 = ppg;
  for (int cbase = 0; cbase < cv; cbase += 256) {           // C > 256*V only for very wide layers
    const int cvt = (cv - cbase) < 256 ? (cv - cbase) : 256;
    const int lanes = 256 / cvt;                             // pixel lanes in this block
    const int vec = threadIdx.x % cvt, pl = threadIdx.x / cvt;
    float a[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { a[i] = 0.f; b[i] = 0.f; }
    if (pl < lanes) {
      const int c0 = (cbase + vec) * V;
      typename Op::Ctx cx;
      op.prep(cx, gpix0, c0, C, g);
#pragma unroll 4
      for (int64_t p = p0 + pl; p < p1; p += lanes) op(cx, gpix0 + p, c0, C, a, b);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { sred[threadIdx.x * 2 * V + i] = a[i]; sred[threadIdx.x * 2 * V + V + i] = b[i]; }
    __syncthreads();
    // thread t < cvt*2V sums column (vec, i) over the pixel lanes in a fixed order
    for (int o = threadIdx.x; o < cvt * 2 * V; o += 256) {
      const int v2 = o / (2 * V), i = o % (2 * V);
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sred[(l * cvt + v2) * 2 * V + i];
      const int c = (cbase + v2) * V + (i % V);
      float* dst = partial + (((int64_t)g * chunks + chunk) * 2) * C;
      if (i < V) dst[c] = s; else dst[C + c] = s;
    }
    __syncthreads();
  }
}
template <typename Op>
static inline void launch_colreduce_vec(const Op& op, int G, int64_t ppg, int C, int chunks, float* partial, cudaStream_t s) {
  dim3 grid(chunks, 1, G);
  k_colreduce_vec<<<grid, 256, 256 * 2 * Op::V * sizeof(float), s>>>(op, ppg, C, chunks, red_pixels_per_chunk(C), partial);
}

// finalize statistics: mean / invstd per (g,c); optional running-stat update (sequential over g)
// one thread per (g, c): fold the chunk partials (fixed order), write mean / invstd and the biased variance
template <typename T>
__global__ void k_stats_finalize(const T* __restrict__ x, const float* __restrict__ partial, int G, int64_t ppg, int C,
                                 int chunks, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                 float* __restrict__ var_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  int g = i / C, c = i - g * C;
  float s1 = 0.f, s2 = 0.f;
  const float* src = partial + ((int64_t)g * chunks * 2) * C + c;
  int k = 0;
  for (; k + 8 <= chunks; k += 8) {          // eight chunks of loads in flight, summed in the fixed chunk order
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = __ldg(src + (int64_t)(k + u) * 2 * C); b[u] = __ldg(src + (int64_t)(k + u) * 2 * C + C); }
#pragma unroll
    for (int u = 0; u < 8; ++u) { s1 += a[u]; s2 += b[u]; }
  }
  for (; k < chunks; ++k) { s1 += __ldg(src + (int64_t)k * 2 * C); s2 += __ldg(src + (int64_t)k * 2 * C + C); }
  float n = (float)ppg;
  float shift = ldf<T>(x + (int64_t)g * ppg * C + c);
  float m1 = s1 / n;
  float var = s2 / n - m1 * m1;
  if (var < 0.f) var = 0.f;
  mean[i] = shift + m1;
  invstd[i] = rsqrtf(var + eps);
  if (var_out) var_out[i] = var;
}
// finalize + running-statistics fold in ONE launch (train-mode BatchNorm, G <= 16 calls of the module): block = 32 channels x G
// groups; thread (channel, g) folds its chunk partials, then the g == 0 thread of each channel folds the G statistics into the
// running buffers in call order (same arithmetic as k_stats_finalize + k_running_update)
// Lanes: the chunk partials of a (group, channel) are summed by L threads (chunk l, l + L, ...) and folded in lane order — with one
// thread per (group, channel) the low-channel / many-pixel layers walked ~240 dependent strided loads (10-20 us per launch, 16 launches).
template <typename T>
__global__ void k_stats_finalize_running(const T* __restrict__ x, const float* __restrict__ partial, int G, int64_t ppg, int C, int chunks,
                                         float eps, float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ var_out,
                                         float* running_mean, float* running_var, int64_t* nbt, float momentum, int L) {
  __shared__ float sm_mean[16][33], sm_var[16][33];
  __shared__ float red1[1024], red2[1024];
  const int cl = threadIdx.x & 31, rest = threadIdx.x >> 5;
  const int g = rest / L, l = rest - g * L;
  const int c = blockIdx.x * 32 + cl;
  float s1 = 0.f, s2 = 0.f;
  if (c < C && g < G) {
    const float* src = partial + ((int64_t)g * chunks * 2) * C + c;
    for (int k = l; k < chunks; k += L) { s1 += __ldg(src + (int64_t)k * 2 * C); s2 += __ldg(src + (int64_t)k * 2 * C + C); }
  }
  red1[threadIdx.x] = s1; red2[threadIdx.x] = s2;
  __syncthreads();
  if (c < C && g < G && l == 0) {
    for (int j = 1; j < L; ++j) { s1 += red1[threadIdx.x + 32 * j]; s2 += red2[threadIdx.x + 32 * j]; }
    const float n = (float)ppg;
    const float shift = ldf<T>(x + (int64_t)g * ppg * C + c);
    const float m1 = s1 / n;
    float var = s2 / n - m1 * m1;
    if (var < 0.f) var = 0.f;
    const int i = g * C + c;
    mean[i] = shift + m1;
    invstd[i] = rsqrtf(var + eps);
    if (var_out) var_out[i] = var;
    sm_mean[g][cl] = shift + m1;
    sm_var[g][cl] = var;
  }
  __syncthreads();
  if (rest == 0 && c < C) {
    float rm = running_mean[c], rv = running_var[c];
    const float n = (float)ppg;
    for (int gg = 0; gg < G; ++gg) {
      const float v = sm_var[gg][cl];
      const float unb = (ppg > 1) ? v * n / (n - 1.f) : v;
      rm = (1.f - momentum) * rm + momentum * sm_mean[gg][cl];
      rv = (1.f - momentum) * rv + momentum * unb;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
  if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += G;
}
// running statistics: the G group statistics folded in group order (one BatchNorm module called G times)
__global__ void k_running_update(const float* __restrict__ mean, const float* __restrict__ var, int G, int64_t ppg, int C,
                                 float* running_mean, float* running_var, int64_t* nbt, float momentum) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float rm = running_mean[c], rv = running_var[c];
    float n = (float)ppg;
    for (int g = 0; g < G; ++g) {
      float v = var[g * C + c];
      float unb = (ppg > 1) ? v * n / (n - 1.f) : v;
      rm = (1.f - momentum) * rm + momentum * mean[g * C + c];
      rv = (1.f - momentum) * rv + momentum * unb;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
  if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += G;
}

extern "C" int rd_norm_stats(rd_ctx* ctx, const void* x, int G, int64_t ppg, int C, int dtype, float eps, float* partial,
                             float* mean, float* invstd, float* running_mean, float* running_var, int64_t* nbt,
                             float momentum, rd_stream st) {
  int chunks = rd_norm_partial_chunks(ppg, C);
  dim3 grid(chunks, rd_div_up(C, 32), G), block(32, 8);
  cudaStream_t s = (cudaStream_t)st;
  // workspace tail [G][2][C] (also used by the backward) holds the biased variances for the running-stat pass
  float* var_ws = running_mean ? partial + (int64_t)G * chunks * 2 * C : nullptr;
  RD_DISPATCH_DTYPE(dtype, {
    if (C % VecIO<T>::V == 0) {
      VOpStats<T> op{(const T*)x};
      launch_colreduce_vec(op, G, ppg, C, chunks, partial, s);
    } else {
      OpStats<T> op{(const T*)x};
      k_colreduce_partial<<<grid, block, 0, s>>>(op, ppg, C, chunks, red_pixels_per_chunk(C), partial);
    }
    RD_CHECK_LAUNCH(ctx, "norm_stats_partial");
    if (running_mean && G <= 16) {
      int L = 32 / G;                           // 32 channels x G groups x L lanes <= 1024 threads
      if (L > 8) L = 8;
      if (L < 1) L = 1;
      k_stats_finalize_running<T><<<rd_div_up(C, 32), 32 * G * L, 0, s>>>((const T*)x, partial, G, ppg, C, chunks, eps, mean, invstd, var_ws,
                                                                         running_mean, running_var, nbt, momentum, L);
      RD_CHECK_LAUNCH(ctx, "norm_stats_finalize_running");
    } else {
    k_stats_finalize<T><<<rd_div_up(G * C, 128), 128, 0, s>>>((const T*)x, partial, G, ppg, C, chunks, eps, mean, invstd, var_ws);
    RD_CHECK_LAUNCH(ctx, "norm_stats_finalize");
    }
  });
  if (running_mean && G > 16) {
    k_running_update<<<rd_div_up(C, 128), 128, 0, s>>>(mean, var_ws, G, ppg, C, running_mean, running_var, nbt, momentum);
    RD_CHECK_LAUNCH(ctx, "norm_running_update");
  }
  return RD_OK;
}

__global__ void k_eval_stats(const float* rm, const float* rv, int G, int C, float eps, float* mean, float* invstd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G * C) {
    int c = i % C;
    mean[i] = rm[c];
    invstd[i] = rsqrtf(rv[c] + eps);
  }
}
extern "C" int rd_norm_eval_stats(rd_ctx* ctx, const float* rm, const float* rv, int G, int C, float eps, float* mean,
                                  float* invstd, rd_stream st) {
  k_eval_stats<<<rd_div_up(G * C, 128), 128, 0, (cudaStream_t)st>>>(rm, rv, G, C, eps, mean, invstd);
  RD_CHECK_LAUNCH(ctx, "norm_eval_stats");
  return RD_OK;
}

template <typename T, int V>
__global__ void k_norm_apply(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                             const float* __restrict__ weight, const float* __restrict__ bias, T* __restrict__ y,
                             int64_t ppg, int C, int64_t total_vec) {
  int cv = C / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / cv;
    int c = (int)(i - pix * cv) * V;
    int g = (int)(pix / ppg);
    float v[4];
    if (V == 4) Vec4<T>::load(x + pix * C + c, v); else v[0] = ldf<T>(x + pix * C + c);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = (v[k] - mean[g * C + c + k]) * invstd[g * C + c + k];
      if (weight) t = t * weight[c + k] + bias[c + k];
      v[k] = t;
    }
    if (V == 4) Vec4<T>::store(y + pix * C + c, v); else stf<T>(y + pix * C + c, v[0]);
  }
}
// 16-byte vectors (8 bf16 / 4 fp32 channels per thread), the group index from one division per vector
template <typename T>
__global__ void __launch_bounds__(256) k_norm_apply_vec(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ weight, const float* __restrict__ bias, T* __restrict__ y,
                                                        int64_t ppg, int C, int64_t total_vec) {
  constexpr int V = VecIO<T>::V;
  const int cv = C / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / cv;
    const int c = (int)(i - pix * cv) * V;
    const int g = (int)(pix / ppg);
    float v[V];
    VecIO<T>::load(x + pix * C + c, v);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = (v[k] - mean[g * C + c + k]) * invstd[g * C + c + k];
      if (weight) t = t * weight[c + k] + bias[c + k];
      v[k] = t;
    }
    VecIO<T>::store(y + pix * C + c, v);
  }
}
extern "C" int rd_norm_apply(rd_ctx* ctx, const void* x, const float* mean, const float* invstd, const float* weight,
                             const float* bias, void* y, int G, int64_t ppg, int C, int dtype, rd_stream st) {
  int64_t total = (int64_t)G * ppg * C;
  cudaStream_t s = (cudaStream_t)st;
  const int vfull = dtype == RD_F32 ? 4 : 8;
  if (C % vfull == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0) {
    int grid = rd_grid_1d(total / vfull, 256, ctx->sm_count);
    RD_DISPATCH_DTYPE(dtype, (k_norm_apply_vec<T><<<grid, 256, 0, s>>>((const T*)x, mean, invstd, weight, bias, (T*)y, ppg, C, total / vfull)));
  } else if (C % 4 == 0) {
    int grid = rd_grid_1d(total / 4, 256, ctx->sm_count);
    RD_DISPATCH_DTYPE(dtype, (k_norm_apply<T, 4><<<grid, 256, 0, s>>>((const T*)x, mean, invstd, weight, bias, (T*)y, ppg, C, total / 4)));
  } else {
    int grid = rd_grid_1d(total, 256, ctx->sm_count);
    RD_DISPATCH_DTYPE(dtype, (k_norm_apply<T, 1><<<grid, 256, 0, s>>>((const T*)x, mean, invstd, weight, bias, (T*)y, ppg, C, total)));
  }
  RD_CHECK_LAUNCH(ctx, "norm_apply");
  return RD_OK;
}

// sums[g][2][C] from partials; optional affine-parameter gradients (+=)
__global__ void k_bwd_finalize(const float* __restrict__ partial, int G, int C, int chunks, float* __restrict__ sums) {
  // grid (ceil(C / 32), G), block (32, 8): 8 lanes per (group, channel) sum chunks l, l + 8, ... and are folded in lane order
  __shared__ float r1[8][33], r2[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, g = blockIdx.y, l = threadIdx.y;
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    const float* src = partial + ((int64_t)g * chunks * 2) * C + c;
    for (int k = l; k < chunks; k += 8) { s1 += __ldg(src + (int64_t)k * 2 * C); s2 += __ldg(src + (int64_t)k * 2 * C + C); }
  }
  r1[l][threadIdx.x] = s1; r2[l][threadIdx.x] = s2;
  __syncthreads();
  if (l == 0 && c < C) {
    for (int j = 1; j < 8; ++j) { s1 += r1[j][threadIdx.x]; s2 += r2[j][threadIdx.x]; }
    sums[((int64_t)g * 2) * C + c] = s1;
    sums[((int64_t)g * 2 + 1) * C + c] = s2;
  }
}
__global__ void k_bwd_param_grads(const float* __restrict__ sums, int G, int C, float* dweight, float* dbias) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float tw = 0.f, tb = 0.f;
  for (int g = 0; g < G; ++g) { tb += sums[((int64_t)g * 2) * C + c]; tw += sums[((int64_t)g * 2 + 1) * C + c]; }
  if (dweight) dweight[c] += tw;
  if (dbias) dbias[c] += tb;
}
template <typename T>
__global__ void k_norm_bwd_apply(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, const float* __restrict__ weight,
                                 const float* __restrict__ sums, T* __restrict__ dx, int64_t ppg, int C, int64_t total) {
  float inv_n = 1.f / (float)ppg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / C;
    int c = (int)(i - pix * C);
    int g = (int)(pix / ppg);
    float is = invstd[g * C + c];
    float xh = (ldf<T>(x + i) - mean[g * C + c]) * is;
    float s1 = sums[((int64_t)g * 2) * C + c], s2 = sums[((int64_t)g * 2 + 1) * C + c];
    float w = weight ? weight[c] : 1.f;
    stf<T>(dx + i, w * is * (ldf<T>(dy + i) - s1 * inv_n - xh * s2 * inv_n));
  }
}
// 16-byte vectors over the channel axis (C % V == 0); same expression per element as the scalar kernel
template <typename T>
__global__ void __launch_bounds__(256) k_norm_bwd_apply_vec(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ weight,
                                                            const float* __restrict__ sums, T* __restrict__ dx, int64_t ppg, int C, int64_t total_vec) {
  constexpr int V = VecIO<T>::V;
  const int cv = C / V;
  const float inv_n = 1.f / (float)ppg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / cv;
    const int c = (int)(i - pix * cv) * V;
    const int g = (int)(pix / ppg);
    float xv[V], dv[V], o[V];
    VecIO<T>::load(x + pix * C + c, xv);
    VecIO<T>::load(dy + pix * C + c, dv);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float is = invstd[g * C + c + k];
      const float xh = (xv[k] - mean[g * C + c + k]) * is;
      const float s1 = sums[((int64_t)g * 2) * C + c + k], s2 = sums[((int64_t)g * 2 + 1) * C + c + k];
      const float w = weight ? weight[c + k] : 1.f;
      o[k] = w * is * (dv[k] - s1 * inv_n - xh * s2 * inv_n);
    }
    VecIO<T>::store(dx + pix * C + c, o);
  }
}
extern "C" int rd_norm_bwd(rd_ctx* ctx, const void* x, const void* dy, const float* mean, const float* invstd,
                           const float* weight, void* dx, float* dweight, float* dbias, float* partial, int G,
                           int64_t ppg, int C, int dtype, rd_stream st) {
  int chunks = rd_norm_partial_chunks(ppg, C);
  dim3 grid(chunks, rd_div_up(C, 32), G), block(32, 8);
  cudaStream_t s = (cudaStream_t)st;
  float* sums = partial + (int64_t)G * chunks * 2 * C;   // workspace tail: [G][2][C]
  int64_t total = (int64_t)G * ppg * C;
  RD_DISPATCH_DTYPE(dtype, {
    if (C % VecIO<T>::V == 0) {
      VOpNormBwd<T> op{(const T*)x, (const T*)dy, mean, invstd};
      launch_colreduce_vec(op, G, ppg, C, chunks, partial, s);
    } else {
      OpNormBwd<T> op{(const T*)x, (const T*)dy, mean, invstd};
      k_colreduce_partial<<<grid, block, 0, s>>>(op, ppg, C, chunks, red_pixels_per_chunk(C), partial);
    }
    RD_CHECK_LAUNCH(ctx, "norm_bwd_partial");
    k_bwd_finalize<<<dim3(rd_div_up(C, 32), G), dim3(32, 8), 0, s>>>(partial, G, C, chunks, sums);
    RD_CHECK_LAUNCH(ctx, "norm_bwd_finalize");
    if (dweight || dbias) {
      k_bwd_param_grads<<<rd_div_up(C, 128), 128, 0, s>>>(sums, G, C, dweight, dbias);
      RD_CHECK_LAUNCH(ctx, "norm_bwd_param_grads");
    }
    if (C % VecIO<T>::V == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15u) == 0) {
      const int64_t total_vec = total / VecIO<T>::V;
      k_norm_bwd_apply_vec<T><<<rd_grid_1d(total_vec, 256, ctx->sm_count), 256, 0, s>>>((const T*)x, (const T*)dy, mean, invstd, weight,
                                                                                         sums, (T*)dx, ppg, C, total_vec);
    } else
    k_norm_bwd_apply<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, s>>>((const T*)x, (const T*)dy, mean, invstd, weight,
                                                                                 sums, (T*)dx, ppg, C, total);
    RD_CHECK_LAUNCH(ctx, "norm_bwd_apply");
  });
  return RD_OK;
}

// ============================================================================ SPADE modulation
template <typename T>
__global__ void k_spade_fwd(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                            const T* __restrict__ gb, T* __restrict__ mix, int64_t hw, int C, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / C;
    int c = (int)(i - pix * C);
    int n = (int)(pix / hw);
    float zh = (ldf<T>(z + i) - mean[n * C + c]) * invstd[n * C + c];
    float g = ldf<T>(gb + pix * 2 * C + c), b = ldf<T>(gb + pix * 2 * C + C + c);
    stf<T>(mix + i, zh * (1.f + g) + b);
  }
}
template <typename T>
__global__ void k_spade_fwd_vec(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const T* __restrict__ gb, T* __restrict__ mix, int64_t hw, int C, int64_t total_vec) {
  constexpr int V = VecIO<T>::V;
  const int cv = C / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / cv;
    int c = (int)(i - pix * cv) * V;
    int n = (int)(pix / hw);
    float zv[V], g[V], b[V], o[V];
    VecIO<T>::load(z + pix * C + c, zv);
    VecIO<T>::load(gb + pix * 2 * C + c, g);
    VecIO<T>::load(gb + pix * 2 * C + C + c, b);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = (zv[k] - mean[n * C + c + k]) * invstd[n * C + c + k] * (1.f + g[k]) + b[k];
    VecIO<T>::store(mix + pix * C + c, o);
  }
}
extern "C" int rd_spade_modulate_fwd(rd_ctx* ctx, const void* z, const float* mean, const float* invstd, const void* gb,
                                     void* mix, int N, int64_t hw, int C, int dtype, rd_stream st) {
  int64_t total = (int64_t)N * hw * C;
  int grid = rd_grid_1d(total, 256, ctx->sm_count);
  if ((dtype == RD_BF16 && C % 8 == 0) || (dtype == RD_F32 && C % 4 == 0)) {
    RD_DISPATCH_DTYPE(dtype, (k_spade_fwd_vec<T><<<rd_grid_1d(total / VecIO<T>::V, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)z, mean, invstd, (const T*)gb, (T*)mix, hw, C, total / VecIO<T>::V)));
    RD_CHECK_LAUNCH(ctx, "spade_modulate_fwd");
    return RD_OK;
  }
  RD_DISPATCH_DTYPE(dtype, (k_spade_fwd<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)z, mean, invstd, (const T*)gb, (T*)mix, hw, C, total)));
  RD_CHECK_LAUNCH(ctx, "spade_modulate_fwd");
  return RD_OK;
}
template <typename T>
__global__ void k_spade_bwd_apply(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const T* __restrict__ gb, const T* __restrict__ dmix, const float* __restrict__ sums,
                                  T* __restrict__ dz, int64_t hw, int C, int64_t total, int gs) {
  float inv_n = 1.f / (float)hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / C;
    int c = (int)(i - pix * C);
    int n = (int)(pix / hw);
    float is = invstd[n * C + c];
    float zh = (ldf<T>(z + i) - mean[n * C + c]) * is;
    float dxh = ldf<T>(dmix + i) * (1.f + ldf<T>(gb + pix * gs + c));
    float s1 = sums[((int64_t)n * 2) * C + c], s2 = sums[((int64_t)n * 2 + 1) * C + c];
    stf<T>(dz + i, is * (dxh - s1 * inv_n - zh * s2 * inv_n));
  }
}
template <typename T>
__global__ void k_spade_bwd_apply_vec(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const T* __restrict__ gb, const T* __restrict__ dmix, const float* __restrict__ sums,
                                      T* __restrict__ dz, int64_t hw, int C, int64_t total_vec, int gs) {
  constexpr int V = VecIO<T>::V;
  const int cv = C / V;
  const float inv_n = 1.f / (float)hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t pix = i / cv;
    int c = (int)(i - pix * cv) * V;
    int n = (int)(pix / hw);
    float zv[V], g[V], dm[V], o[V];
    VecIO<T>::load(z + pix * C + c, zv);
    VecIO<T>::load(gb + pix * gs + c, g);
    VecIO<T>::load(dmix + pix * C + c, dm);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float is = invstd[n * C + c + k];
      float zh = (zv[k] - mean[n * C + c + k]) * is;
      float dxh = dm[k] * (1.f + g[k]);
      float s1 = sums[((int64_t)n * 2) * C + c + k], s2 = sums[((int64_t)n * 2 + 1) * C + c + k];
      o[k] = is * (dxh - s1 * inv_n - zh * s2 * inv_n);
    }
    VecIO<T>::store(dz + pix * C + c, o);
  }
}
constexpr int kSpfPPT = 2;                    // pixels per thread and ticket of k_spade_bwd_fused
static inline int spf_chunks_per_image(int64_t hw, int C, int dtype) {
  const int V = dtype == RD_BF16 ? 8 : 4;
  const int cv = C / V < 1 ? 1 : C / V;
  const int ppc = (256 / (cv > 256 ? 256 : cv)) * kSpfPPT;    // pixels per ticket of k_spade_bwd_fused
  return (int)((hw + ppc - 1) / ppc);
}
// floats of workspace rd_spade_modulate_bwd(_g) needs: partial sums per chunk + sums per image, for whichever kernel form runs
extern "C" int64_t rd_spade_bwd_workspace(int N, int64_t hw, int C, int dtype) {
  int64_t a = (int64_t)rd_norm_partial_chunks(hw, C), b = (int64_t)spf_chunks_per_image(hw, C, dtype);
  int64_t chunks = a > b ? a : b;
  return (int64_t)N * chunks * 2 * C + (int64_t)N * 2 * C;
}
static int spade_modulate_bwd_impl(rd_ctx* ctx, const void* z, const float* mean, const float* invstd, const void* gb, int gs,
                                   const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                                   int dtype, rd_stream st);

// ---- single-pass SPADE modulation backward: z, gamma, dmix come from HBM ONCE.
// The two-pass form (k_colreduce_vec<VOpSpadeBwd> then k_spade_bwd_apply_vec) reads the three [N, H, W, C] tensors twice: 9 tensor
// units of HBM traffic per block against 6 when the second read hits the L2.  A CTA takes a ticket = (image, chunk of 256 / (C / V) * 4
// pixels) from an atomic counter, streams the chunk once (d(gamma|beta) and the chunk's partial sums), arrives on the image's counter — the
// LAST CTA of an image folds the partials in fixed chunk order (deterministic) and releases the image's flag — and only THEN finishes the
// PREVIOUS ticket it held: by now that image's flag is normally set, and the chunk's bytes are still in the L2 (between the two visits
// every resident CTA has streamed one ticket: ~20 MB).  Tickets go round-robin (ticket = blockIdx.x + k gridDim.x), every CTA streams its
// round-k ticket BEFORE it waits for the image of its round-(k - 1) ticket, and the grid is at most the number of CTAs the idle GPU holds
// (3 per SM, checked on the host): every chunk an image waits for is streamed without waiting on anything, so the flag wait cannot deadlock.  The last CTA to leave zeroes the counter slot for the next launch (slots are handed out round-robin per launch from rd_ctx).
// (First version, round 2: the chunk was held in REGISTERS across the barrier — 1.79 ms against 0.74 ms for the two passes at sp6: the
// register file bounds the bytes in flight per SM and the barrier chain is ~8 us long.)
template <typename T> __device__ __forceinline__ void spf_unpack(const uint4& t, float (&v)[VecIO<T>::V]);
template <> __device__ __forceinline__ void spf_unpack<float>(const uint4& t, float (&v)[4]) {
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}
template <> __device__ __forceinline__ void spf_unpack<bf16>(const uint4& t, float (&v)[8]) {
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
  v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}
// second visit of a ticket: dz of its pixels from the L2-resident z, gamma, dmix and the image's sums
template <typename T>
__device__ __forceinline__ void spf_phase2(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                                           const T* __restrict__ gb, int gs, const T* __restrict__ dmix, T* __restrict__ dz,
                                           const float* sums, int img, int p0, int lanes, int hw, int C, int c0, float inv_n) {
  constexpr int V = VecIO<T>::V;
  float m[V], is[V], s1v[V], s2v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    m[i] = __ldg(mean + img * C + c0 + i); is[i] = __ldg(invstd + img * C + c0 + i);
    s1v[i] = __ldcg(sums + (int64_t)img * 2 * C + c0 + i) * inv_n;
    s2v[i] = __ldcg(sums + (int64_t)img * 2 * C + C + c0 + i) * inv_n;
  }
  uint4 zr[kSpfPPT], gr[kSpfPPT], dr[kSpfPPT];
#pragma unroll
  for (int j = 0; j < kSpfPPT; ++j) {
    const int p = p0 + j * lanes;
    if (p < hw) {
      const int64_t pix = (int64_t)img * hw + p;
      zr[j] = __ldcg(reinterpret_cast<const uint4*>(z + pix * C + c0));
      gr[j] = __ldcg(reinterpret_cast<const uint4*>(gb + pix * gs + c0));
      dr[j] = __ldcg(reinterpret_cast<const uint4*>(dmix + pix * C + c0));
    }
  }
#pragma unroll
  for (int j = 0; j < kSpfPPT; ++j) {
    const int p = p0 + j * lanes;
    if (p < hw) {
      float zv[V], gv[V], dm[V], o[V];
      spf_unpack<T>(zr[j], zv); spf_unpack<T>(gr[j], gv); spf_unpack<T>(dr[j], dm);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float zh = (zv[i] - m[i]) * is[i];
        const float dxh = dm[i] * (1.f + gv[i]);
        o[i] = is[i] * (dxh - s1v[i] - zh * s2v[i]);
      }
      VecIO<T>::store(dz + ((int64_t)img * hw + p) * C + c0, o);
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256, 3)
k_spade_bwd_fused(const T* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd, const T* __restrict__ gb, int gs,
                  const T* __restrict__ dmix, T* __restrict__ dz, T* __restrict__ dgb, float* partial, float* sums, int* sync,
                  int N, int hw, int C, int cpi) {
  constexpr int V = VecIO<T>::V;
  __shared__ float wred[8][32][2 * V];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cv = C / V;                        // 16-byte vectors per pixel: a power of two <= 32
  const int lanes = 256 / cv;                  // pixel lanes of the block
  const int vec = tid & (cv - 1), pl = tid / cv;
  const int c0 = vec * V;
  const int ppc = lanes * kSpfPPT;
  const float inv_n = 1.f / (float)hw;
  int* arrive = sync + 2;
  int* ready = sync + 2 + N;
  const int total = N * cpi;
  int prev_img = -1, prev_p0 = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {       // round-robin tickets (a shared ticket counter: 61 440 same-address atomics = 0.9 ms)
    const int img = t / cpi, chunk = t - img * cpi;
    const int p0 = chunk * ppc + pl;
    {
      // ---- first visit: stream the chunk
      uint4 zr[kSpfPPT], gr[kSpfPPT], dr[kSpfPPT];
      const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int j = 0; j < kSpfPPT; ++j) {
        const int p = p0 + j * lanes;
        zr[j] = zero4; gr[j] = zero4; dr[j] = zero4;
        if (p < hw) {
          const int64_t pix = (int64_t)img * hw + p;
          zr[j] = *reinterpret_cast<const uint4*>(z + pix * C + c0);
          gr[j] = *reinterpret_cast<const uint4*>(gb + pix * gs + c0);
          dr[j] = *reinterpret_cast<const uint4*>(dmix + pix * C + c0);
        }
      }
      float m[V], is[V], a[V], b[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { m[i] = __ldg(mean + img * C + c0 + i); is[i] = __ldg(invstd + img * C + c0 + i); a[i] = 0.f; b[i] = 0.f; }
#pragma unroll
      for (int j = 0; j < kSpfPPT; ++j) {
        const int p = p0 + j * lanes;
        float zv[V], gv[V], dm[V], o1[V];
        spf_unpack<T>(zr[j], zv); spf_unpack<T>(gr[j], gv); spf_unpack<T>(dr[j], dm);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float zh = (zv[i] - m[i]) * is[i];
          o1[i] = dm[i] * zh;
          const float dxh = dm[i] * (1.f + gv[i]);
          a[i] += dxh; b[i] += dxh * zh;           // out-of-range pixels carry dm = 0
        }
        if (p < hw) {
          const int64_t pix = (int64_t)img * hw + p;
          VecIO<T>::store(dgb + pix * 2 * C + c0, o1);
          *reinterpret_cast<uint4*>(dgb + pix * 2 * C + C + c0) = dr[j];
        }
      }
      // block sums: butterfly over the pixel lanes of a warp, then the 8 warps in order
      for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < V; ++i) { a[i] += __shfl_xor_sync(0xffffffffu, a[i], off); b[i] += __shfl_xor_sync(0xffffffffu, b[i], off); }
      }
      if (lane < cv) {
#pragma unroll
        for (int i = 0; i < V; ++i) { wred[warp][lane][i] = a[i]; wred[warp][lane][V + i] = b[i]; }
      }
    }
    __syncthreads();
    for (int o = tid; o < cv * 2 * V; o += 256) {
      const int v2 = o / (2 * V), i = o - v2 * 2 * V;
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += wred[w][v2][i];
      const int c = v2 * V + (i < V ? i : i - V);
      float* dst = partial + ((int64_t)t * 2) * C;
      __stcg(dst + (i < V ? c : C + c), s);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&arrive[img], 1) == cpi - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int o = tid; o < 2 * C; o += 256) {
        const float* src = partial + ((int64_t)img * cpi * 2) * C + o;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int k = 0;
        for (; k + 4 <= cpi; k += 4) {
          s0 += __ldcg(src + (int64_t)k * 2 * C); s1 += __ldcg(src + (int64_t)(k + 1) * 2 * C);
          s2 += __ldcg(src + (int64_t)(k + 2) * 2 * C); s3 += __ldcg(src + (int64_t)(k + 3) * 2 * C);
        }
        for (; k < cpi; ++k) s0 += __ldcg(src + (int64_t)k * 2 * C);
        __stcg(sums + (int64_t)img * 2 * C + o, (s0 + s1) + (s2 + s3));
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(&ready[img], 1);
    }
    // ---- second visit of the PREVIOUS ticket (its image's flag has had a whole ticket period to be set)
    if (prev_img >= 0) {
      if (tid == 0) {
        while (*reinterpret_cast<volatile int*>(&ready[prev_img]) == 0) __nanosleep(100);
        __threadfence();
      }
      __syncthreads();
      spf_phase2<T>(z, mean, invstd, gb, gs, dmix, dz, sums, prev_img, prev_p0, lanes, hw, C, c0, inv_n);
    }
    prev_img = img; prev_p0 = p0;
    __syncthreads();                               // s_last / wred are reused by the next ticket
  }
  if (prev_img >= 0) {
    if (tid == 0) {
      while (*reinterpret_cast<volatile int*>(&ready[prev_img]) == 0) __nanosleep(100);
      __threadfence();
    }
    __syncthreads();
    spf_phase2<T>(z, mean, invstd, gb, gs, dmix, dz, sums, prev_img, prev_p0, lanes, hw, C, c0, inv_n);
  }
  // the last CTA to leave zeroes the slot for the next launch
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&sync[1], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    for (int i = tid; i < 2 + 2 * N; i += 256) sync[i] = 0;
  }
}
extern "C" int rd_spade_modulate_bwd(rd_ctx* ctx, const void* z, const float* mean, const float* invstd, const void* gb,
                                     const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                                     int dtype, rd_stream st) {
  return spade_modulate_bwd_impl(ctx, z, mean, invstd, gb, 2 * C, dmix, dz, dgb, partial, N, hw, C, dtype, st);
}
extern "C" int rd_spade_modulate_bwd_g(rd_ctx* ctx, const void* z, const float* mean, const float* invstd, const void* gamma,
                                       const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                                       int dtype, rd_stream st) {
  return spade_modulate_bwd_impl(ctx, z, mean, invstd, gamma, C, dmix, dz, dgb, partial, N, hw, C, dtype, st);
}
static int spade_modulate_bwd_impl(rd_ctx* ctx, const void* z, const float* mean, const float* invstd, const void* gb, int gs,
                                   const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                                   int dtype, rd_stream st) {
  int chunks = rd_norm_partial_chunks(hw, C);
  dim3 grid(chunks, rd_div_up(C, 32), N), block(32, 8);
  cudaStream_t s = (cudaStream_t)st;
  int64_t total = (int64_t)N * hw * C;
  {
    // single-pass kernel (see k_spade_bwd_fused): needs the ctx's counter slots, a power-of-two number of 16-byte vectors per pixel,
    // 16-byte aligned rows, and every image's chunks co-resident
    const char* fe = getenv("RD_B200_SPADE_BWD_FUSED");          // read per call: the parity test switches between the two forms
    const bool fused_on = fe && atoi(fe) != 0;                    // measured slower than the two passes: opt-in
    const int V = dtype == RD_BF16 ? 8 : 4;
    const int cv = C / V;
    const int cpi = spf_chunks_per_image(hw, C, dtype);
    const int max_grid = 3 * ctx->sm_count;
    const bool aligned = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gb) | reinterpret_cast<uintptr_t>(dmix) |
                           reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dgb)) & 15u) == 0 && gs % V == 0;
    if (fused_on && ctx->spf_sync && C % V == 0 && cv >= 1 && cv <= 32 && (cv & (cv - 1)) == 0 && aligned && N <= kSpfMaxN &&
        cpi <= max_grid && hw < (1 << 30)) {
      int* sync = ctx->spf_sync + (size_t)(ctx->spf_next++ % kSpfSlots) * kSpfSlotInts;
      float* sums = partial + (int64_t)N * cpi * 2 * C;
      const int64_t tickets = (int64_t)N * cpi;
      const int g = (int)(tickets < max_grid ? tickets : max_grid);
      RD_DISPATCH_DTYPE(dtype, {
        k_spade_bwd_fused<T><<<g, 256, 0, s>>>((const T*)z, mean, invstd, (const T*)gb, gs, (const T*)dmix, (T*)dz, (T*)dgb, partial,
                                               sums, sync, N, (int)hw, C, cpi);
        RD_CHECK_LAUNCH(ctx, "spade_bwd_fused");
      });
      return RD_OK;
    }
  }
  float* sums = partial + (int64_t)N * chunks * 2 * C;
  RD_DISPATCH_DTYPE(dtype, {
    if (C % VecIO<T>::V == 0) {
      VOpSpadeBwd<T> op{(const T*)z, (const T*)gb, (const T*)dmix, (T*)dgb, mean, invstd, gs};
      launch_colreduce_vec(op, N, hw, C, chunks, partial, s);
    } else {
      OpSpadeBwd<T> op{(const T*)z, (const T*)gb, (const T*)dmix, (T*)dgb, mean, invstd, gs};
      k_colreduce_partial<<<grid, block, 0, s>>>(op, hw, C, chunks, red_pixels_per_chunk(C), partial);
    }
    RD_CHECK_LAUNCH(ctx, "spade_bwd_partial");
    k_bwd_finalize<<<dim3(rd_div_up(C, 32), N), dim3(32, 8), 0, s>>>(partial, N, C, chunks, sums);
    RD_CHECK_LAUNCH(ctx, "spade_bwd_finalize");
    if (C % VecIO<T>::V == 0)
      k_spade_bwd_apply_vec<T><<<rd_grid_1d(total / VecIO<T>::V, 256, ctx->sm_count), 256, 0, s>>>(
          (const T*)z, mean, invstd, (const T*)gb, (const T*)dmix, sums, (T*)dz, hw, C, total / VecIO<T>::V, gs);
    else
      k_spade_bwd_apply<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, s>>>((const T*)z, mean, invstd, (const T*)gb,
                                                                                    (const T*)dmix, sums, (T*)dz, hw, C, total, gs);
    RD_CHECK_LAUNCH(ctx, "spade_bwd_apply");
  });
  return RD_OK;
}

// ============================================================================ bilinear resize
// PyTorch upsample_bilinear2d source-index rule (aten/native/UpSample.h area_pixel_compute_source_index):
// align_corners: src = dst*(in-1)/(out-1) ; else src = max((dst+0.5)*in/out - 0.5, 0).
struct BilinCoord { int i0, i1; float l1; };
__device__ __forceinline__ BilinCoord bilin_coord(int dst, int in, int out, int align) {
  float src;
  if (align) {
    float scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = scale * dst;
  } else {
    float scale = (float)in / (float)out;
    src = scale * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  BilinCoord r;
  r.i0 = (int)src;
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + ((r.i0 < in - 1) ? 1 : 0);
  r.l1 = src - (float)r.i0;
  return r;
}
// V channels per thread: 16-byte vectors when the channel count allows (8 bf16 / 4 fp32), else 4-wide, else scalar
template <typename T, int V> struct VecN;
template <typename T> struct VecN<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[1]) { v[0] = ldf<T>(p); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[1]) { stf<T>(p, v[0]); }
};
template <typename T> struct VecN<T, 4> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[4]) { Vec4<T>::load(p, v); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[4]) { Vec4<T>::store(p, v); }
};
template <> struct VecN<bf16, 8> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) { VecIO<bf16>::load(p, v); }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) { VecIO<bf16>::store(p, v); }
};

// grid (x chunks, output rows, images): the row coordinates are block-uniform, no 64-bit div / mod per element
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_fwd(const T* __restrict__ x, T* __restrict__ y, int h, int w, int c, int oh, int ow,
                                                      int align) {
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ow * cv) return;
  const int ox = i / cv, ch = (i - ox * cv) * V;
  const int oy = blockIdx.y, img = blockIdx.z;
  const BilinCoord cy = bilin_coord(oy, h, oh, align), cx = bilin_coord(ox, w, ow, align);
  const T* base = x + (int64_t)img * h * w * c + ch;
  float a[V], b[V], cc[V], d[V], o[V];
  VecN<T, V>::load(base + ((int64_t)cy.i0 * w + cx.i0) * c, a);
  VecN<T, V>::load(base + ((int64_t)cy.i0 * w + cx.i1) * c, b);
  VecN<T, V>::load(base + ((int64_t)cy.i1 * w + cx.i0) * c, cc);
  VecN<T, V>::load(base + ((int64_t)cy.i1 * w + cx.i1) * c, d);
  const float ly1 = cy.l1, ly0 = 1.f - ly1, lx1 = cx.l1, lx0 = 1.f - lx1;
#pragma unroll
  for (int k = 0; k < V; ++k) o[k] = ly0 * (lx0 * a[k] + lx1 * b[k]) + ly1 * (lx0 * cc[k] + lx1 * d[k]);
  VecN<T, V>::store(y + (((int64_t)img * oh + oy) * ow + ox) * c + ch, o);
}
// R output rows per thread (same column and channel vector): all 4 R loads are issued before the first use.  One vector per thread
// left the kernel latency bound (a thread lives ~1 us for 16 bytes of output: 1.3 TB/s); same expression per element.
template <int R>
__global__ void __launch_bounds__(256, 2) k_bilinear_fwd_rows(const bf16* __restrict__ x, bf16* __restrict__ y, int h, int w, int c, int oh, int ow,
                                                              int align) {
  constexpr int V = 8;
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ow * cv) return;
  const int ox = i / cv, ch = (i - ox * cv) * V;
  const int oy0 = blockIdx.y * R, img = blockIdx.z;
  const BilinCoord cx = bilin_coord(ox, w, ow, align);
  const bf16* base = x + (int64_t)img * h * w * c + ch;
  const float lx1 = cx.l1, lx0 = 1.f - lx1;
  uint4 ra[R], rb[R], rc[R], rd[R];        // raw 16-byte vectors: converted at use, 16 registers per row
  float ly1[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int oy = oy0 + r < oh ? oy0 + r : oh - 1;
    const BilinCoord cy = bilin_coord(oy, h, oh, align);
    ly1[r] = cy.l1;
    ra[r] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)cy.i0 * w + cx.i0) * c));
    rb[r] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)cy.i0 * w + cx.i1) * c));
    rc[r] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)cy.i1 * w + cx.i0) * c));
    rd[r] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)cy.i1 * w + cx.i1) * c));
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (oy0 + r < oh) {
      const float l1 = ly1[r], l0 = 1.f - l1;
      const uint32_t wa[4] = {ra[r].x, ra[r].y, ra[r].z, ra[r].w}, wb[4] = {rb[r].x, rb[r].y, rb[r].z, rb[r].w};
      const uint32_t wc[4] = {rc[r].x, rc[r].y, rc[r].z, rc[r].w}, wd[4] = {rd[r].x, rd[r].y, rd[r].z, rd[r].w};
      float o[V];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // bf16 -> fp32 is a 16-bit shift; low half = even channel
        const float a0 = __uint_as_float(wa[q] << 16), a1 = __uint_as_float(wa[q] & 0xffff0000u);
        const float b0 = __uint_as_float(wb[q] << 16), b1 = __uint_as_float(wb[q] & 0xffff0000u);
        const float c0 = __uint_as_float(wc[q] << 16), c1 = __uint_as_float(wc[q] & 0xffff0000u);
        const float d0 = __uint_as_float(wd[q] << 16), d1 = __uint_as_float(wd[q] & 0xffff0000u);
        o[2 * q] = l0 * (lx0 * a0 + lx1 * b0) + l1 * (lx0 * c0 + lx1 * d0);
        o[2 * q + 1] = l0 * (lx0 * a1 + lx1 * b1) + l1 * (lx0 * c1 + lx1 * d1);
      }
      VecIO<bf16>::store(y + (((int64_t)img * oh + oy0 + r) * ow + ox) * c + ch, o);
    }
  }
}
// Exact x2, align_corners = False specialisation (nn.Upsample(scale_factor=2) between the SPADE blocks, src/model.py:2501): one thread
// per INPUT pixel and channel vector writes its 2 x 2 output pixels from the clamped 3 x 3 neighbourhood — out(2i) = 1/4 x[i-1] +
// 3/4 x[i], out(2i+1) = 3/4 x[i] + 1/4 x[i+1] per axis, the same weights and the same expression as the generic kernel — with
// 9 loads per 4 stores instead of 16 and no coordinate arithmetic.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_fwd_x2(const T* __restrict__ x, T* __restrict__ y, int h, int w, int c) {
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) * V;
  const int iy = blockIdx.y, img = blockIdx.z;
  const int ym = iy > 0 ? iy - 1 : 0, yp = iy < h - 1 ? iy + 1 : h - 1;
  const int xm = ix > 0 ? ix - 1 : 0, xp = ix < w - 1 ? ix + 1 : w - 1;
  const T* base = x + (int64_t)img * h * w * c + ch;
  float v[3][3][V];
  const int ys[3] = {ym, iy, yp}, xs[3] = {xm, ix, xp};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) VecN<T, V>::load(base + ((int64_t)ys[a] * w + xs[b]) * c, v[a][b]);
  // PyTorch clamps the source coordinate at 0 (l1 = 0 there): the first output row / column copies pixel 0 exactly
  const float l1y[2] = {iy > 0 ? 0.75f : 0.f, 0.25f}, l1x[2] = {ix > 0 ? 0.75f : 0.f, 0.25f};
  const int ow = 2 * w;
  T* out = y + (((int64_t)img * 2 * h + 2 * iy) * ow + 2 * ix) * c + ch;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      // stencil rows (i0, i1) = (i-1, i) for the even output, (i, i+1) for the odd one
      const int a0 = dy, b0 = dx;
      const float ly1 = l1y[dy], ly0 = 1.f - ly1, lx1 = l1x[dx], lx0 = 1.f - lx1;
      float o[V];
#pragma unroll
      for (int k = 0; k < V; ++k)
        o[k] = ly0 * (lx0 * v[a0][b0][k] + lx1 * v[a0][b0 + 1][k]) + ly1 * (lx0 * v[a0 + 1][b0][k] + lx1 * v[a0 + 1][b0 + 1][k]);
      VecN<T, V>::store(out + ((int64_t)dy * ow + dx) * c, o);
    }
  }
}
// Up-sampling forward, one thread per (output column, 8-channel vector) walking kBilRows consecutive output rows.  The row coordinates
// of the block and the column coordinates of its threads are evaluated once per block into shared memory (the per-thread float
// divisions and coordinate arithmetic were the bulk of the 815 instructions per thread of k_bilinear_fwd_rows, ncu: issue slots 68 %
// busy).  The horizontal blend h(row) = lx0 * x[row][i0] + lx1 * x[row][i1] of an INPUT row is kept in registers and reused by all
// output rows that read it (two per input row when up-sampling x2; the row tests are block-uniform), so an output row costs one
// vertical blend, and every second one a new pair of loads.  Same expression, same rounding as the generic kernel:
// out = l0 * (lx0 a + lx1 b) + l1 * (lx0 c + lx1 d).
constexpr int kBilRows = 16;
__global__ void __launch_bounds__(256) k_bilinear_fwd_roll(const bf16* __restrict__ x, bf16* __restrict__ y, int h, int w, int c, int oh, int ow,
                                                           int align) {
  __shared__ int s_y0[kBilRows], s_y1[kBilRows], s_x0[256], s_x1[256];
  __shared__ float s_ly[kBilRows], s_lx[256];
  const int cv = c >> 3;
  const int i0 = blockIdx.x * blockDim.x;                 // first (column, vector) item of the block
  const int oy0 = blockIdx.y * kBilRows, img = blockIdx.z;
  const int ox_first = i0 / cv;
  if (threadIdx.x < kBilRows) {
    const int oy = oy0 + threadIdx.x < oh ? oy0 + threadIdx.x : oh - 1;
    const BilinCoord cy = bilin_coord(oy, h, oh, align);
    s_y0[threadIdx.x] = cy.i0; s_y1[threadIdx.x] = cy.i1; s_ly[threadIdx.x] = cy.l1;
  }
  {
    const int ox = ox_first + threadIdx.x;                // a block spans at most 256 / cv + 1 <= 256 columns
    if (ox < ow && (int)threadIdx.x <= 255 / cv + 1) {
      const BilinCoord cx = bilin_coord(ox, w, ow, align);
      s_x0[threadIdx.x] = cx.i0; s_x1[threadIdx.x] = cx.i1; s_lx[threadIdx.x] = cx.l1;
    }
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= ow * cv) return;
  const int ox = i / cv, ch = (i - ox * cv) << 3, lxi = ox - ox_first;
  const float lx1 = s_lx[lxi], lx0 = 1.f - lx1;
  const bf16* base = x + (int64_t)img * h * w * c + ch;
  const int64_t xa = (int64_t)s_x0[lxi] * c, xb = (int64_t)s_x1[lxi] * c;
  bf16* out = y + (((int64_t)img * oh + oy0) * ow + ox) * c + ch;
  float hA[8], hB[8];
  int rowA = -1, rowB = -1;
  const int nrows = oh - oy0 < kBilRows ? oh - oy0 : kBilRows;
  auto hblend = [&](int row, float (&hh)[8]) {
    const bf16* r = base + (int64_t)row * w * c;
    const uint4 va = __ldg(reinterpret_cast<const uint4*>(r + xa)), vb = __ldg(reinterpret_cast<const uint4*>(r + xb));
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      hh[2 * q] = lx0 * __uint_as_float(wa[q] << 16) + lx1 * __uint_as_float(wb[q] << 16);
      hh[2 * q + 1] = lx0 * __uint_as_float(wa[q] & 0xffff0000u) + lx1 * __uint_as_float(wb[q] & 0xffff0000u);
    }
  };
  for (int r = 0; r < nrows; ++r) {
    const int y0 = s_y0[r], y1 = s_y1[r];
    if (y0 != rowA) {                                     // block-uniform
      if (y0 == rowB) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hA[k] = hB[k];
      } else hblend(y0, hA);
      rowA = y0;
      rowB = -1;
    }
    if (y1 != rowB) {
      if (y1 == rowA) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hB[k] = hA[k];
      } else hblend(y1, hB);
      rowB = y1;
    }
    const float l1 = s_ly[r], l0 = 1.f - l1;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = l0 * hA[k] + l1 * hB[k];
    VecIO<bf16>::store(out + (int64_t)r * ow * c, o);
  }
}
template <typename T, int V>
static inline void launch_bilinear_fwd(const void* x, void* y, int n, int h, int w, int c, int oh, int ow, int align, cudaStream_t s) {
  dim3 grid(rd_div_up((int64_t)ow * (c / V), 256), oh, n);
  k_bilinear_fwd<T, V><<<grid, 256, 0, s>>>((const T*)x, (T*)y, h, w, c, oh, ow, align);
}
extern "C" int rd_bilinear_fwd(rd_ctx* ctx, const void* x, void* y, int n, int h, int w, int c, int oh, int ow, int align,
                               int dtype, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  if (oh > 65535 || n > 65535) RD_FAIL(ctx, RD_ERR_ARG, "bilinear: oh and n must be <= 65535");
  static const int roll = getenv("RD_B200_BILINEAR_ROLL") ? atoi(getenv("RD_B200_BILINEAR_ROLL")) : 3;     // bit 0: x2 / align False, bit 1: the other up-samplings
  const bool is_x2 = !align && oh == 2 * h && ow == 2 * w && h > 1 && w > 1;
  if (dtype == RD_BF16 && c % 8 == 0 && c <= 2048 && oh >= h && ow >= w && oh >= 16 && ((is_x2 && (roll & 1)) || (!is_x2 && (roll & 2)))) {
    dim3 grid(rd_div_up((int64_t)ow * (c / 8), 256), rd_div_up(oh, kBilRows), n);
    k_bilinear_fwd_roll<<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, h, w, c, oh, ow, align);
  }
  else if (dtype == RD_BF16 && c % 8 == 0 && is_x2 && h <= 65535) {
    dim3 grid(rd_div_up((int64_t)w * (c / 8), 256), h, n);
    k_bilinear_fwd_x2<bf16, 8><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, h, w, c);
  }
  else if (dtype == RD_BF16 && c % 8 == 0 && oh >= 16) {
    dim3 grid(rd_div_up((int64_t)ow * (c / 8), 256), rd_div_up(oh, 4), n);
    k_bilinear_fwd_rows<4><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, h, w, c, oh, ow, align);
  }
  else if (dtype == RD_BF16 && c % 8 == 0) launch_bilinear_fwd<bf16, 8>(x, y, n, h, w, c, oh, ow, align, s);
  else if (c % 4 == 0) { RD_DISPATCH_DTYPE(dtype, (launch_bilinear_fwd<T, 4>(x, y, n, h, w, c, oh, ow, align, s))); }
  else { RD_DISPATCH_DTYPE(dtype, (launch_bilinear_fwd<T, 1>(x, y, n, h, w, c, oh, ow, align, s))); }
  RD_CHECK_LAUNCH(ctx, "bilinear_fwd");
  return RD_OK;
}
// backward as a gather (deterministic, no atomics): each input pixel sums the output pixels whose
// forward stencil touched it, recomputing the forward coordinates exactly.
__device__ __forceinline__ void bilin_range(int i, int in, int out, int align, int& lo, int& hi) {
  float scale, c0;
  if (align) { scale = (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f; c0 = 0.f; }
  else { scale = (float)in / (float)out; c0 = 0.5f * scale - 0.5f; }
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  // src(dst) = scale*dst + c0 ; stencil {floor(src), floor(src)+1} contains i  <=>  src in (i-1, i+1)
  float flo = ((float)(i - 1) - c0) / scale, fhi = ((float)(i + 1) - c0) / scale;
  lo = (int)floorf(flo) - 1;
  hi = (int)ceilf(fhi) + 1;
  if (i == 0) lo = 0;            // clamped sources (src < 0 -> 0) all land on row 0
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}
// grid (x chunks, input rows, images).  The stencil weights separate: w(oy, ox) = wy(oy) * wx(ox); each thread first
// collects its (<= 6) contributing output columns and weights, the rows are walked with their weight computed once per
// row — instead of two coordinate evaluations per tap.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_bwd(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int c, int oh, int ow,
                                                      int align) {
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) * V;
  const int iy = blockIdx.y, img = blockIdx.z;
  int ylo, yhi, xlo, xhi;
  bilin_range(iy, h, oh, align, ylo, yhi);
  bilin_range(ix, w, ow, align, xlo, xhi);
  constexpr int kMaxTaps = 6;           // x2 upsampling touches <= 4 columns; wider ranges fall back to the generic walk
  int xs[kMaxTaps];
  float wxs[kMaxTaps];
  int nx = 0;
  bool overflow = false;
  for (int ox = xlo; ox <= xhi; ++ox) {
    const BilinCoord cx = bilin_coord(ox, w, ow, align);
    float wx = 0.f;
    if (cx.i0 == ix) wx += 1.f - cx.l1;
    if (cx.i1 == ix) wx += cx.l1;
    if (wx == 0.f) continue;
    if (nx < kMaxTaps) { xs[nx] = ox; wxs[nx] = wx; ++nx; } else overflow = true;
  }
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  const T* base = dy + (int64_t)img * oh * ow * c + ch;
  for (int oy = ylo; oy <= yhi; ++oy) {
    const BilinCoord cy = bilin_coord(oy, h, oh, align);
    float wy = 0.f;
    if (cy.i0 == iy) wy += 1.f - cy.l1;
    if (cy.i1 == iy) wy += cy.l1;
    if (wy == 0.f) continue;
    const T* row = base + (int64_t)oy * ow * c;
    if (!overflow) {
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t) {
        if (t < nx) {
          float v[V];
          VecN<T, V>::load(row + (int64_t)xs[t] * c, v);
          const float ww = wy * wxs[t];
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] += ww * v[k];
        }
      }
    } else {                              // strong down-scaling (resize of s to the coarse SPADE scales): generic walk
      for (int ox = xlo; ox <= xhi; ++ox) {
        const BilinCoord cx = bilin_coord(ox, w, ow, align);
        float wx = 0.f;
        if (cx.i0 == ix) wx += 1.f - cx.l1;
        if (cx.i1 == ix) wx += cx.l1;
        if (wx == 0.f) continue;
        float v[V];
        VecN<T, V>::load(row + (int64_t)ox * c, v);
        const float ww = wy * wx;
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += ww * v[k];
      }
    }
  }
  VecN<T, V>::store(dx + (((int64_t)img * h + iy) * w + ix) * c + ch, acc);
}
// Exact x2, align_corners = False specialisation (nn.Upsample(scale_factor=2) between the SPADE blocks, src/model.py:2501):
// low-resolution pixel i receives the high-resolution pixels 2i-1 .. 2i+2 with weights 1/4, 3/4, 3/4, 1/4; at the borders the
// clamped source index folds the missing neighbour into the edge pixel (weight 1 instead of 3/4).  Same values as the generic
// gather above, without any coordinate arithmetic per tap.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_bwd_x2(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int c) {
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) * V;
  const int iy = blockIdx.y, img = blockIdx.z;
  const int oh = 2 * h, ow = 2 * w;
  float wy[4], wx[4];
  wy[0] = iy > 0 ? 0.25f : 0.f;  wy[1] = iy == 0 ? 1.f : 0.75f;  wy[2] = iy == h - 1 ? 1.f : 0.75f;  wy[3] = iy < h - 1 ? 0.25f : 0.f;
  wx[0] = ix > 0 ? 0.25f : 0.f;  wx[1] = ix == 0 ? 1.f : 0.75f;  wx[2] = ix == w - 1 ? 1.f : 0.75f;  wx[3] = ix < w - 1 ? 0.25f : 0.f;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  const T* base = dy + (int64_t)img * oh * ow * c + ch;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int oy = 2 * iy - 1 + a;
    if (wy[a] == 0.f) continue;
    const T* row = base + (int64_t)oy * ow * c;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] == 0.f) continue;
      float v[V];
      VecN<T, V>::load(row + (int64_t)(2 * ix - 1 + b) * c, v);
      const float ww = wy[a] * wx[b];
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += ww * v[k];
    }
  }
  VecN<T, V>::store(dx + (((int64_t)img * h + iy) * w + ix) * c + ch, acc);
}

// The same x2 backward with the separable stencil evaluated row-wise: a thread walks kBilRows input rows of one (column, 8-channel
// vector); the horizontal gathers g(oy) = sum_b wx[b] dy[oy, 2 ix - 1 + b] of the two output rows an input row shares with its successor
// stay in registers, so every input row costs two new gathers (8 loads) instead of 16 loads and 16 weighted adds.
__global__ void __launch_bounds__(256) k_bilinear_bwd_x2_roll(const bf16* __restrict__ dy, bf16* __restrict__ dx, int h, int w, int c) {
  const int cv = c >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) << 3;
  const int iy0 = blockIdx.y * kBilRows, img = blockIdx.z;
  const int oh = 2 * h, ow = 2 * w;
  float wx[4];
  wx[0] = ix > 0 ? 0.25f : 0.f;  wx[1] = ix == 0 ? 1.f : 0.75f;  wx[2] = ix == w - 1 ? 1.f : 0.75f;  wx[3] = ix < w - 1 ? 0.25f : 0.f;
  const bf16* base = dy + (int64_t)img * oh * ow * c + (int64_t)(2 * ix - 1) * c + ch;
  auto gather = [&](int oy, float (&g)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0.f;
    const bf16* row = base + (int64_t)oy * ow * c;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] != 0.f) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + (int64_t)b * c));
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          g[2 * q] += wx[b] * __uint_as_float(wv[q] << 16);
          g[2 * q + 1] += wx[b] * __uint_as_float(wv[q] & 0xffff0000u);
        }
      }
    }
  };
  float gA[8], gB[8], gC[8], gD[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) gA[k] = gD[k] = 0.f;
  if (iy0 > 0) gather(2 * iy0 - 1, gA);
  gather(2 * iy0, gB);
  const int nrows = h - iy0 < kBilRows ? h - iy0 : kBilRows;
  bf16* out = dx + (((int64_t)img * h + iy0) * w + ix) * c + ch;
  for (int r = 0; r < nrows; ++r) {
    const int iy = iy0 + r;
    gather(2 * iy + 1, gC);
    const bool last = iy == h - 1;
    if (!last) gather(2 * iy + 2, gD);
    const float w0 = iy > 0 ? 0.25f : 0.f, w1 = iy == 0 ? 1.f : 0.75f, w2 = last ? 1.f : 0.75f, w3 = last ? 0.f : 0.25f;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = w0 * gA[k] + w1 * gB[k] + w2 * gC[k] + w3 * gD[k];
    VecIO<bf16>::store(out + (int64_t)r * w * c, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) { gA[k] = gC[k]; gB[k] = gD[k]; }
  }
}

// Up-scaling backward (<= 6 contributing output rows / columns per input pixel, e.g. the x2 align_corners = True upsample of the
// anatomy decoder): the tap lists are block-level data — the row's taps are the same for every thread and a block sees at most
// 256 / cv + 1 distinct columns — so they are built ONCE per block in shared memory instead of ~10 coordinate evaluations (each
// with a float division) per thread.  Same weights, same (row-major) accumulation order as k_bilinear_bwd.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_bwd_tab(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int c, int oh, int ow,
                                                          int align) {
  constexpr int kMaxTaps = 6;
  __shared__ int s_ny, s_oy[kMaxTaps], s_nx[257], s_ox[257][kMaxTaps];
  __shared__ float s_wy[kMaxTaps], s_wx[257][kMaxTaps];
  const int cv = c / V;
  const int i0 = blockIdx.x * blockDim.x;
  const int iy = blockIdx.y, img = blockIdx.z;
  const int ix_first = i0 / cv;
  int ix_last = (i0 + (int)blockDim.x - 1) / cv;
  if (ix_last > w - 1) ix_last = w - 1;
  if (threadIdx.x == 0) {
    int lo, hi, n = 0;
    bilin_range(iy, h, oh, align, lo, hi);
    for (int oy = lo; oy <= hi; ++oy) {
      const BilinCoord cy = bilin_coord(oy, h, oh, align);
      float wy = 0.f;
      if (cy.i0 == iy) wy += 1.f - cy.l1;
      if (cy.i1 == iy) wy += cy.l1;
      if (wy != 0.f && n < kMaxTaps) { s_oy[n] = oy; s_wy[n] = wy; ++n; }
    }
    s_ny = n;
  }
  for (int l = threadIdx.x; ix_first + l <= ix_last; l += blockDim.x) {
    const int ix = ix_first + l;
    int lo, hi, n = 0;
    bilin_range(ix, w, ow, align, lo, hi);
    for (int ox = lo; ox <= hi; ++ox) {
      const BilinCoord cx = bilin_coord(ox, w, ow, align);
      float wx = 0.f;
      if (cx.i0 == ix) wx += 1.f - cx.l1;
      if (cx.i1 == ix) wx += cx.l1;
      if (wx != 0.f && n < kMaxTaps) { s_ox[l][n] = ox; s_wx[l][n] = wx; ++n; }
    }
    s_nx[l] = n;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) * V, l = ix - ix_first;
  const int ny = s_ny, nx = s_nx[l];
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  const T* base = dy + (int64_t)img * oh * ow * c + ch;
  for (int a = 0; a < ny; ++a) {
    const T* row = base + (int64_t)s_oy[a] * ow * c;
    const float wy = s_wy[a];
#pragma unroll
    for (int t = 0; t < kMaxTaps; ++t) {
      if (t < nx) {
        float v[V];
        VecN<T, V>::load(row + (int64_t)s_ox[l][t] * c, v);
        const float ww = wy * s_wx[l][t];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += ww * v[k];
      }
    }
  }
  VecN<T, V>::store(dx + (((int64_t)img * h + iy) * w + ix) * c + ch, acc);
}

// Backward of an integer power-of-two DOWN-scaling with align_corners = False (the resize of the anatomy code to the coarse SPADE
// scales, src/model.py:2441): src = f dst + f/2 - 1/2, so every output pixel averages the 2 x 2 block at (f oy + f/2 - 1, f ox + f/2 - 1)
// with weights 1/2 x 1/2.  An input pixel receives 0.25 dy[oy, ox] when its row and column are one of the two central ones of their
// f-block and nothing otherwise — the same single product the generic gather ends with, without its range searches and divisions.
template <typename T, int V>
__global__ void __launch_bounds__(256) k_bilinear_bwd_down(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int c, int oh, int ow,
                                                           int shift) {
  const int cv = c / V;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * cv) return;
  const int ix = i / cv, ch = (i - ix * cv) * V;
  const int iy = blockIdx.y, img = blockIdx.z;
  const int f = 1 << shift, half = f >> 1;
  const int ry = iy & (f - 1), rx = ix & (f - 1);
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  if ((ry == half - 1 || ry == half) && (rx == half - 1 || rx == half)) {
    float v[V];
    VecN<T, V>::load(dy + (((int64_t)img * oh + (iy >> shift)) * ow + (ix >> shift)) * c + ch, v);
    const float ww = 0.5f * 0.5f;
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += ww * v[k];
  }
  VecN<T, V>::store(dx + (((int64_t)img * h + iy) * w + ix) * c + ch, acc);
}
static inline int rd_pow2_downscale(int h, int w, int oh, int ow) {     // log2 f when (h, w) = f (oh, ow), f = 2^k >= 2; else 0
  if (oh < 1 || ow < 1 || h % oh || w % ow || h / oh != w / ow) return 0;
  const int f = h / oh;
  if (f < 2 || (f & (f - 1))) return 0;
  int s = 0;
  while ((1 << s) < f) ++s;
  return s;
}

template <typename T, int V>
static inline void launch_bilinear_bwd(const void* dy, void* dx, int n, int h, int w, int c, int oh, int ow, int align, cudaStream_t s) {
  dim3 grid(rd_div_up((int64_t)w * (c / V), 256), h, n);
  k_bilinear_bwd<T, V><<<grid, 256, 0, s>>>((const T*)dy, (T*)dx, h, w, c, oh, ow, align);
}
extern "C" int rd_bilinear_bwd(rd_ctx* ctx, const void* dy, void* dx, int n, int h, int w, int c, int oh, int ow, int align,
                               int dtype, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  if (h > 65535 || n > 65535) RD_FAIL(ctx, RD_ERR_ARG, "bilinear: h and n must be <= 65535");
  static const int roll = getenv("RD_B200_BILINEAR_ROLL") ? atoi(getenv("RD_B200_BILINEAR_ROLL")) : 3;     // bit 2: the x2 backward row walk (measured: no gain over k_bilinear_bwd_x2, off by default)
  if (dtype == RD_BF16 && c % 8 == 0 && !align && oh == 2 * h && ow == 2 * w && h >= 8 && w > 1 && (roll & 4)) {
    dim3 grid(rd_div_up((int64_t)w * (c / 8), 256), rd_div_up(h, kBilRows), n);
    k_bilinear_bwd_x2_roll<<<grid, 256, 0, s>>>((const bf16*)dy, (bf16*)dx, h, w, c);
  }
  else if (dtype == RD_BF16 && c % 8 == 0 && !align && oh == 2 * h && ow == 2 * w && h > 1 && w > 1) {
    dim3 grid(rd_div_up((int64_t)w * (c / 8), 256), h, n);
    k_bilinear_bwd_x2<bf16, 8><<<grid, 256, 0, s>>>((const bf16*)dy, (bf16*)dx, h, w, c);
  }
  else if (!align && c % 4 == 0 && rd_pow2_downscale(h, w, oh, ow) > 0) {
    const int shift = rd_pow2_downscale(h, w, oh, ow);
    dim3 grid(rd_div_up((int64_t)w * (c / 4), 256), h, n);
    RD_DISPATCH_DTYPE(dtype, (k_bilinear_bwd_down<T, 4><<<grid, 256, 0, s>>>((const T*)dy, (T*)dx, h, w, c, oh, ow, shift)));
  }
  else if (dtype == RD_BF16 && c % 8 == 0 && oh >= h && ow >= w && oh <= 2 * h && ow <= 2 * w) {     // <= 4-5 taps per axis
    dim3 grid(rd_div_up((int64_t)w * (c / 8), 256), h, n);
    k_bilinear_bwd_tab<bf16, 8><<<grid, 256, 0, s>>>((const bf16*)dy, (bf16*)dx, h, w, c, oh, ow, align);
  }
  else if (dtype == RD_BF16 && c % 8 == 0) launch_bilinear_bwd<bf16, 8>(dy, dx, n, h, w, c, oh, ow, align, s);
  else if (c % 4 == 0) { RD_DISPATCH_DTYPE(dtype, (launch_bilinear_bwd<T, 4>(dy, dx, n, h, w, c, oh, ow, align, s))); }
  else { RD_DISPATCH_DTYPE(dtype, (launch_bilinear_bwd<T, 1>(dy, dx, n, h, w, c, oh, ow, align, s))); }
  RD_CHECK_LAUNCH(ctx, "bilinear_bwd");
  return RD_OK;
}

// ============================================================================ activations
template <typename T>
__global__ void k_lrelu_fwd(const T* __restrict__ x, T* __restrict__ y, int64_t n, float slope) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(x + i);
    stf<T>(y + i, v > 0.f ? v : v * slope);
  }
}
template <typename T>
__global__ void k_lrelu_bwd(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n, float slope) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = ldf<T>(dy + i);
    stf<T>(dx + i, ldf<T>(y + i) > 0.f ? d : d * slope);
  }
}
extern "C" int rd_lrelu_fwd(rd_ctx* ctx, const void* x, void* y, int64_t n, float slope, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, (k_lrelu_fwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)x, (T*)y, n, slope)));
  RD_CHECK_LAUNCH(ctx, "lrelu_fwd");
  return RD_OK;
}
extern "C" int rd_lrelu_bwd(rd_ctx* ctx, const void* dy, const void* y, void* dx, int64_t n, float slope, int dtype,
                            rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, (k_lrelu_bwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)dy, (const T*)y, (T*)dx, n, slope)));
  RD_CHECK_LAUNCH(ctx, "lrelu_bwd");
  return RD_OK;
}

// masked softmax: one thread per pixel, C <= 16 channels in registers
// F.softplus (beta 1, threshold 20): the decoder / anatomy output activation of the mean-normalised datasets
// (src/main_missing.py:75-86, src/model.py:3145-3146)
template <typename T>
__global__ void k_softplus_fwd(const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(x + i);
    stf<T>(y + i, v > 20.f ? v : log1pf(expf(v)));
  }
}
template <typename T>
__global__ void k_softplus_bwd(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(x + i);
    float g = v > 20.f ? 1.f : 1.f / (1.f + expf(-v));
    stf<T>(dx + i, ldf<T>(dy + i) * g);
  }
}
extern "C" int rd_softplus_fwd(rd_ctx* ctx, const void* x, void* y, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_softplus_fwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)x, (T*)y, n));
  RD_CHECK_LAUNCH(ctx, "softplus_fwd");
  return RD_OK;
}
extern "C" int rd_softplus_bwd(rd_ctx* ctx, const void* dy, const void* x, void* dx, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_softplus_bwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)dy, (const T*)x, (T*)dx, n));
  RD_CHECK_LAUNCH(ctx, "softplus_bwd");
  return RD_OK;
}

template <typename T>
__global__ void k_msoftmax_fwd(const T* __restrict__ s, const float* __restrict__ mask, T* __restrict__ p, int64_t pixels,
                               int C, int64_t mask_pixels) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    float v[16];
    float ml = mask ? 100.f * mask[i % mask_pixels] : -INFINITY;
    float mx = ml;
    for (int c = 0; c < C; ++c) { v[c] = ldf<T>(s + i * C + c); mx = fmaxf(mx, v[c]); }
    float sum = mask ? expf(ml - mx) : 0.f;
    for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    float inv = 1.f / sum;
    for (int c = 0; c < C; ++c) stf<T>(p + i * C + c, v[c] * inv);
  }
}
template <typename T>
__global__ void k_msoftmax_bwd(const T* __restrict__ p, const T* __restrict__ dp, T* __restrict__ ds, int64_t pixels, int C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    float pv[16], dv[16];
    float dot = 0.f;   // the dropped (mask) channel has zero upstream gradient
    for (int c = 0; c < C; ++c) { pv[c] = ldf<T>(p + i * C + c); dv[c] = ldf<T>(dp + i * C + c); dot += pv[c] * dv[c]; }
    for (int c = 0; c < C; ++c) stf<T>(ds + i * C + c, pv[c] * (dv[c] - dot));
  }
}
extern "C" int rd_masked_softmax_fwd(rd_ctx* ctx, const void* s, const float* mask, int64_t mask_pixels, void* p,
                                     int64_t pixels, int C, int dtype, rd_stream st) {
  if (C > 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "masked_softmax: C <= 16");
  if (mask && mask_pixels <= 0) RD_FAIL(ctx, RD_ERR_ARG, "masked_softmax: mask_pixels must be > 0");
  RD_DISPATCH_DTYPE(dtype, (k_msoftmax_fwd<T><<<rd_grid_1d(pixels, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)s, mask, (T*)p, pixels, C, mask_pixels)));
  RD_CHECK_LAUNCH(ctx, "masked_softmax_fwd");
  return RD_OK;
}
extern "C" int rd_masked_softmax_bwd(rd_ctx* ctx, const void* p, const void* dp, void* ds, int64_t pixels, int C, int dtype,
                                     rd_stream st) {
  if (C > 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "masked_softmax: C <= 16");
  RD_DISPATCH_DTYPE(dtype, (k_msoftmax_bwd<T><<<rd_grid_1d(pixels, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)p, (const T*)dp, (T*)ds, pixels, C)));
  RD_CHECK_LAUNCH(ctx, "masked_softmax_bwd");
  return RD_OK;
}

// ============================================================================ small linears (fp32)
// one warp per output element; lanes stride over the reduction dimension
__global__ void k_linear_fwd(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                             float* __restrict__ y, int rows, int in_f, int out_f, int act, float slope) {
  int warps_per_block = blockDim.x >> 5;
  int64_t total = (int64_t)rows * out_f;
  int lane = threadIdx.x & 31;
  for (int64_t o = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); o < total;
       o += (int64_t)gridDim.x * warps_per_block) {
    int r = (int)(o / out_f), j = (int)(o - (int64_t)r * out_f);
    const float* xr = x + (int64_t)r * in_f;
    const float* wr = W + (int64_t)j * in_f;
    float acc = 0.f;
    for (int k = lane; k < in_f; k += 32) acc += xr[k] * wr[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + (b ? b[j] : 0.f);
      if (act == RD_ACT_LRELU) v = v > 0.f ? v : v * slope;
      y[o] = v;
    }
  }
}
// short reduction (zi_scaler 16 -> 3840): one THREAD per output element — a warp per element leaves half its lanes idle and pays five
// shuffles for 16 products (135 us for 64 x 3840 outputs)
__global__ void k_linear_fwd_small(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                   float* __restrict__ y, int rows, int in_f, int out_f, int act, float slope) {
  const int64_t total = (int64_t)rows * out_f;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(o / out_f), j = (int)(o - (int64_t)r * out_f);
    const float* xr = x + (int64_t)r * in_f;
    const float* wr = W + (int64_t)j * in_f;
    float acc = 0.f;
    for (int k = 0; k < in_f; ++k) acc = fmaf(__ldg(xr + k), __ldg(wr + k), acc);
    float v = acc + (b ? b[j] : 0.f);
    if (act == RD_ACT_LRELU) v = v > 0.f ? v : v * slope;
    y[o] = v;
  }
}
extern "C" int rd_linear_fwd(rd_ctx* ctx, const float* x, const float* W, const float* b, float* y, int rows, int in_f,
                             int out_f, int act, float slope, rd_stream st) {
  int64_t total = (int64_t)rows * out_f;
  if (in_f <= 32 && total >= 4096) {
    k_linear_fwd_small<<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(x, W, b, y, rows, in_f, out_f, act, slope);
    RD_CHECK_LAUNCH(ctx, "linear_fwd_small");
    return RD_OK;
  }
  int grid = rd_grid_1d(total, 8, ctx->sm_count);
  k_linear_fwd<<<grid, 256, 0, (cudaStream_t)st>>>(x, W, b, y, rows, in_f, out_f, act, slope);
  RD_CHECK_LAUNCH(ctx, "linear_fwd");
  return RD_OK;
}
// dx[r,k] = sum_j dy[r,j] W[j,k]  (thread per element, coalesced over k)
__global__ void k_linear_dx(const float* __restrict__ dy, const float* __restrict__ W, float* __restrict__ dx, int rows,
                            int in_f, int out_f) {
  int64_t total = (int64_t)rows * in_f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / in_f), k = (int)(i - (int64_t)r * in_f);
    float acc = 0.f;
    for (int j = 0; j < out_f; ++j) acc += dy[(int64_t)r * out_f + j] * W[(int64_t)j * in_f + k];
    dx[i] = acc;
  }
}
// dW[j,k] += sum_r dy[r,j] x[r,k]; db[j] += sum_r dy[r,j]
__global__ void k_linear_dw(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                            float* __restrict__ db, int rows, int in_f, int out_f) {
  int64_t total = (int64_t)out_f * in_f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i / in_f), k = (int)(i - (int64_t)j * in_f);
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < rows; ++r) acc += __ldg(dy + (int64_t)r * out_f + j) * __ldg(x + (int64_t)r * in_f + k);     // 16 loads in flight, same order
    dW[i] += acc;
    if (db && k == 0) {
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += dy[(int64_t)r * out_f + j];
      db[j] += s;
    }
  }
}
// 16 x 16 tile of dW per block, the rows staged 32 at a time through shared memory (coalesced 64-byte row pieces): the thread-per-
// element kernel above walks `rows` dependent, strided loads per thread (96 us for the 16 -> 3840 zi_scaler at 256 rows)
__global__ void __launch_bounds__(256) k_linear_dw_tiled(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                                                         float* __restrict__ db, int rows, int in_f, int out_f) {
  __shared__ float sd[32][17], sx[32][17];
  const int jj = threadIdx.x >> 4, kk = threadIdx.x & 15;
  const int j0 = blockIdx.y * 16, k0 = blockIdx.x * 16;
  float acc = 0.f, bs = 0.f;
  for (int r0 = 0; r0 < rows; r0 += 32) {
    for (int t = threadIdx.x; t < 512; t += 256) {
      const int rr = t >> 4, cc = t & 15, r = r0 + rr;
      sd[rr][cc] = (r < rows && j0 + cc < out_f) ? __ldg(dy + (int64_t)r * out_f + j0 + cc) : 0.f;
      sx[rr][cc] = (r < rows && k0 + cc < in_f) ? __ldg(x + (int64_t)r * in_f + k0 + cc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) { acc += sd[rr][jj] * sx[rr][kk]; bs += sd[rr][jj]; }
    __syncthreads();
  }
  const int j = j0 + jj, k = k0 + kk;
  if (j < out_f && k < in_f) dW[(int64_t)j * in_f + k] += acc;
  if (db && k0 == 0 && kk == 0 && j < out_f) db[j] += bs;
}
// long reductions (out_f large, e.g. the 16 -> 3840 zi_scaler): one warp per dx element, lanes stride over j
__global__ void k_linear_dx_warp(const float* __restrict__ dy, const float* __restrict__ W, float* __restrict__ dx, int rows,
                                 int in_f, int out_f) {
  int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  int64_t total = (int64_t)rows * in_f;
  for (int64_t o = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); o < total; o += (int64_t)gridDim.x * wpb) {
    int r = (int)(o / in_f), k = (int)(o - (int64_t)r * in_f);
    float acc = 0.f;
    for (int j = lane; j < out_f; j += 32) acc += dy[(int64_t)r * out_f + j] * W[(int64_t)j * in_f + k];
    acc = warp_sum(acc);
    if (lane == 0) dx[o] = acc;
  }
}
// dW for many rows: one warp per dW element would be wasteful; split rows over a 2-D grid instead (kept simple: rows <= 512)
extern "C" int rd_linear_bwd(rd_ctx* ctx, const float* x, const float* W, const float* dy, float* dx, float* dW, float* db,
                             int rows, int in_f, int out_f, rd_stream st) {
  cudaStream_t s = (cudaStream_t)st;
  if (dx) {
    if (out_f >= 256)
      k_linear_dx_warp<<<rd_grid_1d((int64_t)rows * in_f, 8, ctx->sm_count), 256, 0, s>>>(dy, W, dx, rows, in_f, out_f);
    else
      k_linear_dx<<<rd_grid_1d((int64_t)rows * in_f, 256, ctx->sm_count), 256, 0, s>>>(dy, W, dx, rows, in_f, out_f);
    RD_CHECK_LAUNCH(ctx, "linear_dx");
  }
  if (dW) {
    if (rows >= 32) {
      dim3 grid(rd_div_up(in_f, 16), rd_div_up(out_f, 16));
      k_linear_dw_tiled<<<grid, 256, 0, s>>>(x, dy, dW, db, rows, in_f, out_f);
    } else
    k_linear_dw<<<rd_grid_1d((int64_t)out_f * in_f, 256, ctx->sm_count), 256, 0, s>>>(x, dy, dW, db, rows, in_f, out_f);
    RD_CHECK_LAUNCH(ctx, "linear_dw");
  }
  return RD_OK;
}

__global__ void k_sample_fwd(const float* mu, const float* lv, const float* eps, float* z, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) z[i] = mu[i] + eps[i] * expf(0.5f * lv[i]);
}
__global__ void k_sample_bwd(const float* dz, const float* lv, const float* eps, float* dmu, float* dlv, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    dmu[i] = dz[i];
    dlv[i] = dz[i] * eps[i] * 0.5f * expf(0.5f * lv[i]);
  }
}
extern "C" int rd_sample_fwd(rd_ctx* ctx, const float* mu, const float* lv, const float* eps, float* z, int64_t n, rd_stream st) {
  k_sample_fwd<<<rd_div_up(n, 256), 256, 0, (cudaStream_t)st>>>(mu, lv, eps, z, n);
  RD_CHECK_LAUNCH(ctx, "sample_fwd");
  return RD_OK;
}
extern "C" int rd_sample_bwd(rd_ctx* ctx, const float* dz, const float* lv, const float* eps, float* dmu, float* dlv, int64_t n,
                             rd_stream st) {
  k_sample_bwd<<<rd_div_up(n, 256), 256, 0, (cudaStream_t)st>>>(dz, lv, eps, dmu, dlv, n);
  RD_CHECK_LAUNCH(ctx, "sample_bwd");
  return RD_OK;
}

// ============================================================================ attention-gate helpers (output decoder U+SA)
// SpatialAttentionLayer (reference src/model.py:1316-1327): relu(x_post + g_post), sigmoid, alpha * x.
template <typename T>
__global__ void k_add_relu_fwd(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(a + i) + ldf<T>(b + i);
    stf<T>(y + i, v > 0.f ? v : 0.f);
  }
}
template <typename T>
__global__ void k_relu_bwd(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(dx + i, ldf<T>(y + i) > 0.f ? ldf<T>(dy + i) : 0.f);
}
template <typename T>
__global__ void k_sigmoid_fwd(const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(y + i, 1.f / (1.f + expf(-ldf<T>(x + i))));
}
template <typename T>
__global__ void k_sigmoid_bwd(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(y + i);
    stf<T>(dx + i, ldf<T>(dy + i) * v * (1.f - v));
  }
}
// y[p, c] = (off + alpha[p]) * x[p, c]      (off = 0: attention gate; off = 1: residual attention, src/model.py:1414)
template <typename T>
__global__ void k_mul_bcast_fwd(const T* __restrict__ alpha, const T* __restrict__ x, T* __restrict__ y, int64_t pixels, int C, float off) {
  int64_t total = pixels * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(y + i, (off + ldf<T>(alpha + i / C)) * ldf<T>(x + i));
}
// dx = alpha * dy ; dalpha[p] = sum_c dy[p,c] * x[p,c]   (one warp per pixel)
template <typename T>
__global__ void k_mul_bcast_bwd(const T* __restrict__ alpha, const T* __restrict__ x, const T* __restrict__ dy,
                                T* __restrict__ dx, T* __restrict__ dalpha, int64_t pixels, int C, float off) {
  int lane = threadIdx.x & 31;
  int wpb = blockDim.x >> 5;
  for (int64_t p = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); p < pixels; p += (int64_t)gridDim.x * wpb) {
    float a = off + ldf<T>(alpha + p), acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      float d = ldf<T>(dy + p * C + c);
      acc += d * ldf<T>(x + p * C + c);
      stf<T>(dx + p * C + c, a * d);
    }
    acc = warp_sum(acc);
    if (lane == 0) stf<T>(dalpha + p, acc);
  }
}
extern "C" int rd_add_relu_fwd(rd_ctx* ctx, const void* a, const void* b, void* y, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_add_relu_fwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)a, (const T*)b, (T*)y, n));
  RD_CHECK_LAUNCH(ctx, "add_relu_fwd");
  return RD_OK;
}
extern "C" int rd_relu_bwd(rd_ctx* ctx, const void* dy, const void* y, void* dx, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_relu_bwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)dy, (const T*)y, (T*)dx, n));
  RD_CHECK_LAUNCH(ctx, "relu_bwd");
  return RD_OK;
}
extern "C" int rd_sigmoid_fwd(rd_ctx* ctx, const void* x, void* y, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_sigmoid_fwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)x, (T*)y, n));
  RD_CHECK_LAUNCH(ctx, "sigmoid_fwd");
  return RD_OK;
}
extern "C" int rd_sigmoid_bwd(rd_ctx* ctx, const void* dy, const void* y, void* dx, int64_t n, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_sigmoid_bwd<T><<<rd_grid_1d(n, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)dy, (const T*)y, (T*)dx, n));
  RD_CHECK_LAUNCH(ctx, "sigmoid_bwd");
  return RD_OK;
}
extern "C" int rd_mul_bcast_fwd(rd_ctx* ctx, const void* alpha, const void* x, void* y, int64_t pixels, int C, float off, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_mul_bcast_fwd<T><<<rd_grid_1d(pixels * C, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)alpha, (const T*)x, (T*)y, pixels, C, off));
  RD_CHECK_LAUNCH(ctx, "mul_bcast_fwd");
  return RD_OK;
}
extern "C" int rd_mul_bcast_bwd(rd_ctx* ctx, const void* alpha, const void* x, const void* dy, void* dx, void* dalpha,
                                int64_t pixels, int C, float off, int dtype, rd_stream st) {
  RD_DISPATCH_DTYPE(dtype, k_mul_bcast_bwd<T><<<rd_grid_1d(pixels, 8, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)alpha, (const T*)x, (const T*)dy, (T*)dx, (T*)dalpha, pixels, C, off));
  RD_CHECK_LAUNCH(ctx, "mul_bcast_bwd");
  return RD_OK;
}

// ---- channel attention (squeeze and excitation, residual: src/model.py:1417-1433) and the symmetry gate's |g - flip_H(g)| (:1408-1409)
// of the output-decoder variants U+SA+CA / U+SSA+CA.  Small feature maps of the output U-Net; plain streaming kernels.
// y[n, p, c] = (1 + a[n, c]) * x[n, p, c]      (a fp32 [N][C])
template <typename T>
__global__ void k_chan_scale_fwd(const T* __restrict__ x, const float* __restrict__ a, T* __restrict__ y, int64_t hw, int C, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)C * hw);
    stf<T>(y + i, (1.f + a[n * C + c]) * ldf<T>(x + i));
  }
}
// grid (ceil(C / 32), N), block (32, 8): dx = (1 + a) dy, da[n, c] = sum_p dy * x.
template <typename T>
__global__ void k_chan_scale_bwd(const T* __restrict__ x, const float* __restrict__ a, const T* __restrict__ dy, T* __restrict__ dx,
                                 float* __restrict__ da, int64_t hw, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, n = blockIdx.y;
  float acc = 0.f;
  if (c < C) {
    const float s = 1.f + a[(int64_t)n * C + c];
    for (int64_t p = threadIdx.y; p < hw; p += 8) {
      const int64_t i = ((int64_t)n * hw + p) * C + c;
      const float d = ldf<T>(dy + i);
      acc += d * ldf<T>(x + i);
      stf<T>(dx + i, s * d);
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    da[(int64_t)n * C + c] = t;
  }
}
// dx[n, p, c] = v[n, c] * scale   (backward of the global average pool: scale = 1 / hw)
template <typename T>
__global__ void k_chan_bcast(const float* __restrict__ v, T* __restrict__ dx, int64_t hw, int C, int64_t total, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)C * hw);
    stf<T>(dx + i, v[n * C + c] * scale);
  }
}
// out[n, h, w, c] = |g[n, h, w, c] - g[n, H-1-h, w, c]|;  backward: dg[h] = sign(g[h] - g[H-1-h]) * (dout[h] + dout[H-1-h])
template <typename T>
__global__ void k_flip_absdiff(const T* __restrict__ g, const T* __restrict__ dout, T* __restrict__ out, int H, int64_t wc, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / wc, col = i - row * wc;
    const int64_t n = row / H;
    const int h = (int)(row - n * H);
    const int64_t j = (n * H + (H - 1 - h)) * wc + col;
    const float d = ldf<T>(g + i) - ldf<T>(g + j);
    if (dout == nullptr) stf<T>(out + i, fabsf(d));
    else stf<T>(out + i, (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * (ldf<T>(dout + i) + ldf<T>(dout + j)));
  }
}
extern "C" int rd_chan_scale_fwd(rd_ctx* ctx, const void* x, const float* a, void* y, int N, int64_t hw, int C, int dtype, rd_stream st) {
  const int64_t total = (int64_t)N * hw * C;
  RD_DISPATCH_DTYPE(dtype, (k_chan_scale_fwd<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)x, a, (T*)y, hw, C, total)));
  RD_CHECK_LAUNCH(ctx, "chan_scale_fwd");
  return RD_OK;
}
extern "C" int rd_chan_scale_bwd(rd_ctx* ctx, const void* x, const float* a, const void* dy, void* dx, float* da, int N, int64_t hw, int C,
                                 int dtype, rd_stream st) {
  dim3 grid(rd_div_up(C, 32), N), block(32, 8);
  RD_DISPATCH_DTYPE(dtype, (k_chan_scale_bwd<T><<<grid, block, 0, (cudaStream_t)st>>>((const T*)x, a, (const T*)dy, (T*)dx, da, hw, C)));
  RD_CHECK_LAUNCH(ctx, "chan_scale_bwd");
  return RD_OK;
}
extern "C" int rd_chan_bcast(rd_ctx* ctx, const float* v, void* dx, int N, int64_t hw, int C, float scale, int dtype, rd_stream st) {
  const int64_t total = (int64_t)N * hw * C;
  RD_DISPATCH_DTYPE(dtype, (k_chan_bcast<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>(v, (T*)dx, hw, C, total, scale)));
  RD_CHECK_LAUNCH(ctx, "chan_bcast");
  return RD_OK;
}
extern "C" int rd_flip_absdiff_fwd(rd_ctx* ctx, const void* g, void* out, int N, int H, int W, int C, int dtype, rd_stream st) {
  const int64_t total = (int64_t)N * H * W * C;
  RD_DISPATCH_DTYPE(dtype, (k_flip_absdiff<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)g, (const T*)nullptr, (T*)out, H, (int64_t)W * C, total)));
  RD_CHECK_LAUNCH(ctx, "flip_absdiff_fwd");
  return RD_OK;
}
extern "C" int rd_flip_absdiff_bwd(rd_ctx* ctx, const void* g, const void* dout, void* dg, int N, int H, int W, int C, int dtype, rd_stream st) {
  const int64_t total = (int64_t)N * H * W * C;
  RD_DISPATCH_DTYPE(dtype, (k_flip_absdiff<T><<<rd_grid_1d(total, 256, ctx->sm_count), 256, 0, (cudaStream_t)st>>>((const T*)g, (const T*)dout, (T*)dg, H, (int64_t)W * C, total)));
  RD_CHECK_LAUNCH(ctx, "flip_absdiff_bwd");
  return RD_OK;
}
