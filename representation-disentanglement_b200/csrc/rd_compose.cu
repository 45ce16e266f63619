// rd_compose.cu — weight-space composition of the last two convolutions of a SPADE decoder half (reference
// src/model.py:2606-2612: SPADEBlockNew sp6 `out` (3x3, C -> C') is followed by the CondConv 1x1 `out` (C' -> in_num_ch) with
// nothing in between), so per weight group g
//     W_eff[g] = mix(W_B)[g] . mix(W_A)[g]          b_eff[g] = mix(W_B)[g] . b_A + b_B
// is ONE 3x3 convolution.  These kernels do the (tiny: 7 x 16 x 288 per group) products and their chain rule; the convolution
// itself runs on the tensor-core kernels.  fp32 throughout, deterministic (no atomics).
#include "rd_common.cuh"

namespace {

// grid (ceil(T*Cin / 128), G); thread = one (tap, input channel) of group g: all OB outputs.
template <typename T, int OBMAX>
__global__ void k_compose_fwd(const float* __restrict__ pA, const float* __restrict__ pB, const float* __restrict__ bA,
                              const float* __restrict__ bB, int Gm, int OA, int OB, int taps, int Cin, int o_pad,
                              T* __restrict__ packed, T* __restrict__ packedT, float* __restrict__ b_eff) {
  extern __shared__ float sB[];                    // pB[g]: [OB][OA]
  const int g = blockIdx.y, TI = taps * Cin;
  for (int k = threadIdx.x; k < OB * OA; k += blockDim.x) sB[k] = pB[(size_t)g * OB * OA + k];
  __syncthreads();
  const int ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < OB) {       // bias row of the group
    const int o = threadIdx.x, m = g / Gm;
    float acc = bB ? bB[m * OB + o] : 0.f;
    if (bA)
      for (int c = 0; c < OA; ++c) acc = fmaf(sB[o * OA + c], bA[m * OA + c], acc);
    b_eff[g * OB + o] = acc;
  }
  if (ti >= TI) return;
  float acc[OBMAX];
#pragma unroll
  for (int o = 0; o < OBMAX; ++o) acc[o] = 0.f;
  const float* a = pA + (size_t)g * OA * TI + ti;
  for (int c = 0; c < OA; ++c) {
    const float av = a[(size_t)c * TI];
#pragma unroll
    for (int o = 0; o < OBMAX; ++o)
      if (o < OB) acc[o] = fmaf(sB[o * OA + c], av, acc[o]);
  }
  const int t = ti / Cin, i = ti - t * Cin;
#pragma unroll
  for (int o = 0; o < OBMAX; ++o)
    if (o < OB) stf<T>(packed + ((size_t)g * OB + o) * TI + ti, acc[o]);
  T* pt = packedT + (((size_t)g * Cin + i) * taps + t) * o_pad;
#pragma unroll
  for (int o = 0; o < OBMAX; ++o)
    if (o < o_pad) stf<T>(pt + o, o < OB ? acc[o] : 0.f);
}

// grid (ceil(T*Cin / 128), G): dpA[g, c, t, i] = sum_o pB[g, o, c] dK[g, o, t, i]
template <int OBMAX>
__global__ void k_compose_bwd_a(const float* __restrict__ dK, const float* __restrict__ pB, int OA, int OB, int o_pad, int TI,
                                float* __restrict__ dpA) {
  extern __shared__ float sB[];
  const int g = blockIdx.y;
  for (int k = threadIdx.x; k < OB * OA; k += blockDim.x) sB[k] = pB[(size_t)g * OB * OA + k];
  __syncthreads();
  const int ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= TI) return;
  float d[OBMAX];
#pragma unroll
  for (int o = 0; o < OBMAX; ++o) d[o] = o < OB ? dK[((size_t)g * o_pad + o) * TI + ti] : 0.f;
  for (int c = 0; c < OA; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int o = 0; o < OBMAX; ++o)
      if (o < OB) acc = fmaf(sB[o * OA + c], d[o], acc);
    dpA[((size_t)g * OA + c) * TI + ti] = acc;
  }
}

// grid (G + modules), 256 threads.  Blocks < G: dpB[g, o, c] = sum_{t,i} dK[g, o, t, i] pA[g, c, t, i] + db[g, o] bA[m, c]
// (one warp per (o, c) pair, lanes stride the 9 * Cin products).  Blocks >= G: the bias gradients of module m, its groups summed in
// order: dbA[m, c] += sum_g sum_o pB[g, o, c] db[g, o];  dbB[m, o] += sum_g db[g, o].
__global__ void k_compose_bwd_b(const float* __restrict__ dK, const float* __restrict__ db, const float* __restrict__ pA,
                                const float* __restrict__ pB, const float* __restrict__ bA, int G, int Gm, int OA, int OB, int o_pad,
                                int TI, float* __restrict__ dpB, float* __restrict__ dbA, float* __restrict__ dbB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((int)blockIdx.x < G) {
    const int g = blockIdx.x, m = g / Gm;
    for (int pr = warp; pr < OB * OA; pr += nw) {
      const int o = pr / OA, c = pr - o * OA;
      const float* dk = dK + ((size_t)g * o_pad + o) * TI;
      const float* a = pA + ((size_t)g * OA + c) * TI;
      float acc = 0.f;
      for (int k = lane; k < TI; k += 32) acc = fmaf(dk[k], a[k], acc);
      acc = warp_sum(acc);
      if (lane == 0) dpB[((size_t)g * OB + o) * OA + c] = acc + (bA ? db[g * o_pad + o] * bA[m * OA + c] : 0.f);
    }
    return;
  }
  const int m = blockIdx.x - G;
  for (int k = threadIdx.x; k < OA + OB; k += blockDim.x) {
    float acc = 0.f;
    if (k < OA) {
      if (!dbA) continue;
      for (int g = m * Gm; g < (m + 1) * Gm; ++g)
        for (int o = 0; o < OB; ++o) acc = fmaf(pB[((size_t)g * OB + o) * OA + k], db[g * o_pad + o], acc);
      dbA[m * OA + k] += acc;
    } else {
      if (!dbB) continue;
      const int o = k - OA;
      for (int g = m * Gm; g < (m + 1) * Gm; ++g) acc += db[g * o_pad + o];
      dbB[m * OB + o] += acc;
    }
  }
}

}  // namespace

extern "C" int rd_compose_tail_fwd(rd_ctx* ctx, const float* pA, const float* pB, const float* bA, const float* bB, int G, int modules,
                                   int OA, int OB, int taps, int Cin, int o_pad, int dtype, void* packed, void* packedT, float* b_eff,
                                   rd_stream st) {
  if (OB > 16 || o_pad > 16 || o_pad < OB || G % modules) RD_FAIL(ctx, RD_ERR_ARG, "compose_tail_fwd: OB %d o_pad %d G %d modules %d", OB, o_pad, G, modules);
  const int TI = taps * Cin;
  dim3 grid(rd_div_up(TI, 128), G);
  RD_DISPATCH_DTYPE(dtype, (k_compose_fwd<T, 16><<<grid, 128, (size_t)OB * OA * sizeof(float), (cudaStream_t)st>>>(
                               pA, pB, bA, bB, G / modules, OA, OB, taps, Cin, o_pad, (T*)packed, (T*)packedT, b_eff)));
  RD_CHECK_LAUNCH(ctx, "compose_tail_fwd");
  return RD_OK;
}

extern "C" int rd_compose_tail_bwd(rd_ctx* ctx, const float* dK, const float* db, const float* pA, const float* pB, const float* bA,
                                   int G, int modules, int OA, int OB, int taps, int Cin, int o_pad, float* dpA, float* dpB,
                                   float* dbA, float* dbB, rd_stream st) {
  if (OB > 16 || o_pad < OB || G % modules) RD_FAIL(ctx, RD_ERR_ARG, "compose_tail_bwd: OB %d o_pad %d G %d modules %d", OB, o_pad, G, modules);
  const int TI = taps * Cin;
  dim3 grid(rd_div_up(TI, 128), G);
  k_compose_bwd_a<16><<<grid, 128, (size_t)OB * OA * sizeof(float), (cudaStream_t)st>>>(dK, pB, OA, OB, o_pad, TI, dpA);
  RD_CHECK_LAUNCH(ctx, "compose_tail_bwd_a");
  k_compose_bwd_b<<<G + modules, 256, 0, (cudaStream_t)st>>>(dK, db, pA, pB, bA, G, G / modules, OA, OB, o_pad, TI, dpB, dbA, dbB);
  RD_CHECK_LAUNCH(ctx, "compose_tail_bwd_b");
  return RD_OK;
}
