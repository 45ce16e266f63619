// tools/mma_rate.cu — microbenchmark: cycles per tcgen05.mma (K = 16, bf16) as a function of M (128 / 64), N, the shared-memory
// layout of A, and of WHERE A lives (shared memory or tensor memory), operands resident (contents irrelevant), one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/mma_rate.cu && /tmp/mma_rate
// Used to decide tile shapes for the small-N convolution layers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  uint64_t hi = 64u | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return lo | (hi << 32);
}
// A operand in tensor memory (the `[a_tmem]` form): only B is read from shared memory
__device__ __forceinline__ void umma_ta(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: A K-major SWIZZLE_128B; mode 1: A K-major no-swizzle (halo layout: LBO 2880, SBO 160); B always SW128.
// chains: number of independent TMEM accumulators the MMAs rotate over.
// commit_every > 0: a tcgen05.commit (arrive on a second mbarrier nobody waits on) after every `commit_every` loop iterations of 4 MMAs,
// the pattern of a pipelined kernel that releases an operand stage per tile; fence != 0 adds the tcgen05.fence::after_thread_sync a kernel
// issues after waiting for the stage.
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int m, int n, int chains, int iters, long long* out, int commit_every = 0, int fence = 0) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ __align__(8) uint64_t bar3;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem_raw)[i] = 0x3C003C00u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x < 32) {      // warp-uniform control flow, one elected lane issues (see rd_conv_halo.cu)
    // modes 3-5 (round 2, the weight-gradient kernels' operands): MN-major NO-SWIZZLE tiles as k_wgrad_halo lays them out (A = X halo tile,
    // LBO = one halo row of nb = 2 channel blocks, SBO = 160 B; B = dY tile, LBO = one tile row of n / 8 blocks, SBO = 128 B);
    // 3 = both MN-major, 4 = only A MN-major (B K-major SW128), 5 = only B MN-major (A K-major SW128)
    const bool a_mn = mode == 3 || mode == 4, b_mn = mode == 3 || mode == 5;
    const uint32_t idesc = make_idesc(m, n) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
    const uint64_t a0 = a_mn ? desc_nosw(base, 320u, 160u) : ((mode == 0 || mode == 5) ? desc_sw128(base) : desc_nosw(base, 2880u, 160u));
    const uint64_t b0 = b_mn ? desc_nosw(base + 24 * 1024, (uint32_t)(n / 8) * 128u, 128u) : desc_sw128(base + 24 * 1024);
    const uint64_t ak = a_mn ? 40 : ((mode == 0 || mode == 5) ? 2 : 360);
    const uint32_t t1off = chains == 2 ? (uint32_t)n : 0u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      if (el && mode == 2) {                 // A from tensor memory (columns 384.. : 8 columns per M x K16 bf16 block)
        umma_ta(tmem, tmem + 384u, b0, idesc, 1u);
        umma_ta(tmem + t1off, tmem + 392u, b0 + 2, idesc, 1u);
        umma_ta(tmem, tmem + 400u, b0 + 4, idesc, 1u);
        umma_ta(tmem + t1off, tmem + 408u, b0 + 6, idesc, 1u);
      } else if (el) {
        umma(tmem, a0, b0, idesc, 1u);
        umma(tmem + t1off, a0 + ak, b0 + 2, idesc, 1u);
        umma(tmem, a0 + 1, b0 + 4, idesc, 1u);
        umma(tmem + t1off, a0 + ak + 1, b0 + 6, idesc, 1u);
      }
      // fence: 0 = commit under the C++ `if (elected)`, 1 = + tcgen05.fence::after_thread_sync, 2 = commit predicated inside one asm block (no
      // divergent branch), 3 = one-shot (a single commit at iteration commit_every, none after), 4 = alternate between two mbarriers
      if (commit_every > 0 && (fence == 3 ? (i + 1) == commit_every : (i + 1) % commit_every == 0)) {
        if (fence == 2) {
          asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                       ::"r"(smem_u32(&bar2)) : "memory");
        } else {
          const uint32_t bsel = (fence == 4 && ((i / commit_every) & 1)) ? smem_u32(&bar3) : smem_u32(&bar2);
          if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bsel) : "memory");
          if (fence == 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      __syncwarp();
    }
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// Round 2: is the small-N rate set by the ISSUING THREAD's instruction stream, and do two issuing warps overlap?  Each issuer warp runs
// `iters` iterations of 4 MMAs (K-major SW128 operands, its own accumulator columns) followed by `pad` dependent integer operations (the
// barrier polls / address arithmetic a pipelined kernel executes between tiles) and a tcgen05.commit every 6 iterations.
__global__ void __launch_bounds__(128, 1) k_rate2(int n, int issuers, int pad, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem_raw)[i] = 0x3C003C00u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int w = threadIdx.x >> 5;
  if (w < issuers) {
    const uint32_t tmem = tmem_base_s + (uint32_t)w * 256u;
    const uint32_t idesc = make_idesc(128, n);
    const uint64_t a0 = desc_sw128(base + (uint32_t)w * 8192u);
    const uint64_t b0 = desc_sw128(base + 24 * 1024 + (uint32_t)w * 8192u);
    const uint32_t fin = smem_u32(&bars[w]), rel = smem_u32(&bars[2 + w]);
    uint32_t x = threadIdx.x + 1u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      if (el) {
        umma(tmem, a0, b0, idesc, 1u);
        umma(tmem, a0 + 2, b0 + 2, idesc, 1u);
        umma(tmem, a0 + 4, b0 + 4, idesc, 1u);
        umma(tmem, a0 + 6, b0 + 6, idesc, 1u);
      }
      for (int k = 0; k < pad; ++k) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(x));      // dependent chain
      if (el && (i % 6) == 5) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rel) : "memory");
      __syncwarp();
    }
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fin) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(fin) : "memory");
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 2 + w] = (t1 - t0) + (x == 12345u ? 1 : 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
  }
}

int main(int argc, char** argv) {
  const bool fast = argc > 1;          // any argument: only the MN-major modes
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  const int ns[] = {16, 32, 64, 128, 256};
  const char* names[6] = {"smem SW128", "smem no-swz", "TMEM", "MN/MN no-swz", "A MN no-swz", "B MN no-swz"};
  printf("cycles per tcgen05.mma (K=16, bf16, cta_group::1), %d MMAs issued back to back by one thread, grid 148\n", iters * 4);
  printf("ideal = M x N x 16 MACs at 4096 dense bf16 MAC/clk/SM (2.25 PFLOP/s / 148 SMs / 1.9 GHz)\n");
  for (int m : {128, 64})
    for (int mode = (fast ? 3 : 0); mode < 6; ++mode)
      for (int chains = 1; chains <= 2; ++chains)
        for (int n : ns) {
          if (chains * n > 384) continue;
          if (m == 64 && mode == 1) continue;
          if (mode >= 3 && chains == 2) continue;
          k_rate<<<148, 128, 64 * 1024>>>(mode, m, n, chains, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error (M %d mode %d N %d): %s\n", m, mode, n, cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148; ++i) s += (double)h[i];
          const double cyc = s / 148 / (iters * 4);
          printf("M %3d  A %-12s chains %d  N %3d : %6.1f cycles/MMA  (ideal %5.1f, %4.0f %% of the tensor peak)\n", m, names[mode], chains, n, cyc,
                 (double)m * n / 256.0, 100.0 * m * n / 256.0 / cyc);
        }
  // round 2: what a per-tile tcgen05.commit costs (both operands MN-major no-swizzle as in k_wgrad_halo; the K-major numbers are the same)
  const char* fn[5] = {"", " + fence::after_thread_sync", " (predicated in asm, no divergent branch)", " (ONE commit only, then none)", " (two mbarriers alternating)"};
  for (int n : {32, 64, 256})
    for (int ce : {0, 12, 6, 1})
      for (int fence = 0; fence < 5; ++fence) {
        if (ce == 0 && fence) continue;
        if (n != 32 && (fence == 1 || fence == 4)) continue;
        k_rate<<<148, 128, 64 * 1024>>>(3, 128, n, 1, iters, d, ce, fence);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("M 128 N %3d, tcgen05.commit every %2d MMAs%s : %6.1f cycles/MMA\n", n, ce * 4, fn[fence], s / 148 / (iters * 4));
      }
  // round 2: issue-stream experiment (k_rate2)
  {
    long long* d2;
    cudaMalloc(&d2, 296 * sizeof(long long));
    cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {32, 64, 128})
      for (int pad : {0, 16, 48, 96})
        for (int issuers = 1; issuers <= 2; ++issuers) {
          cudaMemset(d2, 0, 296 * sizeof(long long));
          k_rate2<<<148, 128, 64 * 1024>>>(n, issuers, pad, iters, d2);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[296];
          cudaMemcpy(h, d2, sizeof(h), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148; ++i) s += (double)(h[2 * i] > h[2 * i + 1] ? h[2 * i] : h[2 * i + 1]);
          printf("issue stream: N %3d, %2d dependent integer ops per 4 MMAs, %d issuing warp(s): %6.1f cycles per MMA (aggregate over the SM)\n", n, pad, issuers,
                 s / 148 / (iters * 4 * issuers));
        }
  }
  return 0;
}
