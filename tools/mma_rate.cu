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
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int m, int n, int chains, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((uint32_t*)smem_raw)[i] = 0x3C003C00u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
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
    const uint32_t idesc = make_idesc(m, n);
    const uint64_t a0 = mode == 0 ? desc_sw128(base) : desc_nosw(base, 2880u, 160u);
    const uint64_t b0 = desc_sw128(base + 24 * 1024);
    const uint64_t ak = mode == 0 ? 2 : 360;
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

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  const int ns[] = {16, 32, 64, 128, 256};
  const char* names[3] = {"smem SW128", "smem no-swz", "TMEM"};
  printf("cycles per tcgen05.mma (K=16, bf16, cta_group::1), %d MMAs issued back to back by one thread, grid 148\n", iters * 4);
  printf("ideal = M x N x 16 MACs at 4096 dense bf16 MAC/clk/SM (2.25 PFLOP/s / 148 SMs / 1.9 GHz)\n");
  for (int m : {128, 64})
    for (int mode = 0; mode < 3; ++mode)
      for (int chains = 1; chains <= 2; ++chains)
        for (int n : ns) {
          if (chains * n > 384) continue;
          if (m == 64 && mode == 1) continue;
          k_rate<<<148, 128, 64 * 1024>>>(mode, m, n, chains, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error (M %d mode %d N %d): %s\n", m, mode, n, cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          double s = 0;
          for (int i = 0; i < 148; ++i) s += (double)h[i];
          const double cyc = s / 148 / (iters * 4);
          printf("M %3d  A %-11s chains %d  N %3d : %6.1f cycles/MMA  (ideal %5.1f, %4.0f %% of the tensor peak)\n", m, names[mode], chains, n, cyc,
                 (double)m * n / 256.0, 100.0 * m * n / 256.0 / cyc);
        }
  return 0;
}
