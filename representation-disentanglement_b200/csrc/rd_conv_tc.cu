// rd_conv_tc.cu — implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
//   forward :  Y[p, co]  = sum_{tap, ci} X[src(p, tap), ci] * K[g][co][tap][ci]        (+bias, LeakyReLU)
//   dgrad   :  dX[q, ci] = sum_{tap, co} dY[srcT(q, tap), co] * Kt[g][ci][tap][co]     (transposed gather)
//
// GEMM view per CTA: D[128 pixels x n_tile channels] (fp32, in TMEM) += A[128 x 64] * B[n_tile x 64]^T per
// K-block of 64 (tap, channel) elements, bf16 operands.  A (activations, im2col on the fly) and B (packed
// weights) are staged in shared memory in the canonical K-major SWIZZLE_128B layout the UMMA descriptors
// expect; one elected thread issues tcgen05.mma (M=128, N=n_tile, K=16 x4 per stage) and releases stages
// with tcgen05.commit; four epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 bit),
// fuse bias + activation, convert to bf16 and write NHWC rows with 16-byte stores.
//
// Warp roles (160 threads): warps 0-3 = gather producers (cp.async 16 B, zero-fill for padding / tails)
// then epilogue; warp 4 = barrier init, TMEM alloc/dealloc, MMA issue.  S-stage mbarrier ring
// (full: 128 producer arrivals after cp.async.wait_group + fence.proxy.async; empty: tcgen05.commit).
// Two CTAs fit per SM (<= 100 KB smem, <= 256 TMEM columns each) so one CTA's epilogue overlaps the
// other's main loop.
#include <stdlib.h>
#include "rd_common.cuh"
#include "rd_tc_common.cuh"

namespace {

constexpr int kBlockM = 128;      // pixels per CTA tile (UMMA M)
constexpr int kBlockK = 64;       // bf16 elements per K-block = one 128-byte swizzle row
constexpr int kMaxStages = 4;
constexpr int kThreads = 160;

struct TcParams {
  const bf16* x; const bf16* w; const float* bias; bf16* y;
  int H, W, Cin;          // source tensor of the gather (x for fwd, dy for dgrad)
  int OH, OW, Cout;       // destination tensor
  int KH, KW, stride, pad, mode;
  int ipg;                // images per group
  int64_t ppg;            // destination pixels per group
  int tiles_pg;           // M tiles per group
  int k_total, k_blocks;
  int n_tile;             // UMMA N (multiple of 16, <= 128)
  int stages;
  int tmem_cols;
  int act; float slope;
  int* err_flag;
  int bias_gpr;           // weight groups per bias row (0: one bias row for all groups)
  // stride-2 dgrad by output-pixel parity class: a pixel (oy, ox) only receives the taps kh = (oy + pad) mod 2 (+2), kw
  // likewise, so a tile made of ONE class runs a K loop over its 4 (k = 4) or 1 / 2 / 4 (k = 3) valid taps instead of all
  // k*k with 3/4 of the gathered rows zero-filled.  tiles_pg = 4 * tiles_pc.
  int parity, tiles_pc;
  int fast;               // 32-bit fast gather (see the producer)
  int64_t ppc;            // destination pixels per class and group = ipg * OH/2 * OW/2
};

__device__ __forceinline__ bool tap_src(const TcParams& P, int oy, int ox, int kh, int kw, int& iy, int& ix) {
  if (P.mode == 0) {
    iy = oy * P.stride - P.pad + kh;
    ix = ox * P.stride - P.pad + kw;
  } else {
    int ty = oy + P.pad - kh, tx = ox + P.pad - kw;
    if (ty < 0 || tx < 0) return false;
    if (P.stride > 1) {
      if ((ty % P.stride) | (tx % P.stride)) return false;
      ty /= P.stride; tx /= P.stride;
    }
    iy = ty; ix = tx;
  }
  return iy >= 0 && iy < P.H && ix >= 0 && ix < P.W;
}

__global__ void __launch_bounds__(kThreads) k_conv_tc(const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = kBlockM * 128u;
  const uint32_t b_bytes = (uint32_t)P.n_tile * 128u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int S = P.stages;

  const int grp = blockIdx.x / P.tiles_pg;
  int tile = blockIdx.x - grp * P.tiles_pg;
  const int n0 = blockIdx.y * P.n_tile;
  // parity-class decomposition (stride-2 dgrad): class-uniform tap subset
  int py = 0, px = 0, k0y = 0, k0x = 0, nkx = P.KW;
  int k_total = P.k_total, k_blocks = P.k_blocks, vtaps = P.KH * P.KW;
  if (P.parity) {
    const int cls = tile / P.tiles_pc;
    tile -= cls * P.tiles_pc;
    py = cls >> 1; px = cls & 1;
    k0y = (py + P.pad) & 1; k0x = (px + P.pad) & 1;
    const int nky = (P.KH - k0y + 1) >> 1;
    nkx = (P.KW - k0x + 1) >> 1;
    vtaps = nky * nkx;
    k_total = vtaps * P.Cin;
    k_blocks = (k_total + kBlockK - 1) / kBlockK;
  }

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 128);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      mbar_init(smem_u32(&accum_bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int c = tid & 7;             // 16-byte chunk (8 bf16) within the 128-byte K row
    const int rsub = tid >> 3;         // 0..15
    const uint32_t row_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)(rsub & 7) * 128u + (uint32_t)((c ^ (rsub & 7)) << 4);
    const bf16* wg = P.w + ((int64_t)grp * P.Cout + n0) * P.k_total;
    const int nb = P.n_tile >> 4;
    const int taps = vtaps;
    const bool taps_inner = (P.Cin % kBlockK) == 0;
    if (P.fast) {
      // Fast gather (tensors below 2^31 elements; forward, stride-1 dgrad, parity-class stride-2 dgrad): the source coordinate of row j
      // and tap (dy, dx) is (iy0[j] + dy, ix0[j] + dx) with a signed tap displacement that is the same for all rows of the thread, so a
      // K-block costs two adds, two compares, one select and one 32->64-bit address add per row (the first version re-derived
      // everything per row and K-block: ~300 instructions per K-block and thread, 25 us per 128-pixel tile of the stride-2 encoders).
      int iy0[8], ix0[8], rbase[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t lp = (uint32_t)tile * kBlockM + rsub + 16 * j;
        uint32_t oxj, oyj, img;
        bool rv;
        if (P.parity) {
          rv = lp < (uint32_t)P.ppc;
          const uint32_t w2 = P.OW >> 1, h2 = P.OH >> 1;
          const uint32_t t = lp / w2;
          oxj = 2 * (lp - t * w2) + px;
          const uint32_t im = t / h2;
          oyj = 2 * (t - im * h2) + py;
          img = (uint32_t)grp * P.ipg + im;
        } else {
          rv = lp < (uint32_t)P.ppg;
          const uint32_t pp = (uint32_t)grp * (uint32_t)P.ppg + lp;
          const uint32_t t = pp / (uint32_t)P.OW;
          oxj = pp - t * P.OW;
          img = t / (uint32_t)P.OH;
          oyj = t - img * P.OH;
        }
        int y, x;
        if (P.mode == 0) { y = (int)oyj * P.stride - P.pad; x = (int)oxj * P.stride - P.pad; }
        else if (P.parity) { y = ((int)oyj + P.pad - k0y) >> 1; x = ((int)oxj + P.pad - k0x) >> 1; }
        else { y = (int)oyj + P.pad; x = (int)oxj + P.pad; }
        iy0[j] = rv ? y : -(1 << 28);                  // never inside the source image
        ix0[j] = x;
        rbase[j] = rv ? (int)(((img * (uint32_t)P.H + y) * (uint32_t)P.W + x) * (uint32_t)P.Cin) : 0;   // mod 2^32: exact once a valid tap is added
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t a_s = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t b_s = a_s + a_bytes;
        int kk, tap = 0, ci = 0, kh = 0, kw = 0;
        if (taps_inner) {
          const int chunk = kb / taps;
          tap = kb - chunk * taps;
          ci = chunk * kBlockK + c * 8;
          kk = tap * P.Cin + ci;
        } else {
          kk = kb * kBlockK + c * 8;
        }
        const bool kvalid = kk < k_total;
        int dy = 0, dx = 0;
        if (kvalid) {
          if (!taps_inner) { tap = kk / P.Cin; ci = kk - tap * P.Cin; }
          if (P.parity) { const int vy = tap / nkx, vx = tap - vy * nkx; kh = k0y + 2 * vy; kw = k0x + 2 * vx; dy = -vy; dx = -vx; }
          else { kh = tap / P.KW; kw = tap - kh * P.KW; dy = P.mode == 0 ? kh : -kh; dx = P.mode == 0 ? kw : -kw; }
          kk = (kh * P.KW + kw) * P.Cin + ci;          // column of the packed weight row
        }
        const int doff = (dy * P.W + dx) * P.Cin + ci;
        const int ylim = kvalid ? P.H : 0;             // an invalid K column (tail of the last block) zero-fills every row
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool v = (uint32_t)(iy0[j] + dy) < (uint32_t)ylim && (uint32_t)(ix0[j] + dx) < (uint32_t)P.W;
          const uint32_t off = v ? (uint32_t)(rbase[j] + doff) : 0u;
          cp_async16_ca(a_s + row_off + 2048u * j, P.x + off, v ? 16u : 0u);
        }
        const bf16* wrow = wg + (int64_t)rsub * P.k_total + kk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nb) {
            const bool v = kvalid && (n0 + rsub + 16 * j) < P.Cout;   // weight rows beyond Cout are zero-filled
            cp_async16(b_s + row_off + 2048u * j, v ? (const void*)(wrow + (int64_t)(16 * j) * P.k_total) : (const void*)P.w, v ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive(smem_u32(&full_bar[s]));
      }
    } else {
    int oy[8], ox[8];
    int64_t img_off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int64_t lp = (int64_t)tile * kBlockM + rsub + 16 * j;
      if (P.parity) {
        if (lp < P.ppc) {
          const int w2 = P.OW >> 1, h2 = P.OH >> 1;
          ox[j] = 2 * (int)(lp % w2) + px;
          int64_t t = lp / w2;
          oy[j] = 2 * (int)(t % h2) + py;
          img_off[j] = ((int64_t)grp * P.ipg + t / h2) * (int64_t)P.H * P.W;
        } else {
          ox[j] = 0; oy[j] = -(1 << 28); img_off[j] = 0;
        }
      } else if (lp < P.ppg) {
        int64_t p = (int64_t)grp * P.ppg + lp;
        ox[j] = (int)(p % P.OW);
        int64_t t = p / P.OW;
        oy[j] = (int)(t % P.OH);
        img_off[j] = (t / P.OH) * (int64_t)P.H * P.W;
      } else {
        ox[j] = 0; oy[j] = -(1 << 28); img_off[j] = 0;   // never inside the source image
      }
    }
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
      const uint32_t a_s = smem_base + (uint32_t)s * stage_bytes;
      const uint32_t b_s = a_s + a_bytes;
      // K-block order: when Cin is a multiple of 64 the taps are the INNER loop (block = (channel chunk, tap)) so
      // that consecutive stages re-read the same input neighbourhood (L1 hits); otherwise k = tap*Cin + ci linear.
      int kk, tap = 0, ci = 0, kh = 0, kw = 0;
      // (virtual) tap index -> (kh, kw): all taps, or the class's subset {k0 + 2 m}
      if (taps_inner) {
        const int chunk = kb / taps;
        tap = kb - chunk * taps;
        ci = chunk * kBlockK + c * 8;
        kk = tap * P.Cin + ci;
      } else {
        kk = kb * kBlockK + c * 8;
      }
      const bool kvalid = kk < k_total;
      if (kvalid) {
        if (!taps_inner) { tap = kk / P.Cin; ci = kk - tap * P.Cin; }
        if (P.parity) { const int vy = tap / nkx; kh = k0y + 2 * vy; kw = k0x + 2 * (tap - vy * nkx); }
        else { kh = tap / P.KW; kw = tap - kh * P.KW; }
        kk = (kh * P.KW + kw) * P.Cin + ci;          // column of the packed weight row
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int iy, ix;
        bool v = kvalid && tap_src(P, oy[j], ox[j], kh, kw, iy, ix);
        const bf16* src = v ? P.x + (img_off[j] + (int64_t)iy * P.W + ix) * P.Cin + ci : P.x;
        cp_async16_ca(a_s + row_off + 2048u * j, src, v ? 16u : 0u);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          int co = rsub + 16 * j;
          bool v = kvalid && (n0 + co) < P.Cout;   // weight rows beyond Cout are zero-filled
          const bf16* src = v ? wg + (int64_t)co * P.k_total + kk : P.w;
          cp_async16(b_s + row_off + 2048u * j, src, v ? 16u : 0u);
        }
      }
      // arrive-on-completion of this thread's copies (one of the 128 expected arrivals): the gather of the next stages is issued
      // without waiting.  (cp.async.wait_group + fence.proxy.async here compiled to MEMBAR.ALL.CTA, which waits for EVERY copy in
      // flight: the gather latency of each K-block was fully exposed, 2-3 us per block.)
      cp_async_mbar_arrive(smem_u32(&full_bar[s]));
    }
    }

    // ------------------------------------------------------------------ epilogue
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const int64_t lp = (int64_t)tile * kBlockM + tid;
    bool pvalid = lp < P.ppg;
    int64_t dpix = (int64_t)grp * P.ppg + (pvalid ? lp : 0);
    if (P.parity) {
      pvalid = lp < P.ppc;
      const int w2 = P.OW >> 1, h2 = P.OH >> 1;
      const int64_t l2 = pvalid ? lp : 0;
      const int64_t t = l2 / w2;
      dpix = (((int64_t)grp * P.ipg + t / h2) * P.OH + (2 * (int)(t % h2) + py)) * P.OW + (2 * (int)(l2 % w2) + px);
    }
    bf16* yrow = P.y + dpix * P.Cout + n0;
    const float* biasg = P.bias ? P.bias + (size_t)(P.bias_gpr ? grp / P.bias_gpr : 0) * P.Cout : nullptr;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < P.n_tile; cb += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)cb, r);
      if (pvalid && (P.Cout & 7)) {
        // channel count not a multiple of 8 (4-channel anatomy logits, 7-channel images): scalar stores
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          int co = n0 + cb + q;
          if (co < P.Cout) {
            float v0 = __uint_as_float(r[q]);
            if (biasg) v0 += biasg[co];
            if (P.act == RD_ACT_LRELU) v0 = v0 > 0.f ? v0 : v0 * P.slope;
            yrow[cb + q] = __float2bfloat16_rn(v0);
          }
        }
      } else if (pvalid) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int co = n0 + cb + h * 8;
          if (co < P.Cout) {   // Cout % 8 == 0, so a group of 8 is entirely valid or entirely out
            uint32_t packed[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v0 = __uint_as_float(r[h * 8 + 2 * q]), v1 = __uint_as_float(r[h * 8 + 2 * q + 1]);
              if (biasg) { v0 += biasg[co + 2 * q]; v1 += biasg[co + 2 * q + 1]; }
              if (P.act == RD_ACT_LRELU) { v0 = v0 > 0.f ? v0 : v0 * P.slope; v1 = v1 > 0.f ? v1 : v1 * P.slope; }
              __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
              packed[q] = *reinterpret_cast<uint32_t*>(&b2);
            }
            *reinterpret_cast<uint4*>(yrow + cb + h * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kBlockM, P.n_tile);
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t a_s = smem_base + (uint32_t)s * stage_bytes;
        const uint64_t adesc = make_desc_k_sw128(a_s);
        const uint64_t bdesc = make_desc_k_sw128(a_s + a_bytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)   // +32 bytes (= 2 x 16 B units) per UMMA_K step inside the swizzle row
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(&accum_bar));
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

}  // namespace

int rd_conv_tc_supported(const rd_conv_desc* d, int mode) {
  if (d->dtype != RD_BF16) return 0;
  int cin = mode == 0 ? d->cin : d->cout, cout = mode == 0 ? d->cout : d->cin;
  if (cin % 8 || cin < 8) return 0;      // gathered tensor: 16-byte channel vectors
  if (cout < 1) return 0;                // written tensor: any channel count (scalar stores when not % 8)
  return 1;
}

int rd_conv_tc_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                      cudaStream_t st) {
  static const bool no_tma = getenv("RD_B200_NO_TMA") != nullptr;
  const char* opn = mode == 0 ? "fwd" : "dgrad";
  if (d->algo == RD_ALGO_HALO) { rd_trace_conv(opn, "halo", d); return rd_conv_halo_launch(ctx, d, mode, x, w, bias, y, st); }
  if (!no_tma && rd_conv_halo_supported(d, mode, ctx->sm_count, 0)) { rd_trace_conv(opn, "halo", d); return rd_conv_halo_launch(ctx, d, mode, x, w, bias, y, st); }
  if (!no_tma && rd_conv_tma_supported(d, mode)) { rd_trace_conv(opn, "tma", d); return rd_conv_tma_launch(ctx, d, mode, x, w, bias, y, st); }
  rd_trace_conv(opn, "tc_gather", d);
  TcParams P;
  P.x = (const bf16*)x; P.w = (const bf16*)w; P.bias = bias; P.y = (bf16*)y;
  if (mode == 0) { P.H = d->h; P.W = d->w; P.Cin = d->cin; P.OH = d->oh; P.OW = d->ow; P.Cout = d->cout; }
  else { P.H = d->oh; P.W = d->ow; P.Cin = d->cout; P.OH = d->h; P.OW = d->w; P.Cout = d->cin; }
  P.KH = d->kh; P.KW = d->kw; P.stride = d->stride; P.pad = d->pad; P.mode = mode;
  P.ipg = d->n / d->groups;
  P.ppg = (int64_t)P.ipg * P.OH * P.OW;
  P.tiles_pg = rd_div_up(P.ppg, kBlockM);
  P.k_total = d->kh * d->kw * P.Cin;
  P.k_blocks = rd_div_up(P.k_total, kBlockK);
  int n_tile = ((P.Cout + 15) / 16) * 16;
  if (n_tile > 128) n_tile = 128;
  P.n_tile = n_tile;
  P.stages = n_tile <= 64 ? 4 : 3;
  int cols = 32;
  while (cols < n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.act = mode == 0 ? d->act : RD_ACT_NONE;
  P.slope = d->act_slope;
  P.err_flag = nullptr;
  P.bias_gpr = (mode == 0 && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;
  P.parity = 0; P.tiles_pc = 0; P.ppc = 0;
  static const bool no_parity = getenv("RD_B200_NO_PARITY") != nullptr;
  if (mode == 1 && d->stride == 2 && (P.OH % 2 == 0) && (P.OW % 2 == 0) && d->kh >= 2 && d->kw >= 2 && !no_parity) {
    P.parity = 1;
    P.ppc = (int64_t)P.ipg * (P.OH / 2) * (P.OW / 2);
    P.tiles_pc = rd_div_up(P.ppc, kBlockM);
    P.tiles_pg = 4 * P.tiles_pc;
  }
  {
    static const bool no_fast = getenv("RD_B200_TC_NO_FAST") != nullptr;
    const int64_t src_elems = (int64_t)d->n * P.H * P.W * P.Cin, dst_pix = (int64_t)d->n * P.OH * P.OW;
    P.fast = !no_fast && src_elems < (1ll << 31) && dst_pix + kBlockM < (1ll << 31) && (mode == 0 || d->stride == 1 || P.parity);
  }
  size_t smem = (size_t)P.stages * (kBlockM * 128 + n_tile * 128) + 1024;
  if (!ctx->tc_attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ctx->tc_attr_set = true;
  }
  dim3 grid(P.tiles_pg * d->groups, rd_div_up(P.Cout, n_tile));
  k_conv_tc<<<grid, kThreads, smem, st>>>(P);
  RD_CHECK_LAUNCH(ctx, mode == 0 ? "conv_tc_fwd" : "conv_tc_dgrad");
  return RD_OK;
}


// =====================================================================================================
// wgrad:  dK[g][co][tap][ci] = sum over the group's output pixels p of dY[p, co] * X[src(p, tap), ci]
//
// GEMM view: the reduction (K) dimension is the PIXEL index, so both operands are "MN-major" in shared memory:
// one 128-byte row per pixel holding 64 consecutive channels (bf16), 8-row groups 1024 B apart (SBO), 64-channel
// blocks `blk_stride` apart (LBO) — the canonical SWIZZLE_128B MN-major layout of cute::UMMA::make_umma_desc.
//   normal     (Cout >= 128): D[co (M=128)] [n' (N<=256)]   A = dY block(s), B = im2col(X) blocks
//   transposed (Cout <  128): D[n' (M=128)] [co (N=Cout)]   A = im2col(X) blocks, B = dY block
// with n' = tap*Cin + ci.  One CTA = (group, pixel chunk, M tile, N range); K-blocks of 64 pixels, 3-stage
// cp.async ring, one tcgen05.mma per 16 pixels, fp32 accumulator in TMEM, epilogue = red.global.add.f32 into dK
// (split-K over pixel chunks).
namespace {

constexpr int kWgPixBlock = 64;      // pixels per stage (K-block)
constexpr int kWgStages = 3;
constexpr int kWgBlockBytes = kWgPixBlock * 128;   // one 64-channel block: 64 rows x 128 B = 8 KB

__device__ __align__(16) const unsigned short g_ones_chunk[8] = {0x3F80, 0, 0, 0, 0, 0, 0, 0};   // bf16 {1,0,0,0,0,0,0,0}

struct WgParams {
  const bf16* x; const bf16* dy; float* dK; float* dbias;
  int dbias_gpr;          // weight groups per bias-gradient row (0: one row)
  int n_ext;              // n_total (+8 when the bias gradient rides along as an extra "ones" im2col column)
  int H, W, Cin, OH, OW, Cout, KH, KW, stride, pad;
  int64_t ppg;            // output pixels per group
  int chunk_pixels;       // pixels per CTA (multiple of 64)
  int chunks_pg;          // pixel chunks per group
  int n_total;            // taps * Cin
  int transposed;
  int m_blocks;           // 64-channel blocks on the M side (always 2 -> M = 128)
  int n_blocks;           // 64-channel blocks on the N side
  int n_width;            // UMMA N (multiple of 16, <= 256)
  int n_splits;           // CTAs along the im2col (n') axis
  int np_per_cta;         // n' covered per CTA (multiple of 64) — on the N side (normal) or M side (transposed: 128)
  int tmem_cols;
  int fast;               // 32-bit decode / offsets
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16);
  uint64_t hi = 64u /* SBO = 1024 B */ | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}
__device__ __forceinline__ uint32_t make_idesc_mn(int m, int n) {   // both operands MN-major: bits 15, 16
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads) k_wgrad_tc(const WgParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)P.m_blocks * kWgBlockBytes;
  const uint32_t b_bytes = (uint32_t)P.n_blocks * kWgBlockBytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  const int grp = blockIdx.z;
  const int chunk = blockIdx.x / P.n_splits;
  const int nsplit = blockIdx.x - chunk * P.n_splits;
  const int tile_y = blockIdx.y;
  // element offsets of this CTA on the two logical axes
  const int co0 = P.transposed ? 0 : tile_y * 128;
  const int np0 = P.transposed ? (tile_y * P.n_splits + nsplit) * 128 : nsplit * P.np_per_cta;
  const int64_t pix0 = (int64_t)chunk * P.chunk_pixels;
  int64_t pix1 = pix0 + P.chunk_pixels;
  if (pix1 > P.ppg) pix1 = P.ppg;
  const int k_blocks = (int)((pix1 - pix0 + kWgPixBlock - 1) / kWgPixBlock);

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kWgStages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 128);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      mbar_init(smem_u32(&accum_bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    const int c = tid & 7;
    const int rsub = tid >> 3;       // 0..15 ; rows rsub + 16*j, j = 0..3
    const uint32_t row_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)(rsub & 7) * 128u + (uint32_t)((c ^ (rsub & 7)) << 4);
    const int dy_blocks = P.transposed ? P.n_blocks : P.m_blocks;
    const int im_blocks = P.transposed ? P.m_blocks : P.n_blocks;
    const uint32_t dy_off = P.transposed ? a_bytes : 0u;        // dY is the B operand when transposed
    const uint32_t im_off = P.transposed ? 0u : a_bytes;
    // fast path state: the (tap, channel) of each im2col block is K-block invariant; the four pixel rows of the thread advance by
    // 64 pixels per K-block (quotient / remainder steps instead of two divisions per row and block)
    int tkh[4], tkw[4];
    uint32_t tdoff[4], tflag[4];
    uint32_t rox[4], roy[4], rim[4], rlp[4];
    const uint32_t step_q = (uint32_t)kWgPixBlock / (uint32_t)P.OW, step_r = (uint32_t)kWgPixBlock % (uint32_t)P.OW;
    if (P.fast) {
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const int np = np0 + blk * 64 + c * 8;
        const bool in_cta = P.transposed || (blk * 64 + c * 8) < P.np_per_cta;
        const bool nv = np < P.n_total && in_cta;
        const bool ones = (np == P.n_total) && (P.n_ext > P.n_total) && in_cta;   // bias-gradient column: dY^T * 1
        int ci = 0, kh = 0, kw = 0;
        if (nv) { const int tap = np / P.Cin; ci = np - tap * P.Cin; kh = tap / P.KW; kw = tap - kh * P.KW; }
        tkh[blk] = kh; tkw[blk] = kw;
        tdoff[blk] = (uint32_t)((kh * P.W + kw) * P.Cin + ci);
        tflag[blk] = (nv ? 1u : 0u) | (ones ? 2u : 0u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t lp = (uint32_t)pix0 + rsub + 16 * j;
        const uint32_t pp = (uint32_t)grp * (uint32_t)P.ppg + lp;
        const uint32_t t = pp / (uint32_t)P.OW;
        rlp[j] = lp; rox[j] = pp - t * P.OW;
        rim[j] = t / (uint32_t)P.OH;
        roy[j] = t - rim[j] * P.OH;
      }
    }
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % kWgStages;
      const uint32_t ph = (uint32_t)(kb / kWgStages) & 1u;
      mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
      const uint32_t st_base = smem_base + (uint32_t)s * stage_bytes;
      if (P.fast) {
        // 32-bit offsets (tensors below 2^31 elements): two adds, two compares, one select and one address add per (row, block)
        int iy0[4], ix0[4];
        uint32_t xbase[4], dybase[4];
        bool rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rv[j] = rlp[j] < (uint32_t)pix1;
          iy0[j] = rv[j] ? (int)roy[j] * P.stride - P.pad : -(1 << 28);
          ix0[j] = (int)rox[j] * P.stride - P.pad;
          xbase[j] = ((rim[j] * (uint32_t)P.H + (uint32_t)iy0[j]) * (uint32_t)P.W + (uint32_t)ix0[j]) * (uint32_t)P.Cin;   // mod 2^32
          dybase[j] = ((rim[j] * (uint32_t)P.OH + roy[j]) * (uint32_t)P.OW + rox[j]) * (uint32_t)P.Cout;
          // advance the row by one K-block (64 pixels)
          rlp[j] += kWgPixBlock;
          rox[j] += step_r; roy[j] += step_q;
          if (rox[j] >= (uint32_t)P.OW) { rox[j] -= P.OW; ++roy[j]; }
          while (roy[j] >= (uint32_t)P.OH) { roy[j] -= P.OH; ++rim[j]; }
        }
        for (int blk = 0; blk < dy_blocks; ++blk) {
          const int co = co0 + blk * 64 + c * 8;
          const bool cv = co < P.Cout;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool v = cv && rv[j];
            cp_async16_ca(st_base + dy_off + (uint32_t)blk * kWgBlockBytes + row_off + 2048u * j, P.dy + (v ? dybase[j] + (uint32_t)co : 0u), v ? 16u : 0u);
          }
        }
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          if (blk < im_blocks) {
            const bool ones = (tflag[blk] & 2u) != 0;
            const uint32_t ylim = (tflag[blk] & 1u) ? (uint32_t)P.H : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              bool v = (uint32_t)(iy0[j] + tkh[blk]) < ylim && (uint32_t)(ix0[j] + tkw[blk]) < (uint32_t)P.W;
              const bf16* src = P.x + (v ? xbase[j] + tdoff[blk] : 0u);
              if (ones && rv[j]) { src = reinterpret_cast<const bf16*>(g_ones_chunk); v = true; }
              cp_async16_ca(st_base + im_off + (uint32_t)blk * kWgBlockBytes + row_off + 2048u * j, src, v ? 16u : 0u);
            }
          }
        }
      } else {
      // pixel decode for this thread's 4 rows
      int oy[4], ox[4];
      int64_t img[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t lp = pix0 + (int64_t)kb * kWgPixBlock + rsub + 16 * j;
        if (lp < pix1) {
          int64_t p = (int64_t)grp * P.ppg + lp;
          pl[j] = p;
          ox[j] = (int)(p % P.OW);
          int64_t t = p / P.OW;
          oy[j] = (int)(t % P.OH);
          img[j] = (t / P.OH) * (int64_t)P.H * P.W;
        } else {
          pl[j] = -1; ox[j] = 0; oy[j] = 0; img[j] = 0;
        }
      }
      // dY blocks: channels co0 + blk*64 + c*8
      for (int blk = 0; blk < dy_blocks; ++blk) {
        const int co = co0 + blk * 64 + c * 8;
        const bool cv = co < P.Cout;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool v = cv && pl[j] >= 0;
          const bf16* src = v ? P.dy + pl[j] * P.Cout + co : P.dy;
          cp_async16_ca(st_base + dy_off + (uint32_t)blk * kWgBlockBytes + row_off + 2048u * j, src, v ? 16u : 0u);
        }
      }
      // im2col blocks: n' = np0 + blk*64 + c*8 -> (tap, ci)
      for (int blk = 0; blk < im_blocks; ++blk) {
        const int np = np0 + blk * 64 + c * 8;
        const bool in_cta = P.transposed || (blk * 64 + c * 8) < P.np_per_cta;
        const bool nv = np < P.n_total && in_cta;
        const bool ones = (np == P.n_total) && (P.n_ext > P.n_total) && in_cta;   // bias-gradient column: dY^T * 1
        int tap = 0, ci = 0, kh = 0, kw = 0;
        if (nv) { tap = np / P.Cin; ci = np - tap * P.Cin; kh = tap / P.KW; kw = tap - kh * P.KW; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int iy = oy[j] * P.stride - P.pad + kh, ix = ox[j] * P.stride - P.pad + kw;
          bool v = nv && pl[j] >= 0 && iy >= 0 && iy < P.H && ix >= 0 && ix < P.W;
          const bf16* src = v ? P.x + (img[j] + (int64_t)iy * P.W + ix) * P.Cin + ci : P.x;
          if (ones && pl[j] >= 0) { src = reinterpret_cast<const bf16*>(g_ones_chunk); v = true; }
          cp_async16_ca(st_base + im_off + (uint32_t)blk * kWgBlockBytes + row_off + 2048u * j, src, v ? 16u : 0u);
        }
      }
      }
      cp_async_mbar_arrive(smem_u32(&full_bar[s]));      // arrive-on-completion, see k_conv_tc
    }

    // ------------------------------------------------------------------ epilogue: TMEM -> red.global.add.f32
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int taps = P.KH * P.KW;
    float* dKg = P.dK + (int64_t)grp * P.Cout * taps * P.Cin;
    for (int cb = 0; cb < P.n_width; cb += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)cb, r);
      if (!P.transposed) {
        const int co = co0 + tid;                     // M row = output channel
        if (co < P.Cout) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int np = np0 + cb + q;              // column = n' = tap*Cin + ci  (dK row-major over (tap, ci))
            if ((cb + q) < P.np_per_cta) {
              if (np < P.n_total) atomicAdd(dKg + (int64_t)co * P.n_total + np, __uint_as_float(r[q]));
              else if (np == P.n_total && P.n_ext > P.n_total) atomicAdd(P.dbias + (size_t)(P.dbias_gpr ? grp / P.dbias_gpr : 0) * P.Cout + co, __uint_as_float(r[q]));
            }
          }
        }
      } else {
        const int np = np0 + tid;                     // M row = n'
        if (np < P.n_total) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int co = cb + q;                    // column = output channel
            if (co < P.Cout) atomicAdd(dKg + (int64_t)co * P.n_total + np, __uint_as_float(r[q]));
          }
        } else if (np == P.n_total && P.n_ext > P.n_total) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int co = cb + q;
            if (co < P.Cout) atomicAdd(P.dbias + (size_t)(P.dbias_gpr ? grp / P.dbias_gpr : 0) * P.Cout + co, __uint_as_float(r[q]));
          }
        }
      }
    }
    tc_fence_before();
  } else {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_mn(128, P.n_width);
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kWgStages;
        const uint32_t ph = (uint32_t)(kb / kWgStages) & 1u;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t st_base = smem_base + (uint32_t)s * stage_bytes;
        const uint64_t adesc = make_desc_mn_sw128(st_base, kWgBlockBytes);
        const uint64_t bdesc = make_desc_mn_sw128(st_base + a_bytes, kWgBlockBytes);
#pragma unroll
        for (int k = 0; k < kWgPixBlock / 16; ++k)   // 16 pixels = two 8-row groups = 2048 B = 128 x 16 B
          umma_bf16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (uint32_t)((kb | k) != 0));
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(&accum_bar));
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

}  // namespace

int rd_wgrad_tc_supported(const rd_conv_desc* d) {
  if (d->dtype != RD_BF16) return 0;
  if (d->cin % 8 || d->cout % 8 || d->cin < 8 || d->cout < 8) return 0;
  return 1;
}

int rd_wgrad_tc_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                       cudaStream_t st) {
  WgParams P;
  P.x = (const bf16*)x; P.dy = (const bf16*)dy; P.dK = dK; P.dbias = dbias;
  P.dbias_gpr = d->bias_groups > 1 ? d->groups / d->bias_groups : 0;
  P.H = d->h; P.W = d->w; P.Cin = d->cin; P.OH = d->oh; P.OW = d->ow; P.Cout = d->cout;
  P.KH = d->kh; P.KW = d->kw; P.stride = d->stride; P.pad = d->pad;
  const int ipg = d->n / d->groups;
  P.ppg = (int64_t)ipg * d->oh * d->ow;
  P.n_total = d->kh * d->kw * d->cin;
  P.n_ext = P.n_total + (dbias ? 8 : 0);
  P.transposed = d->cout < 128 ? 1 : 0;
  P.m_blocks = 2;
  int grid_y;
  if (!P.transposed) {
    grid_y = rd_div_up(d->cout, 128);
    P.n_splits = rd_div_up(P.n_ext, 256);
    int per = rd_div_up(P.n_ext, P.n_splits);
    P.np_per_cta = ((per + 63) / 64) * 64;
    if (P.np_per_cta > 256) P.np_per_cta = 256;
    P.n_splits = rd_div_up(P.n_ext, P.np_per_cta);
    P.n_blocks = P.np_per_cta / 64;
    P.n_width = P.np_per_cta;
  } else {
    int m_tiles = rd_div_up(P.n_ext, 128);
    P.n_splits = 1;
    grid_y = m_tiles;
    P.np_per_cta = 128;
    P.n_blocks = rd_div_up(d->cout, 64);
    P.n_width = ((d->cout + 15) / 16) * 16;
  }
  int cols = 32;
  while (cols < P.n_width) cols <<= 1;
  P.tmem_cols = cols;
  // pixel chunk: enough CTAs to fill the machine a few times, at least 4 K-blocks each
  int64_t other = (int64_t)grid_y * P.n_splits * d->groups;
  int64_t want = (int64_t)ctx->sm_count * 4;
  int64_t chunks = (want + other - 1) / other;
  int64_t max_chunks = (P.ppg + 4 * kWgPixBlock - 1) / (4 * kWgPixBlock);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t cp = (P.ppg + chunks - 1) / chunks;
  cp = ((cp + kWgPixBlock - 1) / kWgPixBlock) * kWgPixBlock;
  P.chunk_pixels = (int)cp;
  P.chunks_pg = (int)((P.ppg + cp - 1) / cp);
  {
    static const bool no_fast = getenv("RD_B200_TC_NO_FAST") != nullptr;
    const int64_t x_elems = (int64_t)d->n * d->h * d->w * d->cin, dy_elems = (int64_t)d->n * d->oh * d->ow * d->cout;
    P.fast = !no_fast && x_elems < (1ll << 31) && dy_elems < (1ll << 31);
  }
  size_t smem = (size_t)kWgStages * (P.m_blocks + P.n_blocks) * kWgBlockBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(P.chunks_pg * P.n_splits, grid_y, d->groups);
  k_wgrad_tc<<<grid, kThreads, smem, st>>>(P);
  RD_CHECK_LAUNCH(ctx, "wgrad_tc");
  return RD_OK;
}
