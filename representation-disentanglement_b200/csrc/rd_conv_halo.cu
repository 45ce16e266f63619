// rd_conv_halo.cu — halo-tile tcgen05 implicit-GEMM convolution for the 3x3 stride-1 "same" convolutions whose packed
// weights fit in shared memory (the full- and half-resolution SPADE layers sp5 / sp6, their si_layers, the last two
// anatomy-decoder blocks): forward and dgrad.
//
// Why: in k_conv_tma every filter tap re-loads the 128-pixel activation tile (shifted by one pixel) and the weight
// tile through L2, 9x the algorithmic bytes; at <= 64 channels that makes the kernel L2->SM bound (ncu: tensor pipe
// 8 %, ~35 B/clk/SM of TMA traffic, profiles/r01_ncu_full_conv_tma_sp6gb.txt).  Here
//   * the weights of the CTA's current group stay RESIDENT in shared memory (TMA, K-major swizzled boxes, reloaded only
//     when the persistent CTA crosses a group boundary of its contiguous tile range), and
//   * the activation tile is loaded ONCE with its halo: (16+2) x (8+2) pixels, stored channel-block-major
//     [c/8][18*10 pixels][8 channels] = the canonical NO-SWIZZLE K-major UMMA layout (core matrix = 8 pixels x 16 B,
//     contiguous).  The A operand of tap (kh, kw) is the same buffer with the descriptor start address advanced by
//     (kh*10 + kw) pixels: 8-pixel core-matrix groups are one halo row (SBO = 10*16 B) apart, 8-channel planes
//     LBO = 180*16 B apart.  9 taps x C/16 MMAs read one 11.5 KB (C=32) tile instead of 9 x 8 KB.
//
// Roles (416 threads): warps 0-3 and 9-12 = two epilogue groups, one per TMEM accumulator buffer (even / odd tiles;
// tcgen05.ld, bias + LeakyReLU, bf16, swizzled staging rows, TMA store), warps 4-7 halo producers (cp.async 16 B,
// zero-fill outside the image = the conv padding), warp 8 = TMEM allocator, weight TMA and MMA issue (one elected lane).
// Pipelines: halo stages full/empty, TMEM accumulator double buffer.  Two epilogue groups because one warp per SM
// sub-partition is latency bound on its own instruction stream (ncu: epilogue warps 88 % busy while the tensor pipe
// idles half of the time).
#include <cuda.h>
#include <stdlib.h>
#include "rd_common.cuh"
#include "rd_tc_common.cuh"

namespace {

constexpr int kHThreads = 416;                    // 13 warps, see the role list above
constexpr int kHTH = 16, kHTW = 8;                 // output tile: 16 rows x 8 columns = 128 pixels (UMMA M)
constexpr int kHHW = kHTW + 2, kHHH = kHTH + 2;    // halo tile 18 x 10
constexpr int kHPix = kHHW * kHHH;                 // 180 pixels
constexpr uint32_t kHPlane = kHPix * 16u;          // one 8-channel plane of the halo tile: 2880 B
constexpr int kHMaxStages = 6;
constexpr int kHMaxAcc = 4;                         // TMEM accumulator buffers in flight (n_acc * n_tile <= 512 columns)

struct HaloParams {
  const bf16* x; const float* bias; bf16* y;
  int H, W, Cin, Cout;
  int sign;                       // +1 forward (tap reads out + k - 1), -1 dgrad (out + 1 - k)
  int ipg;                        // images per weight group
  int tiles_x, tiles_per_img, total_tiles, tiles_per_cta;
  int n_tile;                     // UMMA N
  int kc, chunks;                 // channels per halo stage, Cin / kc
  int w_boxes;                    // 64-column weight boxes: ceil(9 * Cin / 64)
  uint32_t w_box_bytes, w_bytes, w_tx_bytes;
  uint32_t a_stage_bytes;
  int n_acc, acc_shift;           // TMEM accumulator buffers (power of two) and log2
  int stages, lag;                // lag = cp.async groups a producer thread keeps in flight (< stages)
  uint32_t stg_off;               // staging buffers for the TMA-store epilogue (offset from the 1 KB aligned base)
  int stg_bufs, store_cw;         // 0 buffers = direct st.global epilogue; store_cw = channels per store box (<= 64)
  uint32_t tmem_cols;
  int act; float slope;
  int bias_gpr;                   // weight groups per bias row (0: one bias row for all groups)
};

__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// K-major NO-SWIZZLE descriptor: LBO = byte offset between core matrices along K, SBO = along M (8-row groups)
__device__ __forceinline__ uint64_t make_desc_k_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return lo | (hi << 32);
}

// all MMAs of one (tile, channel chunk): 9 taps x KSTEPS 16-channel steps.  The weights sit in shared memory as
// [n_tile][9*Cin] split into 64-element (128-byte, SWIZZLE_128B) column boxes whatever Cin is: the narrower swizzle modes
// (64-byte rows for Cin = 32, 32-byte rows for Cin = 16) make the tensor core's B-operand reads 2-3x slower (measured:
// 37 / 65-73 / 104 cycles per M128 MMA with 128 / 64 / 32-byte weight rows).
template <int KSTEPS>
__device__ __forceinline__ void issue_taps(uint32_t tacc, uint64_t ad, uint64_t bd0, uint32_t idesc, const uint32_t (&tap_off)[9],
                                           uint32_t kk0, uint32_t cin, uint32_t wbox16, bool accumulate) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) {
      const uint32_t kk = kk0 + (uint32_t)tap * cin + 16u * (uint32_t)k;          // position in the 9*Cin reduction axis
      umma_bf16(tacc, ad + (uint64_t)(tap_off[tap] + (uint32_t)k * ((2u * kHPlane) >> 4)),
                bd0 + (uint64_t)((kk >> 6) * wbox16 + ((kk & 63u) >> 3)), idesc, (uint32_t)(accumulate || tap != 0 || k != 0));
    }
  }
}

__global__ void __launch_bounds__(kHThreads, 1)
k_conv_halo(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapY, const HaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kHMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kHMaxStages];
  __shared__ __align__(8) uint64_t acc_full[kHMaxAcc];
  __shared__ __align__(8) uint64_t acc_empty[kHMaxAcc];
  __shared__ __align__(8) uint64_t w_full, w_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[2][256];       // one copy per epilogue group (they may be on different weight groups)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // weights first (1 KB aligned boxes)
  const uint32_t a_base = smem_base + P.w_bytes;
  const int S = P.stages;
  const int t_begin = blockIdx.x * P.tiles_per_cta;
  int t_end = t_begin + P.tiles_per_cta;
  if (t_end > P.total_tiles) t_end = P.total_tiles;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&mapB);
      if (P.stg_bufs) tma_prefetch_desc(&mapY);
      for (int s = 0; s < S; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 128);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int b = 0; b < kHMaxAcc; ++b) {
        mbar_init(smem_u32(&acc_full[b]), 1);
        mbar_init(smem_u32(&acc_empty[b]), 4);
      }
      mbar_init(smem_u32(&w_full), 1);
      mbar_init(smem_u32(&w_free), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ halo producers (128 threads)
    const int ptid = tid - 128;
    const int nb = P.kc >> 3;                       // 16-byte channel blocks per stage (2, 4 or 8)
    const int nb_shift = nb == 8 ? 3 : (nb == 4 ? 2 : 1);
    const int items = kHPix * nb;
    const int lag = P.lag;
    // the per-thread copy list is the same for every tile: precompute (destination offset, source offset relative to the
    // halo origin, halo row / column) once, so the per-tile loop is bounds test + add + cp.async
    constexpr int kMaxIt = (kHPix * 8 + 127) / 128;       // 12
    uint32_t dst_off[kMaxIt];
    int src_off[kMaxIt], hyx[kMaxIt];
#pragma unroll
    for (int j = 0; j < kMaxIt; ++j) {
      const int i = ptid + 128 * j;
      const int hp = i >> nb_shift, cb = i & (nb - 1);
      const int hy = hp / kHHW, hx = hp - hy * kHHW;
      dst_off[j] = (uint32_t)cb * kHPlane + (uint32_t)hp * 16u;
      src_off[j] = (hy * P.W + hx) * P.Cin + cb * 8;
      hyx[j] = i < items ? ((hy << 8) | hx) : -1;
    }
    int fill = 0, stage = 0, done_stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int img = t / P.tiles_per_img;
      const int rem = t - img * P.tiles_per_img;
      const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
      const int y0 = ty * kHTH - 1, x0 = tx * kHTW - 1;
      const bf16* xt = P.x + ((int64_t)img * P.H * P.W + (int64_t)y0 * P.W + x0) * P.Cin;   // halo origin (may lie outside)
      for (int c = 0; c < P.chunks; ++c, ++fill) {
        const int s = stage;
        mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1u);
        const uint32_t a_s = a_base + (uint32_t)s * P.a_stage_bytes;
        const bf16* xc = xt + c * P.kc;
#pragma unroll
        for (int j = 0; j < kMaxIt; ++j) {
          if (hyx[j] >= 0) {
            const bool v = ((unsigned)(y0 + (hyx[j] >> 8)) < (unsigned)P.H) && ((unsigned)(x0 + (hyx[j] & 255)) < (unsigned)P.W);
            cp_async16(a_s + dst_off[j], v ? (const void*)(xc + src_off[j]) : (const void*)P.x, v ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (fill >= lag) {
          if (lag == 1) cp_async_wait<1>(); else if (lag == 2) cp_async_wait<2>(); else cp_async_wait<3>();
          fence_proxy_async();
          mbar_arrive(smem_u32(&full_bar[done_stage]));
          if (++done_stage == S) done_stage = 0;
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int f = (fill > lag ? fill - lag : 0); f < fill; ++f) {
      mbar_arrive(smem_u32(&full_bar[done_stage]));
      if (++done_stage == S) done_stage = 0;
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ weights + MMA issuer
    // The whole warp runs this (warp-uniform control flow and values); one elected lane issues the TMA / tcgen05
    // instructions.  Under `if (lane == 0)` the compiler treats every descriptor as per-thread data and wraps each
    // tcgen05.mma in an R2UR / ELECT / BRA.U.ANY serialisation loop — the single issuing thread then cannot keep the
    // tensor core fed (ncu: tensor pipe 48 % active, issuer 57 % busy executing, profiles/r01_ncu_conv_halo.txt).
    {
      const uint32_t idesc = make_idesc(128, P.n_tile);
      const int ksteps = P.kc >> 4;
      // descriptors are affine in (stage, tap, k-step): build the two bases once, add 16-byte-unit offsets in the loop
      const uint64_t adesc0 = make_desc_k_nosw(a_base, kHPlane, kHHW * 16u);
      const uint64_t bdesc0 = make_desc_k(smem_base, 128u);
      const uint32_t wbox16 = P.w_box_bytes >> 4;
      uint32_t tap_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - kh * 3;
        tap_off[tap] = (uint32_t)(P.sign > 0 ? kh * kHHW + kw : (2 - kh) * kHHW + (2 - kw));
      }
      int stage = 0, it = 0, cur_g = -1;
      uint32_t phase = 0, wphase = 0, fphase = 0;
      const int tiles_per_group = P.tiles_per_img * P.ipg;
      int g = t_begin / tiles_per_group;
      int g_left = tiles_per_group - (t_begin - g * tiles_per_group);     // tiles left in group g
      for (int t = t_begin; t < t_end; ++t, ++it) {
        if (g_left == 0) { ++g; g_left = tiles_per_group; }
        --g_left;
        if (g != cur_g) {
          if (cur_g >= 0) {                          // every MMA that reads the old weights must have completed
            if (elect_one()) umma_commit(smem_u32(&w_free));
            __syncwarp();
            mbar_wait(smem_u32(&w_free), fphase);
            fphase ^= 1u;
          }
          const uint32_t wb = smem_u32(&w_full);
          if (elect_one()) {
            mbar_arrive_expect_tx(wb, P.w_tx_bytes);
            for (int b = 0; b < P.w_boxes; ++b)
              tma_load_2d(smem_base + (uint32_t)b * P.w_box_bytes, &mapB, b * 64, g * P.Cout, wb);
          }
          __syncwarp();
          mbar_wait(wb, wphase);
          wphase ^= 1u;
          cur_g = g;
        }
        const int buf = it & (P.n_acc - 1);
        mbar_wait(smem_u32(&acc_empty[buf]), (((uint32_t)it >> P.acc_shift) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)buf * (uint32_t)P.n_tile;
        for (int c = 0; c < P.chunks; ++c) {
          const int s = stage;
          mbar_wait(smem_u32(&full_bar[s]), phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(((uint32_t)s * P.a_stage_bytes) >> 4);
          const uint32_t kk0 = (uint32_t)(c * P.kc);
          if (elect_one()) {
            if (ksteps == 2) issue_taps<2>(tacc, ad, bdesc0, idesc, tap_off, kk0, (uint32_t)P.Cin, wbox16, c != 0);
            else if (ksteps == 4) issue_taps<4>(tacc, ad, bdesc0, idesc, tap_off, kk0, (uint32_t)P.Cin, wbox16, c != 0);
            else issue_taps<1>(tacc, ad, bdesc0, idesc, tap_off, kk0, (uint32_t)P.Cin, wbox16, c != 0);
            umma_commit(smem_u32(&empty_bar[s]));
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(smem_u32(&acc_full[buf]));
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group 0 = warps 0-3 (even tiles, TMEM
    // buffer 0), group 1 = warps 9-12 (odd tiles, buffer 1); warp & 3 = the TMEM lane quarter the warp may read
    const int grp = warp >= 9 ? 1 : 0;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tyl = row >> 3, txl = row & 7;
    const uint32_t stg_base = smem_base + P.stg_off;
    const float slope = P.act == RD_ACT_LRELU ? P.slope : 1.f;      // max(v, slope * v) == LeakyReLU for 0 < slope <= 1
    int bias_g = -1;
    for (int t = t_begin + grp; t < t_end; t += 2) {
      const int it = t - t_begin;
      {   // (re)load this epilogue group's bias row when the tile's weight group changes (named barrier 1 + grp, 128 threads)
        const int tg0 = (t / P.tiles_per_img) / P.ipg;
        const int tg = P.bias_gpr ? tg0 / P.bias_gpr : 0;            // bias row of the tile's weight group
        if (tg != bias_g) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          const float* src = P.bias ? P.bias + (size_t)tg * P.Cout : nullptr;
          for (int i = (warp & 3) * 32 + lane; i < 256; i += 128) bias_s[grp][i] = (src != nullptr && i < P.Cout) ? src[i] : 0.f;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          bias_g = tg;
        }
      }
      const int buf = it & (P.n_acc - 1);
      const int img = t / P.tiles_per_img;
      const int rem = t - img * P.tiles_per_img;
      const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
      const int gy = ty * kHTH + tyl, gx = tx * kHTW + txl;
      const bool pvalid = gy < P.H && gx < P.W;
      bf16* yrow = P.y + (pvalid ? (((int64_t)img * P.H + gy) * P.W + gx) : 0) * P.Cout;
      mbar_wait(smem_u32(&acc_full[buf]), ((uint32_t)it >> P.acc_shift) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * (uint32_t)P.n_tile;
      if (P.stg_bufs) {
        // TMEM -> registers -> swizzled staging rows -> one TMA store per warp and 64-channel block.  (Lane-per-pixel
        // st.global touches 32 different 128-byte lines per instruction.)  All tcgen05.ld of a block are issued before
        // one wait, so their latencies overlap.
        const int cw = P.store_cw;
        const uint32_t rb = (uint32_t)cw * 2u;                       // staging row bytes: 128 / 64 / 32
        const uint32_t swz_mask = (uint32_t)(cw >> 3) - 1u;
        const uint32_t swz = (((uint32_t)lane * rb) >> 7) & swz_mask;
        const uint32_t wst = stg_base + (uint32_t)((P.stg_bufs == 2 ? grp : 0) * 4 + q) * 32u * rb;
        const uint32_t rowa = wst + (uint32_t)lane * rb;
        const int ngrp = cw >> 4;                                    // 16-column register groups per block: 1, 2 or 4
        for (int c0 = 0; c0 < P.Cout; c0 += cw) {
          uint32_t r[64];
#pragma unroll
          for (int gi = 0; gi < 4; ++gi)
            if (gi < ngrp) tmem_ld16_nowait(taddr + (uint32_t)(c0 + gi * 16), r + gi * 16);
          if (lane == 0) bulk_wait_read<0>();                        // this warp's previous store has read the staging rows
          __syncwarp();
          tmem_ld_wait();
          if (c0 + cw >= P.Cout) {                                   // accumulator drained: hand the TMEM buffer back NOW
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
          }
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            if (gi < ngrp) {
              tmem_ld_fence16(r + gi * 16);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int cl = gi * 16 + h * 8;
                const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[grp][c0 + cl]);
                const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[grp][c0 + cl + 4]);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t packed[4];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                  float v0 = __uint_as_float(r[cl + 2 * qq]) + bb[2 * qq], v1 = __uint_as_float(r[cl + 2 * qq + 1]) + bb[2 * qq + 1];
                  v0 = fmaxf(v0, v0 * slope); v1 = fmaxf(v1, v1 * slope);
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                  packed[qq] = *reinterpret_cast<uint32_t*>(&b2);
                }
                const uint32_t chunk = (uint32_t)(cl >> 3);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((chunk ^ swz) << 4)), "r"(packed[0]),
                             "r"(packed[1]), "r"(packed[2]), "r"(packed[3]) : "memory");
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&mapY, wst, c0, tx * kHTW, ty * kHTH + q * 4, img);
            bulk_commit();
          }
        }
        continue;
      }
      for (int cb = 0; cb < P.n_tile; cb += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)cb, r);
        if (pvalid && (P.Cout & 7)) {
#pragma unroll
          for (int qq = 0; qq < 16; ++qq) {
            int co = cb + qq;
            if (co < P.Cout) {
              float v0 = __uint_as_float(r[qq]);
              v0 += bias_s[grp][co];
              if (P.act == RD_ACT_LRELU) v0 = v0 > 0.f ? v0 : v0 * P.slope;
              yrow[co] = __float2bfloat16_rn(v0);
            }
          }
        } else if (pvalid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int co = cb + h * 8;
            if (co < P.Cout) {
              uint32_t packed[4];
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                float v0 = __uint_as_float(r[h * 8 + 2 * qq]), v1 = __uint_as_float(r[h * 8 + 2 * qq + 1]);
                v0 += bias_s[grp][co + 2 * qq]; v1 += bias_s[grp][co + 2 * qq + 1];
                if (P.act == RD_ACT_LRELU) { v0 = v0 > 0.f ? v0 : v0 * P.slope; v1 = v1 > 0.f ? v1 : v1 * P.slope; }
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                packed[qq] = *reinterpret_cast<uint32_t*>(&b2);
              }
              *reinterpret_cast<uint4*>(yrow + co) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    }
    if (P.stg_bufs && lane == 0) bulk_wait_read<0>();     // staging rows must stay valid until the last store has read them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
  }
}

bool g_halo_attr_set = false;
constexpr uint32_t kHaloSmemMax = 223u * 1024u;

struct HaloPlan {
  int cin, cout, n_tile, kc, chunks, w_boxes, stages, stg_bufs, store_cw;
  uint32_t w_box_bytes, w_bytes, a_stage_bytes, stg_off;
};

bool halo_plan(const rd_conv_desc* d, int mode, HaloPlan& pl) {
  if (d->dtype != RD_BF16) return false;
  if (d->stride != 1 || d->kh != 3 || d->kw != 3 || d->pad != 1) return false;
  pl.cin = mode == 0 ? d->cin : d->cout;
  pl.cout = mode == 0 ? d->cout : d->cin;
  if (pl.cin % 16 || pl.cin < 16) return false;
  pl.n_tile = (pl.cout + 15) / 16 * 16;
  if (pl.n_tile > 256) return false;
  pl.kc = (pl.cin % 64 == 0) ? 64 : ((pl.cin % 32 == 0) ? 32 : 16);
  pl.chunks = pl.cin / pl.kc;
  pl.w_boxes = (9 * pl.cin + 63) / 64;
  pl.w_box_bytes = (uint32_t)pl.n_tile * 128u;                            // n_tile % 16 == 0 -> multiple of 1 KB
  pl.w_bytes = (uint32_t)pl.w_boxes * pl.w_box_bytes;
  pl.a_stage_bytes = (uint32_t)(pl.kc / 8) * kHPlane;                     // multiple of 16 B
  // TMA-store epilogue: 16 / 32 / 64 output channels, or a multiple of 64 (one store box per 64-channel block)
  pl.store_cw = pl.cout >= 64 ? 64 : pl.cout;
  const bool can_store = (pl.cout % 64 == 0) || pl.cout == 32 || pl.cout == 16;
  const uint32_t avail = kHaloSmemMax - 1024u;
  for (int bufs = can_store ? 2 : 0; bufs >= 0; bufs -= 2) {     // one staging buffer per epilogue group, or direct stores
    uint32_t stg = (uint32_t)bufs * 128u * (uint32_t)pl.store_cw * 2u + (bufs ? 1024u : 0u);     // + alignment slack
    if (pl.w_bytes + stg + 2u * pl.a_stage_bytes > avail) continue;
    int st = (int)((avail - pl.w_bytes - stg) / pl.a_stage_bytes);
    pl.stages = st > kHMaxStages ? kHMaxStages : st;
    pl.stg_bufs = bufs;
    pl.stg_off = (pl.w_bytes + (uint32_t)pl.stages * pl.a_stage_bytes + 1023u) & ~1023u;
    return true;
  }
  return false;
}

}  // namespace

int rd_conv_halo_supported(const rd_conv_desc* d, int mode, int sm_count, int forced) {
  HaloPlan pl;
  if (!halo_plan(d, mode, pl)) return 0;
  if (!rd_tensormap_encode_fn()) return 0;
  if (forced) return 1;
  static const bool off = getenv("RD_B200_NO_HALO") != nullptr;
  if (off) return 0;
  // worth it only when every persistent CTA amortises its weight load over several tiles
  int64_t tiles = (int64_t)d->n * rd_div_up(d->h, kHTH) * rd_div_up(d->w, kHTW);
  return tiles >= 4 * (int64_t)sm_count;
}

int rd_conv_halo_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                        cudaStream_t st) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = (EncodeTiledFn)rd_tensormap_encode_fn();
  HaloPlan pl;
  if (!enc || !halo_plan(d, mode, pl)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv_halo: shape not supported");
  HaloParams P;
  P.x = (const bf16*)x; P.bias = bias; P.y = (bf16*)y;
  P.H = d->h; P.W = d->w; P.Cin = pl.cin; P.Cout = pl.cout;
  P.sign = mode == 0 ? 1 : -1;
  P.ipg = d->n / d->groups;
  P.tiles_x = rd_div_up(d->w, kHTW);
  P.tiles_per_img = P.tiles_x * rd_div_up(d->h, kHTH);
  P.total_tiles = d->n * P.tiles_per_img;
  int grid = P.total_tiles < ctx->sm_count ? P.total_tiles : ctx->sm_count;
  P.tiles_per_cta = rd_div_up(P.total_tiles, grid);
  grid = rd_div_up(P.total_tiles, P.tiles_per_cta);
  P.n_tile = pl.n_tile; P.kc = pl.kc; P.chunks = pl.chunks; P.w_boxes = pl.w_boxes;
  P.w_box_bytes = pl.w_box_bytes; P.w_bytes = pl.w_bytes;
  P.w_tx_bytes = pl.w_bytes;
  P.a_stage_bytes = pl.a_stage_bytes;
  P.stages = pl.stages;
  P.lag = pl.stages - 1 < 3 ? pl.stages - 1 : 3;
  P.stg_off = pl.stg_off; P.stg_bufs = pl.stg_bufs; P.store_cw = pl.store_cw;
  P.n_acc = 2; P.acc_shift = 1;
  while (P.n_acc < kHMaxAcc && 2 * P.n_acc * pl.n_tile <= 512) { P.n_acc *= 2; ++P.acc_shift; }
  uint32_t cols = 32;
  while (cols < (uint32_t)(P.n_acc * pl.n_tile)) cols <<= 1;
  P.tmem_cols = cols;
  P.act = mode == 0 ? d->act : RD_ACT_NONE;
  P.slope = d->act_slope;
  P.bias_gpr = (mode == 0 && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;

  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  alignas(64) CUtensorMap mapB;
  {
    const int k_total = 9 * pl.cin;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->groups * pl.cout};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)pl.n_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(halo weights) failed: %d", (int)r);
  }
  alignas(64) CUtensorMap mapY;
  memset(&mapY, 0, sizeof(mapY));
  if (pl.stg_bufs) {
    CUtensorMapSwizzle swy = pl.store_cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (pl.store_cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[4] = {(cuuint64_t)pl.cout, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)pl.cout * 2, (cuuint64_t)d->w * pl.cout * 2, (cuuint64_t)d->h * d->w * pl.cout * 2};
    cuuint32_t box[4] = {(cuuint32_t)pl.store_cw, (cuuint32_t)kHTW, 4u, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&mapY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swy,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(halo output) failed: %d", (int)r);
  }
  size_t smem = (pl.stg_bufs ? (size_t)pl.stg_off + (size_t)pl.stg_bufs * 128 * pl.store_cw * 2 : (size_t)pl.w_bytes + (size_t)pl.stages * pl.a_stage_bytes) + 1024;
  if (!g_halo_attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_conv_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    g_halo_attr_set = true;
  }
  k_conv_halo<<<grid, kHThreads, smem, st>>>(mapB, mapY, P);
  RD_CHECK_LAUNCH(ctx, mode == 0 ? "conv_halo_fwd" : "conv_halo_dgrad");
  return RD_OK;
}
