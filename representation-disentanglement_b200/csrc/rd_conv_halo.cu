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
//   * the halo tile arrives as ONE 5-D TMA box per stage: the NHWC tensor is described as (8 channels, W, H, C/8, N) with a
//     16-byte stride on the channel-block dimension, so the box (8, halo columns, halo rows, blocks, 1) lands as the planes
//     above; out-of-image pixels are zero-filled by the TMA unit (= the conv padding).  The first version gathered the tile
//     with 16-byte cp.async copies from four producer warps: every copy was its own shared-memory wavefront (631 per tile
//     against 87 ideal) and the kernel was bound by the shared-memory pipe (profiles/r01_ncu_full_conv_halo_nt2_sp6gb_v8.txt);
//     that path remains behind RD_B200_HALO_TMA=0.
//
// Roles (416 threads): warps 0-3 and 9-12 = two epilogue groups, one per TMEM accumulator buffer (even / odd tiles;
// tcgen05.ld, bias + LeakyReLU, bf16, swizzled staging rows, TMA store), warp 4 lane 0 = halo TMA issue (warps 4-7 are the
// cp.async producers of the fallback path), warp 8 = TMEM allocator, weight TMA and MMA issue (one elected lane).
// Pipelines: halo stages full/empty, TMEM accumulator double buffer.  Two epilogue groups because one warp per SM
// sub-partition is latency bound on its own instruction stream (ncu: epilogue warps 88 % busy while the tensor pipe
// idles half of the time).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "rd_common.cuh"
#include "rd_tc_common.cuh"

namespace {

constexpr int kHThreads = 416;                    // 13 warps, see the role list above
constexpr int kHTH = 16, kHTW = 8;                 // output tile: 16 rows x 8 columns = 128 pixels (UMMA M)
constexpr int kHHW = kHTW + 2, kHHH = kHTH + 2;    // halo tile 18 x 10
constexpr int kHPix = kHHW * kHHH;                 // 180 pixels (one 8-channel plane of the halo tile: 2880 B)
// NT = 128-pixel tiles per halo stage: 1 (16 x 8 pixels, halo 18 x 10) or 2 (two tiles side by side, halo 18 x 18).  With two
// tiles per stage the producers and the MMA warp pay their barrier round trips once per 256 pixels and the halo overhead falls
// from 41 % to 27 %; used when the (twice as large) stages still fit next to the resident weights.
template <int NT> struct HaloGeom {
  static constexpr int TW = 8 * NT, HHW = TW + 2, HPix = HHW * (kHTH + 2);
  static constexpr uint32_t Plane = (uint32_t)HPix * 16u;
};
constexpr int kHMaxStages = 12;
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
  int dbg;                        // timing experiments only (RD_B200_HALO_DEBUG): 1 = producers skip the copies, 2 = no MMAs, 4 = no epilogue stores
  uint32_t sleep_epi, sleep_mma, sleep_prod;   // nanoseconds between barrier polls of the waiting roles
  int pgroups;                    // producer groups (2: alternate tiles; 1 when the stage ring is too short)
  uint32_t stg_off;               // staging buffers for the TMA-store epilogue (offset from the 1 KB aligned base)
  int stg_bufs, store_cw;         // 0 buffers = direct st.global epilogue; store_cw = channels per store box (<= 64)
  uint32_t tmem_cols;
  int act; float slope;
  int bias_gpr;                   // weight groups per bias row (0: one bias row for all groups)
  int use_tma;                    // halo tiles by ONE 5-D TMA box per stage (default; RD_B200_HALO_TMA=0: the cp.async producers)
  int dual;                       // two MMA-issuing warps (8 and 5), one per tile of a two-tile stage: see halo_mma
  int narrow;                     // coalesced narrow-output epilogue: 8 x 512 B of staging rows at stg_off
  int zpf;                        // SPADE epilogue: prefetch the next tile's z into the L2 (RD_B200_HALO_ZPF=0: off)
  // SPADE modulation fused into the gamma|beta convolution (reference src/model.py:2444-2452): Cout = 2C, accumulator columns
  // [0, C) = gamma, [C, 2C) = beta; the epilogue reads z and writes gamma (saved for the backward) and
  // mix = (z - mean) * invstd * (1 + gamma) + beta — the [N, H, W, 2C] gamma|beta tensor and the separate modulation pass disappear
  int spade;                      // C (0 = plain convolution)
  const bf16* z; const float* mean; const float* invstd;      // z [N, H, W, C]; InstanceNorm statistics [N, C]
  bf16* gamma; bf16* mix;
};

__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// K-major NO-SWIZZLE descriptor: LBO = byte offset between core matrices along K, SBO = along M (8-row groups)
__device__ __forceinline__ uint64_t make_desc_k_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return lo | (hi << 32);
}

// tcgen05.mma with the descriptors given as (low, high) words: the high words (SBO / version / layout) never change, the low
// words (start address, LBO) are one 32-bit add away from loop-invariant registers.
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
}

// Weight TMA + MMA issue (warp 8).  The whole warp runs this (warp-uniform control flow and values); one elected lane issues
// the TMA / tcgen05 instructions.  Under `if (lane == 0)` the compiler treats every descriptor as per-thread data and wraps each
// tcgen05.mma in an R2UR / ELECT / BRA.U.ANY serialisation loop.  All MMAs of one (tile, channel chunk) are 9 taps x KSTEPS
// 16-channel steps.  The weights sit in shared memory as [n_tile][9*Cin] split into 64-element (128-byte, SWIZZLE_128B) column
// boxes whatever Cin is: the narrower swizzle modes (64-byte rows for Cin = 32, 32-byte rows for Cin = 16) make the tensor
// core's B-operand reads 2-3x slower (measured: 37 / 65-73 / 104 cycles per M128 MMA with 128 / 64 / 32-byte weight rows).
// Every per-MMA descriptor offset is loop invariant (A: tap shift inside the halo tile; B: position of (tap, k-step) in the
// resident weights) and is computed ONCE into registers: the tile loop is one add per operand and the MMA (ncu on the first
// version: 430 instructions per tile for 18 MMAs — longer than the MMAs themselves at N <= 64).
//
// TWO ISSUERS (P.dual, two-tile stages): a tcgen05.mma occupies its issuing thread until the tensor core has taken it — the thread's
// other instructions (barrier polls, fences, commits, the S2UR of every commit's cluster address) are NOT overlapped with its own MMAs,
// they add to them (tools/mma_rate.cu, profiles/r02_mma_issue_microbench.txt: one issuer 66 -> 209 cycles per N = 32 MMA as 0 -> 96 dependent
// integer operations follow every 4 MMAs; two issuing warps run at the sum of their rates until the tensor pipe's own 40 / 48 / 64
// cycles at N = 32 / 64 / 128).  With `iss` = 0 (warp 8) issuing the stage's first tile and `iss` = 1 (warp 5) the second, each
// issuer's per-stage overhead hides behind the other's MMAs.  Both wait for the stage, each commits its own accumulator, and the stage
// is released by both commits (empty barrier count 2); at a weight-group boundary both commit `w_free` (count 2), issuer 0 reloads.
// One-tile stages (NT = 1): the issuers alternate tiles (= the epilogue groups' accumulator buffers); both still run the weight-group
// protocol at every boundary, whichever issuer owns the first tile behind it.
template <int KSTEPS, bool FASTB, int CIN, int SIGN, int NT>      // CIN > 0: Cin and the tap direction are compile-time -> every descriptor offset is an immediate
__device__ __forceinline__ void halo_mma(const HaloParams& P, uint32_t smem_base, uint32_t a_base, uint32_t tmem_base,
                                         uint64_t* full_bar, uint64_t* empty_bar, uint64_t* acc_full, uint64_t* acc_empty,
                                         uint64_t* w_full, uint64_t* w_free, const CUtensorMap* mapB, int t_begin, int t_end, int iss) {
  const bool dual = NT == 2 && P.dual != 0;
  const int h_lo = dual ? iss : 0, h_hi = dual ? iss + 1 : NT;        // tiles of a stage this issuer computes
  const int S = P.stages;
  const uint32_t idesc = make_idesc(128, P.n_tile);
  constexpr int HHW = HaloGeom<NT>::HHW;
  constexpr uint32_t Plane = HaloGeom<NT>::Plane;
  const uint64_t adesc0 = make_desc_k_nosw(a_base, Plane, HHW * 16u);
  const uint64_t bdesc0 = make_desc_k(smem_base, 128u);
  const uint32_t a_lo0 = (uint32_t)adesc0, a_hi = (uint32_t)(adesc0 >> 32);
  const uint32_t b_lo0 = (uint32_t)bdesc0, b_hi = (uint32_t)(bdesc0 >> 32);
  const uint32_t wbox16 = P.w_box_bytes >> 4;
  const uint32_t stage16 = P.a_stage_bytes >> 4;
  uint32_t aoff[9], boff[9 * KSTEPS];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int kh = tap / 3, kw = tap - kh * 3;
    aoff[tap] = (uint32_t)(P.sign > 0 ? kh * HHW + kw : (2 - kh) * HHW + (2 - kw));
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) {
      const uint32_t kk = (uint32_t)tap * (uint32_t)P.Cin + 16u * (uint32_t)k;     // position in the 9*Cin reduction axis (chunk 0)
      boff[tap * KSTEPS + k] = (kk >> 6) * wbox16 + ((kk & 63u) >> 3);
    }
  }
  // compile-time variant: B start of weight box m (<= 9 boxes per chunk), the offsets inside a box are immediates
  constexpr int kBoxes = CIN > 0 ? (9 * (CIN > 64 ? 64 : CIN) + 63) / 64 + (CIN > 64 ? 9 : 0) : 1;
  uint32_t wbm[kBoxes];
#pragma unroll
  for (int m = 0; m < kBoxes; ++m) wbm[m] = b_lo0 + (uint32_t)m * wbox16;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), accf0 = smem_u32(acc_full), acce0 = smem_u32(acc_empty);
  constexpr bool fast_b = FASTB;          // one chunk, or 64-channel chunks (= one weight box each): offsets are loop invariant
  int stage = 0, it = 0, cur_g = -1;
  uint32_t phase = 0, wphase = 0, fphase = 0;
  const int tiles_per_group = P.tiles_per_img * P.ipg;
  int g = t_begin / tiles_per_group;
  int g_left = tiles_per_group - (t_begin - g * tiles_per_group);     // tiles left in group g
  for (int t = t_begin; t < t_end; t += NT, it += NT) {
    if (g_left == 0) { ++g; g_left = tiles_per_group; }
    g_left -= NT;                                // tiles per image are a multiple of NT: a stage never straddles a weight group
    if (g != cur_g) {
      if (cur_g >= 0) {                          // every MMA that reads the old weights must have completed (both issuers' when dual)
        if (elect_one()) umma_commit(smem_u32(w_free));
        __syncwarp();
        if (iss == 0) {
          mbar_wait(smem_u32(w_free), fphase);
          fphase ^= 1u;
        }
      }
      const uint32_t wb = smem_u32(w_full);
      if (iss == 0) {
        if (elect_one()) {
          mbar_arrive_expect_tx(wb, P.w_tx_bytes);
          for (int b = 0; b < P.w_boxes; ++b)
            tma_load_2d(smem_base + (uint32_t)b * P.w_box_bytes, mapB, b * 64, g * P.Cout, wb);
        }
        __syncwarp();
      }
      mbar_wait(wb, wphase);
      wphase ^= 1u;
      cur_g = g;
    }
    // accumulators of the stage's tiles: buffers buf0 (and buf0 + 1: `it` is even when NT = 2, both share the barrier parity)
    const uint32_t buf0 = (uint32_t)it & (uint32_t)(P.n_acc - 1);
    const uint32_t accpar = ((((uint32_t)it) >> P.acc_shift) & 1u) ^ 1u;
    if (NT == 1 && P.dual && ((it & 1) != iss)) {     // one-tile stages: the issuers alternate tiles; skip the other issuer's stages
      for (int c = 0; c < P.chunks; ++c)
        if (++stage == S) { stage = 0; phase ^= 1u; }
      continue;
    }
    if (h_lo == 0) mbar_wait(acce0 + 8u * buf0, accpar);
    if (NT == 2 && h_hi == 2) mbar_wait(acce0 + 8u * buf0 + 8u, accpar);
    const uint32_t tacc0 = tmem_base + buf0 * (uint32_t)P.n_tile;
    const uint32_t accf_b = accf0 + 8u * buf0;
    tc_fence_after();
    for (int c = 0; c < P.chunks; ++c) {
      const int s = stage;
      mbar_wait_sleep(full0 + 8u * (uint32_t)s, phase, P.sleep_mma);
      // no proxy fence here: the barrier completes when the copies have been written to shared memory (same protocol as
      // CUTLASS' sm100 cp.async mainloop); a fence.proxy.async in this warp compiles to MEMBAR.ALL.CTA and stalls the MMA issue
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + (uint32_t)s * stage16;
      const uint32_t b_lo = b_lo0 + (fast_b ? (uint32_t)c * wbox16 : 0u);        // a 64-channel chunk = one weight box
      const bool last_chunk = c == P.chunks - 1;
      if (elect_one()) {
#pragma unroll
        for (int h = 0; h < NT; ++h) {
          if (h < h_lo || h >= h_hi) continue;                    // the other issuer's tile
          const uint32_t a_h = a_lo + (uint32_t)(h * 8);          // second tile: 8 halo columns (16-byte units) to the right
          const uint32_t tacc_h = tacc0 + (h ? (uint32_t)P.n_tile : 0u);
          if (CIN > 0 && !(P.dbg & 2)) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kh = tap / 3, kw = tap - kh * 3;
              const uint32_t ao = (uint32_t)(SIGN > 0 ? kh * HHW + kw : (2 - kh) * HHW + (2 - kw));
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                // CIN <= 64: one chunk, kk = tap*CIN + 16k;  CIN = 128: chunk c is box 2*tap + c of the 18, kk & 63 = 16k
                const int kk = tap * (CIN > 64 ? 64 : CIN) + 16 * k;
                const int m = CIN > 64 ? tap * (CIN / 64) : (kk >> 6);
                const uint32_t bsel = CIN > 64 ? wbm[m] + (uint32_t)c * wbox16 : wbm[m];
                umma_bf16_lh(tacc_h, a_h + ao + (uint32_t)k * ((2u * Plane) >> 4), a_hi, bsel + (uint32_t)((kk & 63) >> 3), b_hi, idesc,
                             (tap != 0 || k != 0) ? 1u : (uint32_t)(c != 0));
              }
            }
          } else if (CIN == 0 && !(P.dbg & 2)) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                uint32_t bo = boff[tap * KSTEPS + k];
                if (!fast_b) {                                           // several chunks narrower than a weight box (Cin = 48, 80, 96 ...)
                  const uint32_t kk = (uint32_t)(c * P.kc) + (uint32_t)tap * (uint32_t)P.Cin + 16u * (uint32_t)k;
                  bo = (kk >> 6) * wbox16 + ((kk & 63u) >> 3);
                }
                umma_bf16_lh(tacc_h, a_h + aoff[tap] + (uint32_t)k * ((2u * Plane) >> 4), a_hi, b_lo + bo, b_hi, idesc,
                             (tap != 0 || k != 0) ? 1u : (uint32_t)(c != 0));
              }
            }
          }
          if (last_chunk) umma_commit(accf_b + 8u * (uint32_t)h);      // this tile's accumulator is complete: its epilogue starts now
        }
        umma_commit(empty0 + 8u * (uint32_t)s);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
  }
}

__device__ __forceinline__ void cp_async16_full(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Halo producers (warps 4-7, 128 threads).  One warp per SM sub-partition runs this instruction stream alone, so its
// length is the tile period when the MMAs are short (ncu, sp6 out: 370 warp instructions per tile = 2500 cycles against
// 700 cycles of MMAs, producers never waiting): the copy list (destination, source offset, halo row / column) is built once
// per thread, tile coordinates advance without divisions, and tiles whose halo lies inside the image (3 of 4 at 160 x 192)
// take a path without bounds tests: one 64-bit add + one cp.async per 16 bytes.
template <int NB, int NT>
__device__ __forceinline__ void halo_producer(const HaloParams& P, uint32_t a_base, uint64_t* full_bar, uint64_t* empty_bar,
                                              int ptid, int pgrp, int t_begin, int t_end) {
  constexpr int HHW = HaloGeom<NT>::HHW, HPix = HaloGeom<NT>::HPix, TW = HaloGeom<NT>::TW;
  constexpr uint32_t Plane = HaloGeom<NT>::Plane;
  constexpr int kItems = HPix * NB;
  constexpr int kIt = (kItems + 127) / 128;            // copies per thread and stage (3 / 6 / 12 for one tile, 6 / 11 / 21 for two)
  constexpr int kShift = NB == 8 ? 3 : (NB == 4 ? 2 : 1);
  constexpr int kStep = 128 / NB;                      // halo pixels between two copies of a thread; its channel block never changes
  const int S = P.stages, PG = P.pgroups, chunks = P.chunks;
  if (pgrp >= PG) return;
  const int cb = ptid & (NB - 1), p0 = ptid >> kShift;
  const int hy0 = p0 / HHW, hx0 = p0 - hy0 * HHW;
  const uint32_t dst0 = (uint32_t)cb * Plane + (uint32_t)p0 * 16u;       // copy j lands at dst0 + j * kStep * 16
  int src_off[kIt];
  {
    int hy = hy0, hx = hx0;
#pragma unroll
    for (int j = 0; j < kIt; ++j) {
      src_off[j] = (hy * P.W + hx) * P.Cin + cb * 8;
      hx += kStep % HHW; hy += kStep / HHW;
      if (hx >= HHW) { hx -= HHW; ++hy; }
    }
  }
  const bool last_ok = p0 + kStep * (kIt - 1) < HPix;                    // the last copy slot is partial
  const int safe_off = (P.W + 1) * P.Cin;                     // halo pixel (1, 1) = output pixel (0, 0) of the tile: always inside
  const int tiles_x = P.tiles_x / NT;                          // stages per tile row
  const int tiles_y = P.tiles_per_img / P.tiles_x;
  const int t0 = t_begin / NT + pgrp;                          // stage index (NT tiles each)
  const int st_end = t_end / NT, st_per_img = P.tiles_per_img / NT;
  int img = t0 / st_per_img;
  int ty = (t0 - img * st_per_img) / tiles_x;
  int tx = t0 - img * st_per_img - ty * tiles_x;
  // the stage cursor walks the GLOBAL fill sequence (stage-major, chunk-minor), of which this group owns the stages t0 + k PG
  const int skip = (PG - 1) * chunks;                         // fills of the other group's stage between two own stages
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);      // once: the generic->shared conversion costs an S2UR chain
  int stage = pgrp * chunks;
  uint32_t phase = 0;
  while (stage >= S) { stage -= S; phase ^= 1u; }
  for (int t = t0; t < st_end; t += PG) {
    const int y0 = ty * kHTH - 1, x0 = tx * TW - 1;
    const bf16* xt = P.x + ((int64_t)(img * P.H + y0) * P.W + x0) * P.Cin;   // halo origin (may lie outside the image)
    const bool interior = ty > 0 && tx > 0 && y0 + kHTH + 2 <= P.H && x0 + HHW <= P.W;
    for (int c = 0; c < chunks; ++c) {
      mbar_wait_sleep(empty0 + 8u * (uint32_t)stage, phase ^ 1u, P.sleep_prod);
      const uint32_t a_s = a_base + (uint32_t)stage * P.a_stage_bytes + dst0;
      const bf16* xc = xt + c * P.kc;
      if (P.dbg & 1) {
      } else if (interior) {
#pragma unroll
        for (int j = 0; j < kIt; ++j)
          if (j < kIt - 1 || last_ok) cp_async16_full(a_s + (uint32_t)(j * kStep * 16), xc + src_off[j]);
      } else {
        int y0v, x0v;        // opaque copies: keep the border arithmetic inside this branch (the compiler hoisted it into every tile)
        asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(y0v), "=r"(x0v) : "r"(y0), "r"(x0));
        const int ylo = y0v < 0 ? 1 : 0, xlo = x0v < 0 ? 1 : 0;
        const uint32_t ny = (uint32_t)((P.H - y0v < kHTH + 2 ? P.H - y0v : kHTH + 2) - ylo);
        const uint32_t nx = (uint32_t)((P.W - x0v < HHW ? P.W - x0v : HHW) - xlo);
        int hy = hy0, hx = hx0;
#pragma unroll
        for (int j = 0; j < kIt; ++j) {
          if (j < kIt - 1 || last_ok) {
            const bool v = (uint32_t)(hy - ylo) < ny && (uint32_t)(hx - xlo) < nx;
            cp_async16(a_s + (uint32_t)(j * kStep * 16), xc + (v ? src_off[j] : safe_off), v ? 16u : 0u);     // zero fill = the conv padding
          }
          hx += kStep % HHW; hy += kStep / HHW;
          if (hx >= HHW) { hx -= HHW; ++hy; }
        }
      }
      // the mbarrier tracks this thread's copies itself (arrive-on-completion, counted in the 128 expected arrivals): no
      // wait_group / fence in the producer — MEMBAR + FENCE.VIEW.ASYNC here waited for EVERY copy in flight, i.e. the
      // memory latency of each tile was fully exposed whatever the number of groups kept in flight
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8u * (uint32_t)stage) : "memory");
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    stage += skip;
    while (stage >= S) { stage -= S; phase ^= 1u; }
    for (int a = 0; a < PG; ++a)
      if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++img; } }
  }
  cp_async_wait_all();     // copies must have landed before the thread may exit
}


// Epilogue of one tile and one warp (32 pixels = 4 tile rows x 8 columns) through the TMA store: TMEM -> registers -> bias /
// LeakyReLU / bf16 -> swizzled staging rows -> one store per CW-channel block.  CW = channels per store box (16 / 32 / 64,
// staging row = 2 CW bytes with the matching 32 / 64 / 128-byte swizzle).
template <int CW>
__device__ __forceinline__ void halo_store_tile(const HaloParams& P, const CUtensorMap* mapY, uint32_t taddr, uint32_t wst, const float* bias,
                                                float slope, uint64_t* acc_empty_buf, int lane, int gx0, int gy0, int img) {
  constexpr uint32_t rb = (uint32_t)CW * 2u;                       // staging row bytes: 128 / 64 / 32
  constexpr uint32_t swz_mask = (uint32_t)(CW >> 3) - 1u;
  constexpr int ngrp = CW >> 4;                                    // 16-column register groups per block: 1, 2 or 4
  const uint32_t swz = (((uint32_t)lane * rb) >> 7) & swz_mask;
  const uint32_t rowa = wst + (uint32_t)lane * rb;
  for (int c0 = 0; c0 < P.Cout; c0 += CW) {
    uint32_t r[16 * ngrp];
#pragma unroll
    for (int gi = 0; gi < ngrp; ++gi) tmem_ld16_nowait(taddr + (uint32_t)(c0 + gi * 16), r + gi * 16);
    if (lane == 0) bulk_wait_read<0>();                        // this warp's previous store has read the staging rows
    __syncwarp();
    tmem_ld_wait();
    if (c0 + CW >= P.Cout) {                                   // accumulator drained: hand the TMEM buffer back NOW
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty_buf));
    }
#pragma unroll
    for (int gi = 0; gi < ngrp; ++gi) {
      tmem_ld_fence16(r + gi * 16);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cl = gi * 16 + h * 8;
        uint32_t packed[4];
        // the epilogue warps run alone on their SM sub-partitions and their instruction stream sets the tile period of the N <= 64 layers:
        // input gradients (no bias, no activation) only convert and pack, layers without activation skip the LeakyReLU pair
        if (P.bias == nullptr && P.act != RD_ACT_LRELU) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(r[cl + 2 * qq]), __uint_as_float(r[cl + 2 * qq + 1]));
            packed[qq] = *reinterpret_cast<uint32_t*>(&b2);
          }
        } else {
        const float4 b0 = *reinterpret_cast<const float4*>(&bias[c0 + cl]);
        const float4 b1 = *reinterpret_cast<const float4*>(&bias[c0 + cl + 4]);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float v0 = __uint_as_float(r[cl + 2 * qq]) + bb[2 * qq], v1 = __uint_as_float(r[cl + 2 * qq + 1]) + bb[2 * qq + 1];
          if (P.act == RD_ACT_LRELU) { v0 = fmaxf(v0, v0 * slope); v1 = fmaxf(v1, v1 * slope); }
          __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
          packed[qq] = *reinterpret_cast<uint32_t*>(&b2);
        }
        }
        const uint32_t chunk = (uint32_t)(cl >> 3);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((chunk ^ swz) << 4)), "r"(packed[0]),
                     "r"(packed[1]), "r"(packed[2]), "r"(packed[3]) : "memory");
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0 && !(P.dbg & 4)) {
      tma_store_4d(mapY, wst, c0, gx0, gy0, img);
      bulk_commit();
    }
  }
}

// SPADE epilogue of one tile and one warp (32 pixels): per 16-channel sub-block, gamma and beta columns from TMEM, z from global
// memory (32 contiguous bytes per pixel, requested two sub-blocks ahead — the first two before the accumulator wait), one float4 of
// per-(image, channel) constants {bias_gamma, bias_beta, invstd, -mean * invstd} from shared memory -> gamma and mix as 64-byte staging
// rows (32 channels, SWIZZLE_64B) -> two TMA stores per 32-channel block, or direct 16-byte stores when the staging buffers did not fit
// next to the resident weights (STAGED = false).  The epilogue warps run alone on their SM sub-partitions, so the instruction count per
// channel sets the tile period: 1 LDS.128 + 5 FP32 operations + unpack / pack per channel.
template <bool STAGED>
__device__ __forceinline__ void halo_store_spade(const HaloParams& P, const CUtensorMap* mapG, const CUtensorMap* mapM, uint32_t taddr,
                                                 uint32_t wst, const float4* cst, uint32_t accf_bar, uint32_t accf_par,
                                                 uint64_t* acc_empty_buf, int lane, int gx0, int gy0, int img, int64_t pix, bool pvalid) {
  const int C = P.spade;
  const uint32_t swz = ((uint32_t)lane >> 1) & 3u;                  // 64-byte rows: 16-byte chunk index ^ ((row * 64) >> 7) & 3
  const uint32_t rowa = wst + (uint32_t)lane * 64u;
  const uint4* zrow = reinterpret_cast<const uint4*>(P.z + pix * C);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  uint4 za0 = zero4, za1 = zero4, zb0 = zero4, zb1 = zero4;         // sub-block k (za) and k + 1 (zb)
  if (pvalid) { za0 = __ldg(zrow); za1 = __ldg(zrow + 1); zb0 = __ldg(zrow + 2); zb1 = __ldg(zrow + 3); }
  mbar_wait_sleep(accf_bar, accf_par, P.sleep_epi);
  tc_fence_after();
  for (int c16 = 0; c16 < C; c16 += 16) {
    uint32_t rg[16], rb[16];
    tmem_ld16_nowait(taddr + (uint32_t)c16, rg);
    tmem_ld16_nowait(taddr + (uint32_t)(C + c16), rb);
    if (STAGED && (c16 & 16) == 0) {
      if (lane == 0) bulk_wait_read<0>();                      // this warp's previous stores have read the staging rows
      __syncwarp();
    }
    tmem_ld_wait();
    const bool last = c16 + 16 >= C;
    if (last) {                                                // accumulator drained: hand the TMEM buffer back NOW
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty_buf));
    }
    tmem_ld_fence16(rg); tmem_ld_fence16(rb);
    uint4 zn0 = zero4, zn1 = zero4;
    if (c16 + 32 < C && pvalid) { zn0 = __ldg(zrow + (c16 >> 3) + 4); zn1 = __ldg(zrow + (c16 >> 3) + 5); }      // z of sub-block k + 2
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cl = c16 + h * 8;
      const uint4 zq = h ? za1 : za0;
      const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
      uint32_t pg[4], pm[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float4 k0 = cst[cl + 2 * qq], k1 = cst[cl + 2 * qq + 1];        // {bias_gamma, bias_beta, invstd, -mean * invstd}
        const float g0 = __uint_as_float(rg[h * 8 + 2 * qq]) + k0.x, g1 = __uint_as_float(rg[h * 8 + 2 * qq + 1]) + k1.x;
        const float t0 = __uint_as_float(rb[h * 8 + 2 * qq]) + k0.y, t1 = __uint_as_float(rb[h * 8 + 2 * qq + 1]) + k1.y;
        const float zh0 = fmaf(__uint_as_float(zw[qq] << 16), k0.z, k0.w), zh1 = fmaf(__uint_as_float(zw[qq] & 0xffff0000u), k1.z, k1.w);
        const float m0 = fmaf(zh0, g0, zh0) + t0, m1 = fmaf(zh1, g1, zh1) + t1;       // zhat * (1 + gamma) + beta
        __nv_bfloat162 g2 = __floats2bfloat162_rn(g0, g1), m2 = __floats2bfloat162_rn(m0, m1);
        pg[qq] = *reinterpret_cast<uint32_t*>(&g2);
        pm[qq] = *reinterpret_cast<uint32_t*>(&m2);
      }
      if (STAGED) {
        const uint32_t off = (((uint32_t)(((c16 & 16) >> 3) + h) ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + off), "r"(pg[0]), "r"(pg[1]), "r"(pg[2]), "r"(pg[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + 2048u + off), "r"(pm[0]), "r"(pm[1]), "r"(pm[2]), "r"(pm[3]) : "memory");
      } else if (pvalid) {
        *reinterpret_cast<uint4*>(P.gamma + pix * C + cl) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
        *reinterpret_cast<uint4*>(P.mix + pix * C + cl) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
      }
    }
    if (STAGED && (c16 & 16)) {                                // a 32-channel block is staged: one store box per output tensor
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(P.dbg & 4)) {
        tma_store_4d(mapG, wst, c16 - 16, gx0, gy0, img);
        tma_store_4d(mapM, wst + 2048u, c16 - 16, gx0, gy0, img);
        bulk_commit();
      }
    }
    za0 = zb0; za1 = zb1; zb0 = zn0; zb1 = zn1;
  }
}

template <int NT, bool SPADE>
__global__ void __launch_bounds__(kHThreads, 1)
k_conv_halo(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX,
            const __grid_constant__ CUtensorMap mapM, const HaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kHMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kHMaxStages];
  __shared__ __align__(8) uint64_t acc_full[kHMaxAcc];
  __shared__ __align__(8) uint64_t acc_empty[kHMaxAcc];
  __shared__ __align__(8) uint64_t w_full, w_free;
  __shared__ uint32_t tmem_base_s;
  // one copy per epilogue group (they may be on different weight groups / images): the bias row, or (SPADE) per channel
  // {bias_gamma, bias_beta, invstd, -mean * invstd} of the group's current image
  __shared__ __align__(16) float bias_raw[SPADE ? 2 * 128 * 4 : 2 * 256];
  float (*bias_s)[256] = reinterpret_cast<float (*)[256]>(bias_raw);
  float4 (*cst_s)[128] = reinterpret_cast<float4 (*)[128]>(bias_raw);

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
      if (SPADE && P.stg_bufs) tma_prefetch_desc(&mapM);
      for (int s = 0; s < S; ++s) {
        mbar_init(smem_u32(&full_bar[s]), P.use_tma ? 1 : 128);
        mbar_init(smem_u32(&empty_bar[s]), (NT == 2 && P.dual) ? 2 : 1);
      }
      for (int b = 0; b < kHMaxAcc; ++b) {
        mbar_init(smem_u32(&acc_full[b]), 1);
        mbar_init(smem_u32(&acc_empty[b]), 4);
      }
      mbar_init(smem_u32(&w_full), 1);
      mbar_init(smem_u32(&w_free), P.dual ? 2 : 1);
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

  if (((warp >= 4 && warp < 8) || warp >= 13) && !(P.dual && warp == 5)) {
    // ------------------------------------------------------------------ halo producers: two groups of 128 threads (warps 4-7
    // and 13-16) that fill alternate tiles — one warp per SM sub-partition cannot issue a tile's copies in the time its MMAs take
    const int nb = P.kc >> 3;                       // 16-byte channel blocks per stage (2, 4 or 8)
    if (P.use_tma) {
      // one thread, one 5-D box per stage: (8 ch, halo columns, halo rows, channel blocks, image) lands as [plane][row][pixel][16 B],
      // out-of-image pixels zero-filled by the TMA unit (= the conv padding)
      if (tid == 128) {
        tma_prefetch_desc(&mapX);
        constexpr int TWn = HaloGeom<NT>::TW;
        const int stx = P.tiles_x / NT, sty = P.tiles_per_img / P.tiles_x, st_per_img = P.tiles_per_img / NT;
        const int s0 = t_begin / NT, s1 = t_end / NT;
        int img = s0 / st_per_img;
        int ty = (s0 - img * st_per_img) / stx;
        int tx = s0 - img * st_per_img - ty * stx;
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = s0; t < s1; ++t) {
          for (int c = 0; c < P.chunks; ++c) {
            mbar_wait_sleep(empty0 + 8u * (uint32_t)stage, phase ^ 1u, P.sleep_prod);
            if (P.dbg & 1) mbar_arrive(full0 + 8u * (uint32_t)stage);
            else {
            mbar_arrive_expect_tx(full0 + 8u * (uint32_t)stage, P.a_stage_bytes);
            tma_load_5d(a_base + (uint32_t)stage * P.a_stage_bytes, &mapX, 0, tx * TWn - 1, ty * kHTH - 1, c * nb, img, full0 + 8u * (uint32_t)stage);
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
          if (++tx == stx) { tx = 0; if (++ty == sty) { ty = 0; ++img; } }
        }
      }
    } else {
    const int pgrp = warp >= 13 ? 1 : 0;
    const int ptid = pgrp ? tid - 416 : tid - 128;
    if (nb == 8) halo_producer<8, NT>(P, a_base, full_bar, empty_bar, ptid, pgrp, t_begin, t_end);
    else if (nb == 4) halo_producer<4, NT>(P, a_base, full_bar, empty_bar, ptid, pgrp, t_begin, t_end);
    else halo_producer<2, NT>(P, a_base, full_bar, empty_bar, ptid, pgrp, t_begin, t_end);
    }
  } else if (warp == 8 || (P.dual && warp == 5)) {
    const int iss = warp == 8 ? 0 : 1;
    // ------------------------------------------------------------------ weights + MMA issuer(s)
    // The whole warp runs this (warp-uniform control flow and values); one elected lane issues the TMA / tcgen05
    // instructions.  Under `if (lane == 0)` the compiler treats every descriptor as per-thread data and wraps each
    // tcgen05.mma in an R2UR / ELECT / BRA.U.ANY serialisation loop — the single issuing thread then cannot keep the
    // tensor core fed (ncu: tensor pipe 48 % active, issuer 57 % busy executing, profiles/r01_ncu_conv_halo.txt).
    const int ksteps = P.kc >> 4;
#define RD_HALO_MMA(KS, FB, CI, SG) halo_mma<KS, FB, CI, SG, NT>(P, smem_base, a_base, tmem_base, full_bar, empty_bar, acc_full, acc_empty, &w_full, &w_free, &mapB, t_begin, t_end, iss)
    if (NT == 2) {                                         // two tiles per stage: planned only for the compile-time shapes (Cin 16 / 32 / 64, one chunk)
      if (P.Cin == 16) { if (P.sign > 0) RD_HALO_MMA(1, true, 16, 1); else RD_HALO_MMA(1, true, 16, -1); }
      else if (P.Cin == 32) { if (P.sign > 0) RD_HALO_MMA(2, true, 32, 1); else RD_HALO_MMA(2, true, 32, -1); }
      else { if (P.sign > 0) RD_HALO_MMA(4, true, 64, 1); else RD_HALO_MMA(4, true, 64, -1); }
    } else if (P.dbg & 2) {                                // timing experiments: runtime-offset variant (it honours the no-MMA switch)
      if (ksteps == 4) RD_HALO_MMA(4, true, 0, 0); else if (ksteps == 2) RD_HALO_MMA(2, true, 0, 0); else RD_HALO_MMA(1, true, 0, 0);
    } else if (P.chunks == 1 && P.Cin == 16) { if (P.sign > 0) RD_HALO_MMA(1, true, 16, 1); else RD_HALO_MMA(1, true, 16, -1); }
    else if (P.chunks == 1 && P.Cin == 32) { if (P.sign > 0) RD_HALO_MMA(2, true, 32, 1); else RD_HALO_MMA(2, true, 32, -1); }
    else if (P.chunks == 1 && P.Cin == 64) { if (P.sign > 0) RD_HALO_MMA(4, true, 64, 1); else RD_HALO_MMA(4, true, 64, -1); }
    else if (P.chunks == 2 && P.Cin == 128) { if (P.sign > 0) RD_HALO_MMA(4, true, 128, 1); else RD_HALO_MMA(4, true, 128, -1); }
    else if (P.chunks == 1 || P.kc == 64) {
      if (ksteps == 4) RD_HALO_MMA(4, true, 0, 0); else if (ksteps == 2) RD_HALO_MMA(2, true, 0, 0); else RD_HALO_MMA(1, true, 0, 0);
    } else {
      if (ksteps == 2) RD_HALO_MMA(2, false, 0, 0); else RD_HALO_MMA(1, false, 0, 0);      // several chunks narrower than a weight box (Cin = 48, 96 ...)
    }
#undef RD_HALO_MMA
  } else {
    // ------------------------------------------------------------------ epilogue: group 0 = warps 0-3 (even tiles, TMEM
    // buffer 0), group 1 = warps 9-12 (odd tiles, buffer 1); warp & 3 = the TMEM lane quarter the warp may read
    const int grp = warp >= 9 ? 1 : 0;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tyl = row >> 3, txl = row & 7;
    const uint32_t stg_base = smem_base + P.stg_off;
    const float slope = P.act == RD_ACT_LRELU ? P.slope : 1.f;      // max(v, slope * v) == LeakyReLU for 0 < slope <= 1
    int bias_g = -1, last_img = -1;
    const uint32_t accf0 = smem_u32(acc_full);
    // tile coordinates advance by two tiles per iteration without divisions (the weight group / bias row is recomputed
    // only when the image changes)
    const int tiles_y = P.tiles_per_img / P.tiles_x;
    int img, ty, tx;
    {
      const int t0 = t_begin + grp;
      img = t0 / P.tiles_per_img;
      ty = (t0 - img * P.tiles_per_img) / P.tiles_x;
      tx = t0 - img * P.tiles_per_img - ty * P.tiles_x;
    }
    for (int t = t_begin + grp; t < t_end; t += 2) {
      const int it = t - t_begin;
      if (img != last_img) {   // (re)load this epilogue group's bias row when the tile's weight group changes (named barrier 1 + grp, 128 threads)
        last_img = img;
        const int tg0 = img / P.ipg;
        const int tg = P.bias_gpr ? tg0 / P.bias_gpr : 0;            // bias row of the tile's weight group
        if (tg != bias_g || SPADE) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          const float* src = P.bias ? P.bias + (size_t)tg * P.Cout : nullptr;
          if (SPADE) {                                               // bias pair + the image's InstanceNorm statistics
            const int i = (warp & 3) * 32 + lane;
            if (i < P.spade) {
              const float is = P.invstd[(size_t)img * P.spade + i];
              cst_s[grp][i] = make_float4(src ? src[i] : 0.f, src ? src[P.spade + i] : 0.f, is, -P.mean[(size_t)img * P.spade + i] * is);
            }
          } else if (tg != bias_g) {
            for (int i = (warp & 3) * 32 + lane; i < 256; i += 128) bias_s[grp][i] = (src != nullptr && i < P.Cout) ? src[i] : 0.f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          bias_g = tg;
        }
      }
      const int buf = it & (P.n_acc - 1);
      const int cimg = img, cty = ty, ctx = tx;                      // this tile; then step two tiles ahead
#pragma unroll
      for (int adv = 0; adv < 2; ++adv)
        if (++tx == P.tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++img; } }
      const int gy = cty * kHTH + tyl, gx = ctx * kHTW + txl;
      const bool pvalid = gy < P.H && gx < P.W;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * (uint32_t)P.n_tile;
      if (SPADE) {
        const int64_t pix = pvalid ? (((int64_t)cimg * P.H + gy) * P.W + gx) : 0;
        // z of this group's NEXT tile -> L2 now: the epilogue requests a tile's z just before its accumulator wait and needs it right
        // after, so the whole DRAM latency sat on every tile (ncu source view: 18 % of the kernel's stall samples on the first use of z,
        // 6 % at the accumulator barrier); one tile period later the loads hit the L2
        if (t + 2 < t_end && P.zpf) {
          const int ngy = ty * kHTH + tyl, ngx = tx * kHTW + txl;
          if (ngy < P.H && ngx < P.W) {
            const bf16* zn = P.z + (((int64_t)img * P.H + ngy) * P.W + ngx) * P.spade;
            if (P.zpf == 2) {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(zn));
              if (P.spade > 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(zn + 32));
            } else {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(zn));
              if (P.spade > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(zn + 32));
            }
          }
        }
        const uint32_t abar = accf0 + 8u * (uint32_t)buf, apar = ((uint32_t)it >> P.acc_shift) & 1u;
        if (P.stg_bufs) {
          const uint32_t wst = stg_base + (uint32_t)(grp * 4 + q) * 4096u;
          halo_store_spade<true>(P, &mapY, &mapM, taddr, wst, cst_s[grp], abar, apar, &acc_empty[buf], lane,
                                 ctx * kHTW, cty * kHTH + q * 4, cimg, pix, pvalid);
        } else {
          halo_store_spade<false>(P, &mapY, &mapM, taddr, 0u, cst_s[grp], abar, apar, &acc_empty[buf], lane,
                                  ctx * kHTW, cty * kHTH + q * 4, cimg, pix, pvalid);
        }
        continue;
      }
      bf16* yrow = P.y + (pvalid ? (((int64_t)cimg * P.H + gy) * P.W + gx) : 0) * P.Cout;
      mbar_wait_sleep(accf0 + 8u * (uint32_t)buf, ((uint32_t)it >> P.acc_shift) & 1u, P.sleep_epi);
      tc_fence_after();
      if (P.stg_bufs) {
        // TMEM -> registers -> swizzled staging rows -> one TMA store per warp and 64-channel block.  (Lane-per-pixel
        // st.global touches 32 different 128-byte lines per instruction.)  All tcgen05.ld of a block are issued before
        // one wait, so their latencies overlap.
        const uint32_t wst = stg_base + (uint32_t)((P.stg_bufs == 2 ? grp : 0) * 4 + q) * 64u * (uint32_t)P.store_cw;
        if (P.store_cw == 64) halo_store_tile<64>(P, &mapY, taddr, wst, bias_s[grp], slope, &acc_empty[buf], lane, ctx * kHTW, cty * kHTH + q * 4, cimg);
        else if (P.store_cw == 32) halo_store_tile<32>(P, &mapY, taddr, wst, bias_s[grp], slope, &acc_empty[buf], lane, ctx * kHTW, cty * kHTH + q * 4, cimg);
        else halo_store_tile<16>(P, &mapY, taddr, wst, bias_s[grp], slope, &acc_empty[buf], lane, ctx * kHTW, cty * kHTH + q * 4, cimg);
        continue;
      }
      if (P.narrow && ctx * kHTW + kHTW <= P.W) {
        // Narrow outputs (the decoder's 7 image channels, the 4 anatomy logits): a pixel row is Cout * 2 bytes, so lane-per-pixel
        // 2-byte stores touch ~16 sectors per instruction and channel.  The 8 pixels of a tile row are 16 * Cout CONTIGUOUS, 16-byte
        // aligned bytes (W is a multiple of 8): stage the warp's 32 pixels (4 tile rows) and write them as 4 * Cout 16-byte stores.
        uint32_t r[16];
        tmem_ld16(taddr, r);
        uint8_t* stn = smem_raw + (smem_base - smem_u32(smem_raw)) + P.stg_off + (uint32_t)(grp * 4 + q) * 512u;
#pragma unroll
        for (int co = 0; co < 7; ++co) {
          if (co < P.Cout) {
            float v0 = __uint_as_float(r[co]) + bias_s[grp][co];
            if (P.act == RD_ACT_LRELU) v0 = v0 > 0.f ? v0 : v0 * P.slope;
            *reinterpret_cast<bf16*>(stn + (lane * P.Cout + co) * 2) = __float2bfloat16_rn(v0);
          }
        }
        __syncwarp();
        const int row0 = cty * kHTH + q * 4;
        for (int j = lane; j < 4 * P.Cout; j += 32) {
          const int sg = j / P.Cout, c16 = j - sg * P.Cout;
          if (row0 + sg < P.H) {
            const uint4 v = *reinterpret_cast<const uint4*>(stn + j * 16);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(P.y) + ((((int64_t)cimg * P.H + row0 + sg) * P.W + ctx * kHTW) * P.Cout) * 2 + c16 * 16) = v;
          }
        }
        __syncwarp();                       // the staging rows are rewritten by this warp's next tile
      } else
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
  int cin, cout, n_tile, kc, chunks, w_boxes, stages, stg_bufs, store_cw, nt;
  int narrow;                      // Cout < 8: 8 x 512 B of staging rows at stg_off for the coalesced narrow-output epilogue
  uint32_t w_box_bytes, w_bytes, a_stage_bytes, stg_off;
};

bool halo_plan_nt(const rd_conv_desc* d, int mode, int nt, HaloPlan& pl) {
  if (d->dtype != RD_BF16) return false;
  pl.nt = nt;
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
  pl.a_stage_bytes = (uint32_t)(pl.kc / 8) * (nt == 2 ? HaloGeom<2>::Plane : HaloGeom<1>::Plane);     // multiple of 16 B
  // TMA-store epilogue: 16 / 32 / 64 output channels, or a multiple of 64 (one store box per 64-channel block)
  pl.store_cw = pl.cout >= 64 ? 64 : pl.cout;
  const bool can_store = (pl.cout % 64 == 0) || pl.cout == 32 || pl.cout == 16;
  const uint32_t avail = kHaloSmemMax - 1024u;
  pl.narrow = (pl.cout < 8 && d->w % 8 == 0) ? 1 : 0;        // 32 pixels x Cout x 2 bytes <= 448 of a warp's 512 staging bytes
  for (int bufs = can_store ? 2 : 0; bufs >= 0; bufs -= 2) {     // one staging buffer per epilogue group, or direct stores
    uint32_t stg = (uint32_t)bufs * 128u * (uint32_t)pl.store_cw * 2u + (bufs ? 1024u : 0u);     // + alignment slack
    if (pl.narrow) stg = 8u * 512u + 1024u;
    if (pl.w_bytes + stg + 2u * pl.a_stage_bytes > avail) continue;
    int st = (int)((avail - pl.w_bytes - stg) / pl.a_stage_bytes);
    pl.stages = st > kHMaxStages ? kHMaxStages : st;
    pl.stg_bufs = bufs;
    pl.stg_off = (pl.w_bytes + (uint32_t)pl.stages * pl.a_stage_bytes + 1023u) & ~1023u;
    return true;
  }
  return false;
}

// Two tiles per stage when the image is an even number of tiles wide, four accumulators fit in TMEM (each tile of the pair has
// its own, and the next pair must start while this one drains) and the ring still holds four of the larger stages.
bool halo_plan(const rd_conv_desc* d, int mode, HaloPlan& pl) {
  static const char* e_nt = getenv("RD_B200_HALO_NT");          // tuning knob: 1 = always one tile per stage
  const int max_nt = e_nt ? atoi(e_nt) : 2;
  if (max_nt >= 2 && rd_div_up(d->w, kHTW) % 2 == 0) {
    HaloPlan p2;
    if (halo_plan_nt(d, mode, 2, p2) && p2.n_tile <= 128 && p2.chunks == 1 && p2.stages >= 4 &&
        (p2.cin == 16 || p2.cin == 32 || p2.cin == 64)) { pl = p2; return true; }
  }
  return halo_plan_nt(d, mode, 1, pl);
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

namespace {
struct HaloSpadeArgs { const void* z; const float* mean; const float* invstd; void* gamma; void* mix; };
int halo_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                const HaloSpadeArgs* sp, cudaStream_t st);
}
int rd_conv_halo_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                        cudaStream_t st) {
  return halo_launch(ctx, d, mode, x, w, bias, y, nullptr, st);
}
// gamma|beta convolution (Cout = 2C) with the SPADE modulation in its epilogue; see HaloParams::spade
int rd_conv_halo_spade_supported(const rd_conv_desc* d, int sm_count) {
  HaloPlan pl;
  if (d->cout % 64 || d->cout > 256 || d->act != RD_ACT_NONE) return 0;       // C = Cout / 2 in 32-channel blocks, <= 128
  if (!rd_conv_halo_supported(d, 0, sm_count, d->algo == RD_ALGO_HALO) || !halo_plan(d, 0, pl)) return 0;
  // the SPADE variants hold 2 KB more static shared memory (the per-channel constants): their dynamic part must stay within 222 KB
  const size_t smem = (pl.stg_bufs ? (size_t)pl.stg_off + (size_t)pl.stg_bufs * 128 * pl.store_cw * 2 : (size_t)pl.w_bytes + (size_t)pl.stages * pl.a_stage_bytes) + 1024;
  return (pl.store_cw == 64 && smem <= 222u * 1024u) ? 1 : 0;
}
int rd_conv_halo_spade_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* w, const float* bias, const void* z,
                              const float* mean, const float* invstd, void* gamma, void* mix, cudaStream_t st) {
  if (!rd_conv_halo_spade_supported(d, ctx->sm_count)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv_halo_spade: shape not supported");
  HaloSpadeArgs sp{z, mean, invstd, gamma, mix};
  return halo_launch(ctx, d, 0, x, w, bias, gamma, &sp, st);
}
namespace {
int halo_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                const HaloSpadeArgs* sp, cudaStream_t st) {
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
  if (pl.nt == 2) P.tiles_per_cta += P.tiles_per_cta & 1;         // a stage (tile pair) never straddles two CTAs
  grid = rd_div_up(P.total_tiles, P.tiles_per_cta);
  P.n_tile = pl.n_tile; P.kc = pl.kc; P.chunks = pl.chunks; P.w_boxes = pl.w_boxes;
  P.w_box_bytes = pl.w_box_bytes; P.w_bytes = pl.w_bytes;
  P.w_tx_bytes = pl.w_bytes;
  P.a_stage_bytes = pl.a_stage_bytes;
  P.stages = pl.stages;
  {
    static const char* e_st = getenv("RD_B200_HALO_STAGES");      // tuning knob: cap on the halo stage ring
    if (e_st) { int v = atoi(e_st); if (v >= 2 && v < P.stages) P.stages = v; }
  }
  // One producer group: a second one (warps 13-16, alternate tiles) is supported by the kernel but 17 warps cap the kernel at 96
  // registers per thread (warp slots are allocated in fours) and the spills cost more than the group gains (measured).
  P.pgroups = 1; P.lag = 0;
  { const char* e_du = getenv("RD_B200_HALO_DUAL"); P.dual = (e_du && atoi(e_du) == 0) ? 0 : 1; }      // A/B switch, read per call
  P.sleep_epi = 200; P.sleep_mma = 40; P.sleep_prod = 80;
  { const char* e_dbg = getenv("RD_B200_HALO_DEBUG"); P.dbg = e_dbg ? atoi(e_dbg) : 0; }      // per call: tools/ablate_halo.py sweeps it
  {
    static const char* e_sl = getenv("RD_B200_HALO_SLEEP");       // tuning knob: "epi,mma,prod" in ns
    if (e_sl) { unsigned a = 0, b = 0, c = 0; if (sscanf(e_sl, "%u,%u,%u", &a, &b, &c) == 3) { P.sleep_epi = a; P.sleep_mma = b; P.sleep_prod = c; } }
  }
  P.stg_off = pl.stg_off; P.stg_bufs = pl.stg_bufs; P.store_cw = pl.store_cw;
  { const char* e_zp = getenv("RD_B200_HALO_ZPF"); P.zpf = e_zp ? atoi(e_zp) : 1; }
  { const char* e_nw = getenv("RD_B200_HALO_NARROW"); P.narrow = (pl.narrow && !(e_nw && atoi(e_nw) == 0)) ? 1 : 0; }      // A/B switch
  P.n_acc = 2; P.acc_shift = 1;
  while (P.n_acc < kHMaxAcc && 2 * P.n_acc * pl.n_tile <= 512) { P.n_acc *= 2; ++P.acc_shift; }
  uint32_t cols = 32;
  while (cols < (uint32_t)(P.n_acc * pl.n_tile)) cols <<= 1;
  P.tmem_cols = cols;
  P.act = mode == 0 ? d->act : RD_ACT_NONE;
  P.slope = d->act_slope;
  P.bias_gpr = (mode == 0 && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;
  P.spade = 0; P.z = nullptr; P.mean = P.invstd = nullptr; P.gamma = P.mix = nullptr;
  if (sp) {
    P.spade = pl.cout / 2;
    P.z = (const bf16*)sp->z; P.mean = sp->mean; P.invstd = sp->invstd; P.gamma = (bf16*)sp->gamma; P.mix = (bf16*)sp->mix;
  }

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
  alignas(64) CUtensorMap mapY, mapM;
  memset(&mapY, 0, sizeof(mapY));
  memset(&mapM, 0, sizeof(mapM));
  if (pl.stg_bufs && sp) {          // two [N, H, W, C] outputs, 32-channel store boxes (64-byte staging rows)
    const int C = P.spade;
    void* outs[2] = {sp->gamma, sp->mix};
    CUtensorMap* maps[2] = {&mapY, &mapM};
    for (int k = 0; k < 2; ++k) {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)d->w * C * 2, (cuuint64_t)d->h * d->w * C * 2};
      cuuint32_t box[4] = {32u, (cuuint32_t)kHTW, 4u, 1u};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(maps[k], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, outs[k], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(halo spade output) failed: %d", (int)r);
    }
  } else if (pl.stg_bufs) {
    CUtensorMapSwizzle swy = pl.store_cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (pl.store_cw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[4] = {(cuuint64_t)pl.cout, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)pl.cout * 2, (cuuint64_t)d->w * pl.cout * 2, (cuuint64_t)d->h * d->w * pl.cout * 2};
    cuuint32_t box[4] = {(cuuint32_t)pl.store_cw, (cuuint32_t)kHTW, 4u, 1u};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&mapY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swy,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(halo output) failed: %d", (int)r);
  }
  alignas(64) CUtensorMap mapX;
  memset(&mapX, 0, sizeof(mapX));
  {
    static const char* e_tma = getenv("RD_B200_HALO_TMA");
    P.use_tma = (e_tma && atoi(e_tma) == 0) ? 0 : 1;        // default since it measured 5-17 % faster on every halo layer; 0 = cp.async producers
    if (!P.use_tma) P.dual = 0;                             // warp 5 is a cp.async producer in the fallback path
    { const char* e_d1 = getenv("RD_B200_HALO_DUAL1"); if (pl.nt != 2 && e_d1 && atoi(e_d1) == 0) P.dual = 0; }      // A/B switch for the one-tile stages
    if (P.dual && pl.nt != 2) {
      // Alternating tiles: an issuer must see EVERY fill of the stages it waits on (an mbarrier wait only knows the phase parity — a
      // wait for fill k of a stage whose fill k - 1 went to the other issuer can pass on the wrong phase).  With the ring a multiple of
      // two tiles' stages each stage always belongs to the same issuer; rings too short for that keep the single issuer.
      const int m = 2 * P.chunks;
      if (P.stages >= m) P.stages -= P.stages % m; else P.dual = 0;
    }
    if (P.use_tma) {
      const int hhw = pl.nt == 2 ? HaloGeom<2>::HHW : HaloGeom<1>::HHW;
      cuuint64_t dims[5] = {8u, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)(pl.cin / 8), (cuuint64_t)d->n};
      cuuint64_t strides[4] = {(cuuint64_t)pl.cin * 2, (cuuint64_t)d->w * pl.cin * 2, 16u, (cuuint64_t)d->h * d->w * pl.cin * 2};
      cuuint32_t box[5] = {8u, (cuuint32_t)hhw, (cuuint32_t)(kHTH + 2), (cuuint32_t)(pl.kc / 8), 1u};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(halo input) failed: %d", (int)r);
    }
  }
  size_t smem = (pl.stg_bufs ? (size_t)pl.stg_off + (size_t)pl.stg_bufs * 128 * pl.store_cw * 2 : (size_t)pl.w_bytes + (size_t)pl.stages * pl.a_stage_bytes) + 1024;
  if (pl.narrow) smem = (size_t)pl.stg_off + 8 * 512 + 1024;
  if (!g_halo_attr_set) {
    RD_CUDA(ctx, (cudaFuncSetAttribute(k_conv_halo<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)));
    RD_CUDA(ctx, (cudaFuncSetAttribute(k_conv_halo<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)));
    RD_CUDA(ctx, (cudaFuncSetAttribute(k_conv_halo<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024)));
    RD_CUDA(ctx, (cudaFuncSetAttribute(k_conv_halo<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024)));
    g_halo_attr_set = true;
  }
  if (sp) {
    if (pl.nt == 2) k_conv_halo<2, true><<<grid, kHThreads, smem, st>>>(mapB, mapY, mapX, mapM, P);
    else k_conv_halo<1, true><<<grid, kHThreads, smem, st>>>(mapB, mapY, mapX, mapM, P);
  } else if (pl.nt == 2) k_conv_halo<2, false><<<grid, kHThreads, smem, st>>>(mapB, mapY, mapX, mapM, P);
  else k_conv_halo<1, false><<<grid, kHThreads, smem, st>>>(mapB, mapY, mapX, mapM, P);
  RD_CHECK_LAUNCH(ctx, mode == 0 ? "conv_halo_fwd" : "conv_halo_dgrad");
  return RD_OK;
}
}  // namespace

// =====================================================================================================
// Halo-tile wgrad for the same layers (3x3 stride-1, Cin in {16, 32, 64}):
//     dK[g][co][(kh, kw)][ci] = sum_p dY[p, co] * X[p + (kh-1, kw-1), ci]
// k_wgrad_tma loads one shifted X box per tap through L2 (9x the tile) and is L2->SM bound for these layers (2x above
// its MMA floor).  Here the X tile is loaded ONCE with its halo in the layout [halo row][channel block][halo column][8 ch]
// and dY in [row][channel block][column][8 ch].  Both are canonical NO-SWIZZLE MN-major operands (core matrix = 8 pixels
// x 16 bytes): 8-pixel K groups = one tile row (LBO = one halo / tile row), 8-channel MN blocks SBO apart.  Because the
// halo ROW stride is exactly (channel blocks) x SBO, the blocks of consecutive kh taps continue the same arithmetic
// progression: ONE M = 128 instruction covers 16 / nb kh-slots x Cin channels (slots >= 3 read rows past the tap range:
// garbage accumulator rows, never stored), kw is a +16-byte shift of the start address.  Per 128-pixel tile:
// 8 K-steps x 3 kw x ceil(3 nb / 16) MMAs of M128 x N=Cout.  Accumulators stay in TMEM over the CTA's whole tile range
// (one weight group per CTA), the epilogue adds them into dK with red.global.add.f32; the bias gradient is summed from
// the dY tiles in shared memory by the four otherwise idle epilogue warps.
namespace {

constexpr int kWHThreads = 288;       // warps 0-3: bias sums + epilogue, 4-7: producers, 8: MMA issue / TMEM
constexpr int kWHMaxStages = 6;

struct WgHaloParams {
  const bf16* x; const bf16* dy; float* dK; float* dbias;
  int H, W, Cin, Cout;
  int nb, nbo;                    // 8-channel blocks of X / dY
  int mt;                         // M tiles: ceil(3 * nb / 16)
  int ipg, groups, ctas_per_group;
  int n_split, cout_cta;          // Cout split over n_split CTAs (TMEM holds 3 * mt * cout_cta <= 512 accumulator columns)
  int tiles_x, tiles_per_img, tiles_pg;
  uint32_t x_bytes, dy_bytes, stage_bytes;
  int stages, lag;
  uint32_t tmem_cols;
  int dbias_gpr;
  int use_tma;                    // X halo tile and dY tile by two 5-D TMA boxes per stage (default) instead of the cp.async producers
  int dbg;                        // timing experiments only (RD_B200_WGH_DEBUG): 1 = no TMA loads, 2 = no MMAs, 4 = no bias sums, 8 = dY box only, 16 = X box only
  int dual;                       // two MMA-issuing warps (8: even tiles, 5: odd tiles) with an accumulator set each (see halo_mma for why)
  uint32_t set_cols;              // TMEM columns of one accumulator set: 3 * mt * cout_cta
  // dY tile as MN-major SWIZZLED rows [pixel][bo channels] (bo = min(64, cout_cta): 128- / 64- / 32-byte rows, one 4-D TMA box per 64-channel
  // block) instead of the no-swizzle [row][block][column][16 B] planes: the TMA unit delivers a tile in 16-byte pieces at ~1.1 cycles per
  // piece (role ablation: sp6 gamma|beta wgrad took 0.40 of its 0.45 ms with the MMAs switched off, dY alone 0.24 ms for 1 024 pieces per
  // tile); row-contiguous boxes are 128 (64 / 32-byte: 128) pieces per tile
  int dy_sw;                      // 1: swizzled rows (TMA path only)
  int bo, dy_blocks;              // channels per dY box, boxes per stage
  uint32_t dy_row_bytes, dy_blk_bytes;
};

// MN-major NO-SWIZZLE descriptor: SBO = byte offset between 8-element MN blocks, LBO = between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return lo | (hi << 32);
}
// MN-major SWIZZLED descriptor (rows of 128 / 64 / 32 bytes = 64 / 32 / 16 channels of one pixel): SBO = 8 rows, LBO = next channel block
__device__ __forceinline__ uint64_t make_desc_mn_sw(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
  uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
  return lo | (hi << 32);
}
__device__ __forceinline__ uint32_t make_idesc_mn2(int m, int n) {   // both operands MN-major (bits 15, 16)
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// all MMAs of one 128-pixel tile: 8 K-steps (tile rows 2j, 2j+1) x 3 kw x MT M-tiles, fully unrolled
template <int MT>
__device__ __forceinline__ void issue_wgrad_tile(uint32_t tmem_base, uint64_t xa0, uint64_t da0, uint32_t xrow16, uint32_t drow16,
                                                 uint32_t cout, uint32_t idesc, bool first) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t xa = xa0 + (uint64_t)((uint32_t)j * xrow16);
    const uint64_t da = da0 + (uint64_t)((uint32_t)j * drow16);
    const uint32_t acc = (uint32_t)(!(first && j == 0));
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        umma_bf16(tmem_base + (uint32_t)(kw * MT + mt) * cout, xa + (uint64_t)(kw + mt * 160), da, idesc, acc);
  }
}

__global__ void __launch_bounds__(kWHThreads, 1) k_wgrad_halo(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapD,
                                                              const WgHaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWHMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWHMaxStages];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_red[128 * 8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzled dY rows: 1 KB aligned atoms
  const int S = P.stages;
  const int gs = blockIdx.x / P.ctas_per_group;            // (group, Cout split) pair
  const int g = gs / P.n_split;
  const int co0 = (gs - g * P.n_split) * P.cout_cta;       // first output channel of this CTA
  const int cg = blockIdx.x - gs * P.ctas_per_group;
  const int per = (P.tiles_pg + P.ctas_per_group - 1) / P.ctas_per_group;
  const int t_begin = cg * per;
  int t_end = t_begin + per;
  if (t_end > P.tiles_pg) t_end = P.tiles_pg;
  const bool do_bias = P.dbias != nullptr;

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(smem_u32(&full_bar[s]), P.use_tma ? 1 : 128);
        mbar_init(smem_u32(&empty_bar[s]), do_bias ? 5 : 1);
      }
      mbar_init(smem_u32(&acc_bar), P.dual ? 2 : 1);
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
  const int img_base = g * P.ipg;

  if (warp >= 4 && warp < 8 && !(P.dual && warp == 5)) {
    // ------------------------------------------------------------------ producers: X halo tile + dY tile per stage
    // The per-thread copy lists are tile-invariant (128 is a multiple of nb and nbo, so a thread always copies the same
    // channel block): X items from a precomputed table, dY items by a fixed pixel stride.
    if (P.use_tma) {
      // one thread, two 5-D boxes per stage: X (8 ch, 10 halo columns, nb blocks, 18 halo rows, image) -> [row][block][column][16 B]
      // and dY (8 ch, 8 columns, nbo blocks, 16 rows, image) -> [row][block][column][16 B]; out-of-image pixels are zero-filled
      if (tid == 128) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapD);
        const int tiles_y = P.tiles_per_img / P.tiles_x;
        int imgl = t_begin / P.tiles_per_img;
        int ty = (t_begin - imgl * P.tiles_per_img) / P.tiles_x;
        int tx = t_begin - imgl * P.tiles_per_img - ty * P.tiles_x;
        const uint32_t tx_bytes = (uint32_t)(kHHH * P.nb * 160 + 16 * P.nbo * 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t xs = smem_base + (uint32_t)stage * P.stage_bytes;
          if (P.dbg & 1) mbar_arrive(fb);
          else if (P.dbg & 8) {
            mbar_arrive_expect_tx(fb, (uint32_t)(16 * P.nbo * 128));
            if (P.dy_sw) { for (int b = 0; b < P.dy_blocks; ++b) tma_load_4d(xs + P.x_bytes + (uint32_t)b * P.dy_blk_bytes, &mapD, co0 + b * P.bo, tx * kHTW, ty * kHTH, img_base + imgl, fb); }
            else tma_load_5d(xs + P.x_bytes, &mapD, 0, tx * kHTW, co0 >> 3, ty * kHTH, img_base + imgl, fb);
          }
          else if (P.dbg & 16) { mbar_arrive_expect_tx(fb, (uint32_t)(kHHH * P.nb * 160)); tma_load_5d(xs, &mapX, 0, tx * kHTW - 1, 0, ty * kHTH - 1, img_base + imgl, fb); }
          else if (P.dy_sw) {
            mbar_arrive_expect_tx(fb, tx_bytes);
            tma_load_5d(xs, &mapX, 0, tx * kHTW - 1, 0, ty * kHTH - 1, img_base + imgl, fb);
            for (int b = 0; b < P.dy_blocks; ++b)
              tma_load_4d(xs + P.x_bytes + (uint32_t)b * P.dy_blk_bytes, &mapD, co0 + b * P.bo, tx * kHTW, ty * kHTH, img_base + imgl, fb);
          }
          else {
          mbar_arrive_expect_tx(fb, tx_bytes);
          tma_load_5d(xs, &mapX, 0, tx * kHTW - 1, 0, ty * kHTH - 1, img_base + imgl, fb);
          tma_load_5d(xs + P.x_bytes, &mapD, 0, tx * kHTW, co0 >> 3, ty * kHTH, img_base + imgl, fb);
          }
          if (++tx == P.tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++imgl; } }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    } else {
    const int ptid = tid - 128;
    const int nb = P.nb, nbo = P.nbo;
    const int x_items = kHPix * nb;
    constexpr int kMaxX = (kHPix * 8 + 127) / 128;          // 12
    uint32_t xdst[kMaxX];
    int xsrc[kMaxX], xhyx[kMaxX];
#pragma unroll
    for (int k = 0; k < kMaxX; ++k) {
      const int i = ptid + 128 * k;
      const int cb = i % nb, hp = i / nb;
      const int hy = hp / kHHW, hx = hp - hy * kHHW;
      xdst[k] = (uint32_t)((hy * nb + cb) * 160 + hx * 16);
      xsrc[k] = (hy * P.W + hx) * P.Cin + cb * 8;
      xhyx[k] = i < x_items ? ((hy << 8) | hx) : -1;
    }
    const int dcb = ptid % nbo, dp0 = ptid / nbo, dstep = 128 / nbo;      // pixel p = dp0 + k * dstep, k < nbo
    const int nmine = (x_items - ptid + 127) / 128;                       // this thread's X copies (the last slot is partial)
    // dY copies: dstep is a multiple of 8 (one tile row), so the thread keeps its tile column and walks down the rows
    const int dyy0 = dp0 >> 3, dxx = dp0 & 7, drows = dstep >> 3;
    const uint32_t ddst0 = (uint32_t)((dyy0 * nbo + dcb) * 128 + dxx * 16), ddst_step = (uint32_t)(drows * nbo * 128);
    const int dsrc0 = (dyy0 * P.W + dxx) * P.Cout, dsrc_step = drows * P.W * P.Cout;
    const int tiles_y = P.tiles_per_img / P.tiles_x;
    int imgl = t_begin / P.tiles_per_img;
    int ty = (t_begin - imgl * P.tiles_per_img) / P.tiles_x;
    int tx = t_begin - imgl * P.tiles_per_img - ty * P.tiles_x;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int y0 = ty * kHTH, x0 = tx * kHTW;
      const int64_t ibase = (int64_t)(img_base + imgl) * P.H * P.W;
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t xs = smem_base + (uint32_t)stage * P.stage_bytes;
      const uint32_t ds = xs + P.x_bytes;
      const bf16* xt = P.x + (ibase + (int64_t)(y0 - 1) * P.W + (x0 - 1)) * P.Cin;      // halo origin (may lie outside)
      const bf16* dt = P.dy + (ibase + (int64_t)y0 * P.W + x0) * P.Cout + co0 + dcb * 8;
      // tiles whose halo lies inside the image (3 of 4 at 160 x 192) copy without bounds tests: the producer warps' instruction
      // stream is what sets the tile period of this kernel (see halo_producer)
      const bool interior = ty > 0 && tx > 0 && y0 + kHTH + 1 <= P.H && x0 + kHTW + 1 <= P.W;
      if (interior && nb <= 8) {
#pragma unroll
        for (int k = 0; k < kMaxX; ++k)
          if (k < nmine) cp_async16_full(xs + xdst[k], xt + xsrc[k]);
        uint32_t dd = ds + ddst0;
        const bf16* dsrc = dt + dsrc0;
        for (int k = 0; k < nbo; ++k, dd += ddst_step, dsrc += dsrc_step) cp_async16_full(dd, dsrc);
      } else {
      if (nb <= 8) {
#pragma unroll
        for (int k = 0; k < kMaxX; ++k) {
          if (xhyx[k] >= 0) {
            const bool v = ((unsigned)(y0 - 1 + (xhyx[k] >> 8)) < (unsigned)P.H) && ((unsigned)(x0 - 1 + (xhyx[k] & 255)) < (unsigned)P.W);
            cp_async16(xs + xdst[k], v ? (const void*)(xt + xsrc[k]) : (const void*)P.x, v ? 16u : 0u);
          }
        }
      } else {                                     // 128 input channels: 16 blocks, the thread keeps its block, 8 pixels per round
        const int cb = ptid & 15;
        for (int hp = ptid >> 4; hp < kHPix; hp += 8) {
          const int hy = hp / kHHW, hx = hp - hy * kHHW;
          const bool v = ((unsigned)(y0 - 1 + hy) < (unsigned)P.H) && ((unsigned)(x0 - 1 + hx) < (unsigned)P.W);
          cp_async16(xs + (uint32_t)((hy * 16 + cb) * 160 + hx * 16), v ? (const void*)(xt + (hy * P.W + hx) * P.Cin + cb * 8) : (const void*)P.x,
                     v ? 16u : 0u);
        }
      }
      for (int k = 0, p = dp0; k < nbo; ++k, p += dstep) {
        const int yy = p >> 3, xx = p & 7;
        const bool v = (y0 + yy < P.H) && (x0 + xx < P.W);
        cp_async16(ds + (uint32_t)((yy * nbo + dcb) * 128 + xx * 16), v ? (const void*)(dt + ((int64_t)yy * P.W + xx) * P.Cout) : (const void*)P.dy,
                   v ? 16u : 0u);
      }
      }
      if (++tx == P.tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++imgl; } }
      cp_async_mbar_arrive(smem_u32(&full_bar[stage]));      // arrive-on-completion of this thread's copies, see halo_producer
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    cp_async_wait_all();
    }
  } else if (warp == 8 || warp == 5) {
    // ------------------------------------------------------------------ MMA issue (warp-uniform, one elected lane).  P.dual: warp 8 issues
    // the even tiles of the CTA's range into accumulator set 0, warp 5 the odd tiles into set 1 (the epilogue adds the sets): the barrier
    // wait / fence / commit of one issuer hide behind the other's MMAs instead of adding to them
    const int iss = warp == 8 ? 0 : 1;
    const uint32_t tmem_set = tmem_base + (uint32_t)iss * P.set_cols;
    const uint32_t idesc = make_idesc_mn2(128, P.cout_cta);
    const uint32_t xrow = (uint32_t)P.nb * 160u, drow = (uint32_t)P.nbo * 128u;     // one halo / tile row
    const uint64_t xdesc0 = make_desc_mn_nosw(smem_base, xrow, 160u);
    const uint64_t ddesc0 = P.dy_sw ? make_desc_mn_sw(smem_base + P.x_bytes, P.dy_row_bytes, P.dy_blk_bytes)
                                    : make_desc_mn_nosw(smem_base + P.x_bytes, drow, 128u);
    const uint32_t xrow16 = (2u * xrow) >> 4;                                        // one K-step (2 tile rows), 16-byte units
    const uint32_t drow16 = P.dy_sw ? (16u * P.dy_row_bytes) >> 4 : (2u * drow) >> 4;      // 16 pixels = 16 rows of the swizzled tile
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int t = t_begin; t < t_end; ++t) {
      if (!P.dual || ((t - t_begin) & 1) == iss) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tc_fence_after();
      const uint64_t soff = (uint64_t)(((uint32_t)stage * P.stage_bytes) >> 4);
      if (elect_one()) {
        const uint64_t xa0 = xdesc0 + soff, da0 = ddesc0 + soff;
        if (P.dbg & 2) {}
        else if (P.mt == 1) issue_wgrad_tile<1>(tmem_set, xa0, da0, xrow16, drow16, (uint32_t)P.cout_cta, idesc, first);
        else if (P.mt == 2) issue_wgrad_tile<2>(tmem_set, xa0, da0, xrow16, drow16, (uint32_t)P.cout_cta, idesc, first);
        else issue_wgrad_tile<3>(tmem_set, xa0, da0, xrow16, drow16, (uint32_t)P.cout_cta, idesc, first);
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      first = false;
      }
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&acc_bar));
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ warps 0-3: bias sums during the loop, epilogue after it
    const int et = tid;                                       // 0..127
    float bsum[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
    if (do_bias) {
      // Thread et sums the 8 channels of block cb at tile column px over the tile rows r0, r0 + 16 / nbo, ...: the 8 threads of a
      // quarter warp read 128 contiguous bytes (conflict free; with the channel block fastest, 128-byte strides put all eight on the
      // same banks and the 8-way replays also took shared-memory cycles from the MMAs' operand reads: sp6 gamma|beta 0.70 -> 0.47 ms).
      const int nbo = P.nbo;
      const int px = et & 7, cb = (et >> 3) % nbo, r0 = (et >> 3) / nbo;
      uint32_t off0 = (uint32_t)((r0 * nbo + cb) * 128 + px * 16);              // rows advance by 16 / nbo: always 2 KB
      // swizzled rows: thread et sums the 16-byte chunk scb of the pixels spl, spl + 128 / nbo, ... (a quarter warp reads 128 contiguous
      // bytes); the chunk's position inside its row is XORed with the row's swizzle phase
      const int scb = et % nbo, spl = et / nbo, sstep = 128 / nbo;
      const int cpr = (int)P.dy_row_bytes >> 4;                                 // 16-byte chunks per row
      const uint32_t sblk_off = P.dy_sw ? (uint32_t)(scb / cpr) * P.dy_blk_bytes : 0u;
      const int scl = P.dy_sw ? scb % cpr : 0;
      if (P.dy_sw) off0 = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        const uint8_t* ds = smem_raw + (smem_base - smem_u32(smem_raw)) + (uint32_t)stage * P.stage_bytes + P.x_bytes + off0;
        for (int k = 0; k < ((P.dbg & 4) ? 0 : nbo); ++k) {
          uint32_t koff = (uint32_t)k * 2048u;
          if (P.dy_sw) {
            const uint32_t prow = (uint32_t)(spl + k * sstep);
            const uint32_t sw = ((prow * P.dy_row_bytes) >> 7) & (uint32_t)(cpr - 1);
            koff = sblk_off + prow * P.dy_row_bytes + ((((uint32_t)scl) ^ sw) << 4);
          }
          const uint4 v = *reinterpret_cast<const uint4*>(ds + koff);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { bsum[2 * q] += __uint_as_float(w[q] << 16); bsum[2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u); }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) bias_red[et * 8 + k] = bsum[k];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < P.cout_cta && t_end > t_begin) {
        const int cbk = et >> 3, k = et & 7;
        float s = 0.f;
        if (P.dy_sw) {
          for (int pl = 0; pl < 128 / nbo; ++pl) s += bias_red[(pl * nbo + cbk) * 8 + k];
        } else
        for (int rg = 0; rg < 16 / nbo; ++rg)
          for (int x8 = 0; x8 < 8; ++x8) s += bias_red[((rg * nbo + cbk) * 8 + x8) * 8 + k];
        atomicAdd(P.dbias + (size_t)(P.dbias_gpr ? g / P.dbias_gpr : 0) * P.Cout + co0 + et, s);
      }
    }
    if (t_end > t_begin) {
      mbar_wait(smem_u32(&acc_bar), 0);
      tc_fence_after();
      const int q = warp;
      const int row = q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      float* dKg = P.dK + (size_t)g * P.Cout * 9 * P.Cin;
      for (int mt = 0; mt < P.mt; ++mt) {
        const int blk = mt * 16 + (row >> 3);                 // 8-channel block index = khslot * nb + cb
        const int kh = blk / P.nb, ci = (blk - kh * P.nb) * 8 + (row & 7);
        const bool rvalid = kh < 3;
        for (int kw = 0; kw < 3; ++kw) {
          for (int c0 = 0; c0 < P.cout_cta; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)((kw * P.mt + mt) * P.cout_cta + c0), r);
            if (P.dual && t_end - t_begin >= 2) {             // the second issuer's accumulator set (it had at least one tile)
              uint32_t r2[16];
              tmem_ld16(taddr + P.set_cols + (uint32_t)((kw * P.mt + mt) * P.cout_cta + c0), r2);
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
            }
            // A lane holds 16 output channels of ONE input channel; the 4 lanes of a quad hold 4 consecutive input channels of the same
            // 8-channel block.  A 4 x 4 transpose inside the quad (4 shuffles per 4 values) turns 16 scalar reductions per lane into
            // 4 16-byte ones (REDG.E.ADD.F32x4): the accumulators of all CTAs drain at the same moment at the end of the kernel and
            // the scalar REDs (10 M per launch for sp5 gamma|beta) queued up in the L2.
            const int m = lane & 3;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              float v0 = __uint_as_float(r[4 * k4]), v1 = __uint_as_float(r[4 * k4 + 1]), v2 = __uint_as_float(r[4 * k4 + 2]), v3 = __uint_as_float(r[4 * k4 + 3]);
              float a = (m & 1) ? v0 : v1;  a = __shfl_xor_sync(0xffffffffu, a, 1);  if (m & 1) v0 = a; else v1 = a;
              float b = (m & 1) ? v2 : v3;  b = __shfl_xor_sync(0xffffffffu, b, 1);  if (m & 1) v2 = b; else v3 = b;
              float c = (m & 2) ? v0 : v2;  c = __shfl_xor_sync(0xffffffffu, c, 2);  if (m & 2) v0 = c; else v2 = c;
              float d = (m & 2) ? v1 : v3;  d = __shfl_xor_sync(0xffffffffu, d, 2);  if (m & 2) v1 = d; else v3 = d;
              // now (v0..v3) = output channel co0 + c0 + 4 k4 + m for the quad's input channels (ci & ~3) .. + 3
              if (rvalid) {
                float* dst = dKg + ((size_t)(co0 + c0 + 4 * k4 + m) * 9 + kh * 3 + kw) * P.Cin + (ci & ~3);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
  }
}

bool g_wh_attr_set = false;

bool wgrad_halo_plan(const rd_conv_desc* d, WgHaloParams& P, int sm_count, int dy_sw = 1) {
  if (d->dtype != RD_BF16) return false;
  if (d->stride != 1 || d->kh != 3 || d->kw != 3 || d->pad != 1) return false;
  if (d->cin != 16 && d->cin != 32 && d->cin != 64 && d->cin != 128) return false;
  if (d->cout != 16 && d->cout != 32 && d->cout != 64 && d->cout != 128) return false;
  P.Cin = d->cin; P.Cout = d->cout;
  P.nb = d->cin / 8;
  P.mt = (3 * P.nb + 15) / 16;
  // TMEM holds 3 * mt accumulators of cout_cta columns: split Cout over CTAs when it does not fit (X is re-read per split)
  P.n_split = 1;
  P.cout_cta = d->cout;
  while (3 * P.mt * P.cout_cta > 512) { P.n_split *= 2; P.cout_cta /= 2; }
  if (P.cout_cta < 16 || P.n_split > 2) return false;
  if (P.nb == 16 && P.n_split > 1) return false;          // measured: 128 -> 64 @40x48 (sp4 out) is faster on k_wgrad_tma
  P.nbo = P.cout_cta / 8;
  if (d->groups * P.n_split > sm_count) return false;
  P.x_bytes = (uint32_t)(kHHH * P.nb * 160);
  // accumulator rows of the unused kh slots read up to 8 halo rows past the tile: keep that inside the stage
  uint32_t xpad = (uint32_t)((kHHH + 8) * P.nb * 160);
  P.dy_bytes = (uint32_t)(16 * P.nbo * 128);
  P.dy_sw = (dy_sw && (P.cout_cta >= 64 || dy_sw == 2)) ? 1 : 0;      // measured: pays for 128-byte rows (sp6 gamma|beta 0.472 -> 0.411 ms), neutral below
  P.bo = P.cout_cta < 64 ? P.cout_cta : 64;
  P.dy_blocks = P.cout_cta / P.bo;
  P.dy_row_bytes = (uint32_t)P.bo * 2u;
  P.dy_blk_bytes = 128u * P.dy_row_bytes;                 // 4 / 8 / 16 KB: whole 1 KB swizzle atoms
  if (P.dy_sw) P.x_bytes = (P.x_bytes + 1023u) & ~1023u;  // (from here on x_bytes is only the offset of the dY tile inside a stage)
  if (P.x_bytes + P.dy_bytes < xpad) P.dy_bytes = xpad - P.x_bytes;
  P.stage_bytes = (P.x_bytes + P.dy_bytes + (P.dy_sw ? 1023u : 127u)) & (P.dy_sw ? ~1023u : ~127u);
  int st = (int)((200u * 1024u) / P.stage_bytes);
  P.stages = st > kWHMaxStages ? kWHMaxStages : st;
  if (P.stages < 2) return false;
  P.lag = P.stages - 1 < 3 ? P.stages - 1 : 3;
  return true;
}

}  // namespace

int rd_wgrad_halo_supported(const rd_conv_desc* d, int sm_count) {
  static const bool off = getenv("RD_B200_NO_WGRAD_HALO") != nullptr;
  if (off) return 0;
  WgHaloParams P;
  if (!wgrad_halo_plan(d, P, sm_count)) return 0;
  if (d->algo == RD_ALGO_HALO) return 1;
  int64_t tiles = (int64_t)d->n * rd_div_up(d->h, kHTH) * rd_div_up(d->w, kHTW);
  return tiles >= 4 * (int64_t)sm_count;
}

int rd_wgrad_halo_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias, cudaStream_t st) {
  WgHaloParams P;
  int want_sw = 1;
  {
    const char* e_tma0 = getenv("RD_B200_HALO_TMA");
    const char* e_sw = getenv("RD_B200_WGH_DYSW");           // A/B switch: 0 = the no-swizzle dY planes
    if (rd_tensormap_encode_fn() == nullptr || (e_tma0 && atoi(e_tma0) == 0) || (e_sw && atoi(e_sw) == 0)) want_sw = 0;
    else if (e_sw && atoi(e_sw) == 2) want_sw = 2;            // 2 = also for the 64- / 32-byte rows (tests)
  }
  if (!wgrad_halo_plan(d, P, ctx->sm_count, want_sw)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "wgrad_halo: shape not supported");
  P.x = (const bf16*)x; P.dy = (const bf16*)dy; P.dK = dK; P.dbias = dbias;
  P.H = d->h; P.W = d->w;
  P.ipg = d->n / d->groups; P.groups = d->groups;
  P.tiles_x = rd_div_up(d->w, kHTW);
  P.tiles_per_img = P.tiles_x * rd_div_up(d->h, kHTH);
  P.tiles_pg = P.ipg * P.tiles_per_img;
  P.ctas_per_group = ctx->sm_count / (d->groups * P.n_split);
  if (P.ctas_per_group > P.tiles_pg) P.ctas_per_group = P.tiles_pg;
  if (P.ctas_per_group < 1) P.ctas_per_group = 1;
  P.set_cols = (uint32_t)(3 * P.mt * P.cout_cta);
  { const char* e_du = getenv("RD_B200_HALO_DUAL"); P.dual = (e_du && atoi(e_du) == 0) ? 0 : 1; }      // A/B switch, read per call
  if (2u * P.set_cols > 512u) P.dual = 0;                 // both accumulator sets must fit in tensor memory
  P.dbias_gpr = d->bias_groups > 1 ? d->groups / d->bias_groups : 0;
  { const char* e_dbg = getenv("RD_B200_WGH_DEBUG"); P.dbg = e_dbg ? atoi(e_dbg) : 0; }
  alignas(64) CUtensorMap mapX, mapD;
  memset(&mapX, 0, sizeof(mapX));
  memset(&mapD, 0, sizeof(mapD));
  {
    static const char* e_tma = getenv("RD_B200_HALO_TMA");
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiledFn enc = (EncodeTiledFn)rd_tensormap_encode_fn();
    P.use_tma = (enc != nullptr && !(e_tma && atoi(e_tma) == 0)) ? 1 : 0;
    if (P.use_tma) {
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      {
        cuuint64_t dims[5] = {8u, (cuuint64_t)d->w, (cuuint64_t)(d->cin / 8), (cuuint64_t)d->h, (cuuint64_t)d->n};
        cuuint64_t strides[4] = {(cuuint64_t)d->cin * 2, 16u, (cuuint64_t)d->w * d->cin * 2, (cuuint64_t)d->h * d->w * d->cin * 2};
        cuuint32_t box[5] = {8u, (cuuint32_t)kHHW, (cuuint32_t)P.nb, (cuuint32_t)kHHH, 1u};
        CUresult r = enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo X) failed: %d", (int)r);
      }
      if (P.dy_sw) {
        cuuint64_t dims[4] = {(cuuint64_t)d->cout, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
        cuuint64_t strides[3] = {(cuuint64_t)d->cout * 2, (cuuint64_t)d->w * d->cout * 2, (cuuint64_t)d->h * d->w * d->cout * 2};
        cuuint32_t box[4] = {(cuuint32_t)P.bo, (cuuint32_t)kHTW, (cuuint32_t)kHTH, 1u};
        cuuint32_t es4[4] = {1, 1, 1, 1};
        const CUtensorMapSwizzle swz = P.bo == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (P.bo == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUresult r = enc(&mapD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, es4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo dY, swizzled rows) failed: %d", (int)r);
      } else {
        cuuint64_t dims[5] = {8u, (cuuint64_t)d->w, (cuuint64_t)(d->cout / 8), (cuuint64_t)d->h, (cuuint64_t)d->n};
        cuuint64_t strides[4] = {(cuuint64_t)d->cout * 2, 16u, (cuuint64_t)d->w * d->cout * 2, (cuuint64_t)d->h * d->w * d->cout * 2};
        cuuint32_t box[5] = {8u, (cuuint32_t)kHTW, (cuuint32_t)P.nbo, (cuuint32_t)kHTH, 1u};
        CUresult r = enc(&mapD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo dY) failed: %d", (int)r);
      }
    }
  }
  if (!P.use_tma) P.dual = 0;                             // warp 5 is a cp.async producer in the fallback path
  if (P.dual) P.stages &= ~1;                             // even ring: a stage always belongs to the same issuer (see rd_conv_halo_launch)
  {
    uint32_t cols = 32;
    while (cols < P.set_cols * (P.dual ? 2u : 1u)) cols <<= 1;
    P.tmem_cols = cols;
  }
  size_t smem = (size_t)P.stages * P.stage_bytes + 1024;
  if (!g_wh_attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_wgrad_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    g_wh_attr_set = true;
  }
  k_wgrad_halo<<<P.ctas_per_group * d->groups * P.n_split, kWHThreads, smem, st>>>(mapX, mapD, P);
  RD_CHECK_LAUNCH(ctx, "wgrad_halo");
  return RD_OK;
}
