// rd_conv_tma.cu — persistent, TMA-fed tcgen05 implicit-GEMM convolution for the stride-1 "same" convolutions
// (3x3 pad 1 and 1x1 pad 0: every SPADE gamma/beta/out convolution and the anatomy-decoder convolutions,
// ~93 % of the step's FLOPs), forward and dgrad.
//
// One CTA per SM, looping over output tiles.  A tile is a rectangle of TN images x TH rows x TW columns
// (<= 128 pixels = the UMMA M dimension / the 128 TMEM lanes).  For filter tap (kh, kw) the A operand is the SAME
// rectangle shifted by (kh - pad, kw - pad): one 4-D TMA box load {kc channels, TW, TH, TN} from the NHWC tensor
// at coordinates {c0, x0 + kw - pad, y0 + kh - pad, n0}; out-of-bounds elements are zero-filled by the TMA unit,
// which IS the convolution's zero padding.  No im2col buffer, no per-thread address arithmetic.  The B operand
// (packed CondConv weights [G*Cout][taps*Cin]) is a 2-D TMA box {kc, n_tile}.  Both land in shared memory in the
// K-major swizzled layout (128B / 64B / 32B swizzle for kc = 64 / 32 / 16 channels) the UMMA descriptors expect.
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected thread, mbarrier expect_tx), warp 1 = MMA issuer
// (one thread, tcgen05.mma M=128 N=n_tile K=16, fp32 accumulators in TMEM, tcgen05.commit releases stages),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b, bias + LeakyReLU, bf16, 16-byte NHWC stores).
// Three pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty (MMA <-> epilogue),
// and the persistent tile loop — the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <tuple>
#include <vector>
#include "rd_common.cuh"
#include "rd_tc_common.cuh"

namespace {

constexpr int kTmaThreads = 256;
constexpr int kTmaMaxStages = 8;

struct TmaParams {
  const float* bias; bf16* y;
  int H, W, Cin, Cout;
  int taps, KW, pad, sign;   // sign = +1 forward (shift = k - pad), -1 dgrad (shift = pad - k)
  int stride;                // 1, or 2 (forward only): the box walks the input with TMA element strides, H / W are the OUTPUT size
  int TW, TH, TN;            // tile rectangle; rows_valid = TW*TH*TN <= 128
  int tiles_x, tiles_y, img_blocks_pg, ipg;   // per group: img_blocks_pg * tiles_y * tiles_x pixel tiles
  int ptiles_total;          // pixel tiles over all groups
  int n_tile, n_tiles;       // UMMA N and number of N tiles
  int kc, k_chunks;          // channels per K-block, K-blocks per tap
  int stages;
  int dual;                  // two MMA-issuing warps (1 and 3), each on one half of the N tile (see k_conv_tma)
  uint32_t a_bytes, b_bytes; // smem bytes per stage (A: 128 rows; B: n_tile rows rounded up to 1 KB)
  uint32_t tx_bytes;         // bytes the two TMA boxes deliver per stage (full boxes, OOB parts zero-filled)
  uint32_t tmem_cols;
  int act; float slope;
  int bias_gpr;              // weight groups per bias row (0: one bias row for all groups)
  // Tap table.  classes = 1: the taps of the convolution in (kh, kw) order.  classes = 4: the input gradient of a stride-2 convolution,
  // one stride-1 problem per output-parity class cls = 2 (y & 1) + (x & 1): dx(2u + py, 2v + px) only receives the taps with
  // kh = py + 1 (mod 2), kw = px + 1 (mod 2) — 2 x 2 taps for k = 4; 1, 2, 2 or 4 for k = 3 — each reading dy at (u + sy, v + sx),
  // sy, sx in {-1, 0, 1}.  The tile grid then lives on the dy image and the epilogue writes every second pixel of every second row.
  int classes;
  int tps_max;               // stage = tps_max A boxes followed by tps_max B boxes
  int tps[4];                // taps per pipeline stage of class cls (divides ntaps[cls]); > 1 for the 16- / 32-channel layers, whose
                             // single-tap stages (4 KB + 1 KB) were bound by the barrier round trip per stage, not by the MMAs
  int ntaps[4];
  signed char tap_sy[4][16], tap_sx[4][16];      // A-box shift of tap j (input coordinates, relative to the tile origin x0 * stride)
  unsigned char tap_w[4][16];                    // its tap index in the packed weights (column block tap_w * Cin)
};

__global__ void __launch_bounds__(kTmaThreads, 1)
k_conv_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TmaParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTmaMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kTmaMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (P.a_bytes + P.b_bytes) * (uint32_t)P.tps_max;
  const int S = P.stages;
  const int tiles_per_class = P.ptiles_total * P.n_tiles;
  const int total_tiles = tiles_per_class * P.classes;       // the parity class is the SLOWEST tile index: every CTA gets its share of each class
  const uint32_t row_bytes = (uint32_t)P.kc * 2u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), P.dual ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), P.dual ? 2 : 1);
      mbar_init(smem_u32(&acc_empty[b]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp in warp-uniform control flow, one elected lane issues: coordinates and descriptors stay in uniform
    //  registers instead of being re-derived per thread — see the note in rd_conv_halo.cu)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int pt = t;
        int cls = 0;
        if (P.classes > 1) { cls = pt / tiles_per_class; pt -= cls * tiles_per_class; }
        const int nt = pt % P.n_tiles; pt /= P.n_tiles;
        const int tx = pt % P.tiles_x; pt /= P.tiles_x;
        const int ty = pt % P.tiles_y; pt /= P.tiles_y;
        const int ib = pt % P.img_blocks_pg;
        const int g = pt / P.img_blocks_pg;
        const int img0 = g * P.ipg + ib * P.TN;
        const int x0 = tx * P.TW * P.stride, y0 = ty * P.TH * P.stride;      // input coordinates of the tile's first output pixel
        const int wrow = g * P.Cout + nt * P.n_tile;
        const int ntap = P.ntaps[cls], tps = P.tps[cls];
        // K-block order: channel chunk outer, taps inner (neighbouring boxes stay in L2); table look-ups instead of div / mod —
        // this single thread must issue the TMA loads faster than the tensor core consumes a stage.  A stage holds `tps` taps:
        // [A box] x tps, then [B box] x tps.
        for (int chunk = 0; chunk < P.k_chunks; ++chunk) {
          const int c0 = chunk * P.kc;
          for (int j = 0; j < ntap; j += tps) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t a_s = smem_base + (uint32_t)stage * stage_bytes;
            const uint32_t fb = smem_u32(&full_bar[stage]);
            if (elect_one()) {
              mbar_arrive_expect_tx(fb, P.tx_bytes * (uint32_t)tps);
              for (int u = 0; u < tps; ++u) {
                tma_load_4d(a_s + (uint32_t)u * P.a_bytes, &mapA, c0, x0 + (int)P.tap_sx[cls][j + u], y0 + (int)P.tap_sy[cls][j + u], img0, fb);
                tma_load_2d(a_s + (uint32_t)P.tps_max * P.a_bytes + (uint32_t)u * P.b_bytes, &mapB, c0 + (int)P.tap_w[cls][j + u] * P.Cin, wrow, fb);
              }
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || (warp == 3 && P.dual)) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, elected lane issues).
    // P.dual: a tcgen05.mma occupies its issuing thread, so the waits / fences / commits between the stages ADD to the MMAs of a single
    // issuer (tools/mma_rate.cu, profiles/r02_mma_issue_microbench.txt).  Two issuers — warp 1 on accumulator columns [0, n_tile / 2),
    // warp 3 on [n_tile / 2, n_tile), each with its half of the B rows — overlap their overheads; a stage and an accumulator are
    // released by both commits (barrier count 2).
    {
      const int iss = warp == 1 ? 0 : 1;
      const int n_iss = P.dual ? P.n_tile / 2 : P.n_tile;
      const uint32_t idesc = make_idesc(128, n_iss);
      const uint64_t desc0 = make_desc_k(smem_base, row_bytes);      // descriptors are affine in the stage index
      const int ksteps = P.kc >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&acc_empty[buf]), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)buf * (uint32_t)P.n_tile + (uint32_t)(iss * n_iss);
        const int mcls = P.classes > 1 ? t / tiles_per_class : 0;
        const int tps = P.tps[mcls];
        const int kb_per_tile = P.ntaps[mcls] / tps * P.k_chunks;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint64_t adesc = desc0 + (uint64_t)(((uint32_t)stage * stage_bytes) >> 4);
          const uint64_t bdesc = adesc + (uint64_t)(((uint32_t)P.tps_max * P.a_bytes + (uint32_t)(iss * n_iss) * row_bytes) >> 4);
          if (elect_one()) {
            if (tps > 1) {                        // several taps per stage (kc = 16 / 32)
              for (int u = 0; u < tps; ++u) {
                const uint64_t au = adesc + (uint64_t)(((uint32_t)u * P.a_bytes) >> 4), bu = bdesc + (uint64_t)(((uint32_t)u * P.b_bytes) >> 4);
                for (int k = 0; k < ksteps; ++k) umma_bf16(tacc, au + (uint64_t)(2 * k), bu + (uint64_t)(2 * k), idesc, (uint32_t)((kb | u | k) != 0));
              }
            } else if (ksteps == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            } else if (ksteps == 2) {
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            } else {
              umma_bf16(tacc, adesc, bdesc, idesc, (uint32_t)(kb != 0));
            }
            umma_commit(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit(smem_u32(&acc_full[buf]));
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (warps 4..7 <-> TMEM lane quarters 0..3)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rows_valid = P.TW * P.TH * P.TN;
    const int wl = row % P.TW;
    const int hl = (row / P.TW) % P.TH;
    const int nl = row / (P.TW * P.TH);
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      int pt = t;
      int cls = 0;
      if (P.classes > 1) { cls = pt / tiles_per_class; pt -= cls * tiles_per_class; }
      const int nt = pt % P.n_tiles; pt /= P.n_tiles;
      const int tx = pt % P.tiles_x; pt /= P.tiles_x;
      const int ty = pt % P.tiles_y; pt /= P.tiles_y;
      const int ib = pt % P.img_blocks_pg;
      const int g = pt / P.img_blocks_pg;
      const int n0 = nt * P.n_tile;
      const bool pvalid = row < rows_valid;
      const int64_t pix = P.classes > 1
          ? ((int64_t)(g * P.ipg + ib * P.TN + nl) * (2 * P.H) + (2 * (ty * P.TH + hl) + (cls >> 1))) * (2 * P.W) + (2 * (tx * P.TW + wl) + (cls & 1))
          : ((int64_t)(g * P.ipg + ib * P.TN + nl) * P.H + (ty * P.TH + hl)) * P.W + (tx * P.TW + wl);
      bf16* yrow = P.y + (pvalid ? pix : 0) * P.Cout + n0;
      const float* biasg = P.bias ? P.bias + (size_t)(P.bias_gpr ? g / P.bias_gpr : 0) * P.Cout : nullptr;
      mbar_wait(smem_u32(&acc_full[buf]), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * (uint32_t)P.n_tile;
      for (int cb = 0; cb < P.n_tile; cb += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)cb, r);
        if (pvalid && (P.Cout & 7)) {
#pragma unroll
          for (int qq = 0; qq < 16; ++qq) {
            int co = n0 + cb + qq;
            if (co < P.Cout) {
              float v0 = __uint_as_float(r[qq]);
              if (biasg) v0 += biasg[co];
              if (P.act == RD_ACT_LRELU) v0 = v0 > 0.f ? v0 : v0 * P.slope;
              yrow[cb + qq] = __float2bfloat16_rn(v0);
            }
          }
        } else if (pvalid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int co = n0 + cb + h * 8;
            if (co < P.Cout) {
              uint32_t packed[4];
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                float v0 = __uint_as_float(r[h * 8 + 2 * qq]), v1 = __uint_as_float(r[h * 8 + 2 * qq + 1]);
                if (biasg) { v0 += biasg[co + 2 * qq]; v1 += biasg[co + 2 * qq + 1]; }
                if (P.act == RD_ACT_LRELU) { v0 = v0 > 0.f ? v0 : v0 * P.slope; v1 = v1 > 0.f ? v1 : v1 * P.slope; }
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                packed[qq] = *reinterpret_cast<uint32_t*>(&b2);
              }
              *reinterpret_cast<uint4*>(yrow + cb + h * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

EncodeTiledFn get_encode() {
  if (!g_encode_tried) {
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  }
  return g_encode;
}

// largest rectangle TN x TH x TW <= 128 that tiles (ipg, H, W) exactly
bool choose_tile(int ipg, int H, int W, int& TN, int& TH, int& TW) {
  int best = 0;
  for (int tw = 1; tw <= W && tw <= 128; ++tw) {
    if (W % tw) continue;
    for (int th = 1; th <= H && tw * th <= 128; ++th) {
      if (H % th) continue;
      int tn = 1;
      if (tw == W && th == H) {
        for (int c = 1; c <= ipg && tw * th * c <= 128; ++c)
          if (ipg % c == 0) tn = c;
      }
      int px = tw * th * tn;
      // prefer more pixels; tie-break towards wider rows (longer contiguous runs in NHWC)
      if (px > best || (px == best && tw > TW)) { best = px; TN = tn; TH = th; TW = tw; }
    }
  }
  return best >= 64;      // below half a tile the gather kernel's linear tiling wastes less
}

bool g_attr_set = false;

}  // namespace

void* rd_tensormap_encode_fn() { return (void*)get_encode(); }

int rd_conv_tma_supported(const rd_conv_desc* d, int mode) {
  if (d->dtype != RD_BF16) return 0;
  static const bool s2 = getenv("RD_B200_NO_TMA_S2") == nullptr;
  // stride 2 (the encoders' k4 / k3 pad-1 convolutions), forward only: same kernel, the A box walks the input with element strides
  // ... and their input gradient as four stride-1 problems, one per output-parity class (see TmaParams::classes)
  static const bool s2d = getenv("RD_B200_NO_TMA_S2_DGRAD") == nullptr;
  const bool strided = s2 && (mode == 0 || s2d) && d->stride == 2 && d->kh == d->kw && (d->kh == 3 || d->kh == 4) && d->pad == 1 &&
                       d->oh * 2 == d->h && d->ow * 2 == d->w;
  if (!strided) {
    if (d->stride != 1 || d->kh != d->kw || (d->kh != 1 && d->kh != 3) || d->pad != (d->kh - 1) / 2) return 0;
    if (d->oh != d->h || d->ow != d->w) return 0;
  }
  int cin = mode == 0 ? d->cin : d->cout;
  if (cin % 16) return 0;
  int TN, TH, TW = 0;
  if (!choose_tile(d->n / d->groups, d->oh, d->ow, TN, TH, TW)) return 0;
  if (strided && mode == 0 && ((TW - 1) * 2 + 1 > 256 || (TH - 1) * 2 + 1 > 256)) return 0;
  if (!get_encode()) return 0;
  return 1;
}

int rd_conv_tma_launch(rd_ctx* ctx, const rd_conv_desc* d, int mode, const void* x, const void* w, const float* bias, void* y,
                       cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  TmaParams P;
  P.bias = bias; P.y = (bf16*)y;
  const bool s2_dgrad = mode == 1 && d->stride == 2;
  P.stride = (mode == 0) ? d->stride : 1;
  P.H = (mode == 0 || s2_dgrad) ? d->oh : d->h; P.W = (mode == 0 || s2_dgrad) ? d->ow : d->w;      // the tile grid lives on the OUTPUT image (stride-2 dgrad: on dy, per class)
  P.Cin = mode == 0 ? d->cin : d->cout;
  P.Cout = mode == 0 ? d->cout : d->cin;
  P.taps = d->kh * d->kw; P.KW = d->kw; P.pad = d->pad; P.sign = mode == 0 ? 1 : -1;
  memset(P.ntaps, 0, sizeof(P.ntaps));
  memset(P.tap_sy, 0, sizeof(P.tap_sy)); memset(P.tap_sx, 0, sizeof(P.tap_sx)); memset(P.tap_w, 0, sizeof(P.tap_w));
  if (s2_dgrad) {
    P.classes = 4;
    for (int cls = 0; cls < 4; ++cls) {
      const int py = cls >> 1, px = cls & 1;
      int n = 0;
      for (int kh = 0; kh < d->kh; ++kh) {
        if (((py + d->pad - kh) & 1) != 0) continue;
        for (int kw = 0; kw < d->kw; ++kw) {
          if (((px + d->pad - kw) & 1) != 0) continue;
          // dx(2u + py) <- dy(u + (py + pad - kh) / 2): exact division, the numerator is even
          P.tap_sy[cls][n] = (signed char)((py + d->pad - kh) / 2);
          P.tap_sx[cls][n] = (signed char)((px + d->pad - kw) / 2);
          P.tap_w[cls][n] = (unsigned char)(kh * d->kw + kw);
          ++n;
        }
      }
      P.ntaps[cls] = n;
    }
  } else {
    P.classes = 1;
    if (P.taps > 16) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv_tma: more than 16 taps");
    int n = 0;
    for (int kh = 0; kh < d->kh; ++kh)
      for (int kw = 0; kw < d->kw; ++kw, ++n) {
        P.tap_sy[0][n] = (signed char)(P.sign * (kh - P.pad));
        P.tap_sx[0][n] = (signed char)(P.sign * (kw - P.pad));
        P.tap_w[0][n] = (unsigned char)n;
      }
    P.ntaps[0] = n;
  }
  P.ipg = d->n / d->groups;
  P.TW = 0;
  if (!choose_tile(P.ipg, P.H, P.W, P.TN, P.TH, P.TW)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv_tma: no exact tiling");
  P.tiles_x = P.W / P.TW; P.tiles_y = P.H / P.TH; P.img_blocks_pg = P.ipg / P.TN;
  P.ptiles_total = d->groups * P.img_blocks_pg * P.tiles_y * P.tiles_x;
  P.kc = (P.Cin % 64 == 0) ? 64 : ((P.Cin % 32 == 0) ? 32 : 16);
  P.k_chunks = P.Cin / P.kc;
  int n_tile = ((P.Cout + 15) / 16) * 16;
  if (n_tile > 256) n_tile = 256;
  P.n_tile = n_tile;
  P.n_tiles = rd_div_up(P.Cout, n_tile);
  P.a_bytes = 128u * (uint32_t)P.kc * 2u;
  P.b_bytes = (uint32_t)n_tile * (uint32_t)P.kc * 2u;
  // smem rows of the A tile that the TMA box does not cover (rows_valid..127) keep stale data: harmless, their
  // accumulator rows are never stored.  Stage bases stay 1024-byte aligned: a_bytes, b_bytes are multiples of 1024
  // for kc = 64; for kc = 32 / 16 pad b_bytes up.
  P.b_bytes = (P.b_bytes + 1023u) & ~1023u;
  // taps per stage: the small-channel layers (kc 16 / 32: 4-8 KB of A per tap) put up to 4 taps into one stage
  P.tps_max = 1;
  for (int cls = 0; cls < 4; ++cls) {
    P.tps[cls] = 1;
    static const bool multi = getenv("RD_B200_TMA_NO_MULTITAP") == nullptr;
    if (multi && P.kc <= 32 && P.ntaps[cls] > 1) {
      const int cap = P.kc == 16 ? 4 : 2;
      for (int t = cap; t > 1; --t)
        if (P.ntaps[cls] % t == 0) { P.tps[cls] = t; break; }
    }
    if (cls < P.classes && P.tps[cls] > P.tps_max) P.tps_max = P.tps[cls];
  }
  uint32_t stage_bytes = (P.a_bytes + P.b_bytes) * (uint32_t)P.tps_max;
  int stages = (int)((200u * 1024u) / stage_bytes);
  if (stages > kTmaMaxStages) stages = kTmaMaxStages;
  if (stages < 2) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "conv_tma: stage too large");
  P.stages = stages;
  {
    // Measured (profiles/r02_bench_conv_dual_v18.txt): NO gain for this kernel — its MMAs are wide (N = 128 / 256: 64 / 131 cycles), the
    // issuing thread is only held for the ~40 cycles of the operand fetch, so the per-stage overhead already hides behind them — and
    // 2-10 % slower on the low-resolution layers.  Off by default; RD_B200_TMA_DUAL_FWD=<n> enables it from n_tile >= n.
    const char* e_du = getenv("RD_B200_TMA_DUAL_FWD");
    const int min_n = e_du ? atoi(e_du) : 0;
    P.dual = (min_n > 0 && n_tile >= min_n && n_tile % 32 == 0) ? 1 : 0;
  }
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.act = mode == 0 ? d->act : RD_ACT_NONE;
  P.slope = d->act_slope;
  P.bias_gpr = (mode == 0 && d->bias_groups > 1) ? d->groups / d->bias_groups : 0;

  CUtensorMapSwizzle sw = P.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (P.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  alignas(64) CUtensorMap mapA, mapB;
  {
    const int ih = P.H * P.stride, iw = P.W * P.stride;                   // input image (= output for stride 1; dy for the stride-2 dgrad)
    const cuuint32_t sst = (cuuint32_t)P.stride;
    cuuint64_t dims[4] = {(cuuint64_t)P.Cin, (cuuint64_t)iw, (cuuint64_t)ih, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)P.Cin * 2, (cuuint64_t)iw * P.Cin * 2, (cuuint64_t)ih * iw * P.Cin * 2};
    // with element strides the box is given as the span it covers in the tensor: ceil(span / stride) elements are delivered
    cuuint32_t box[4] = {(cuuint32_t)P.kc, (cuuint32_t)((P.TW - 1) * P.stride + 1), (cuuint32_t)((P.TH - 1) * P.stride + 1), (cuuint32_t)P.TN};
    cuuint32_t es[4] = {1, sst, sst, 1};
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    const int k_total = P.taps * P.Cin;
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)d->groups * P.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)P.kc, (cuuint32_t)n_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  // expect_tx counts the FULL boxes (out-of-bounds parts are zero-filled and still counted)
  P.tx_bytes = (uint32_t)(P.TW * P.TH * P.TN) * (uint32_t)P.kc * 2u + (uint32_t)n_tile * (uint32_t)P.kc * 2u;
  size_t smem = (size_t)stages * stage_bytes + 1024;
  if (!g_attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_conv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    g_attr_set = true;
  }
  int total_tiles = P.ptiles_total * P.n_tiles * P.classes;
  int grid = total_tiles < ctx->sm_count ? total_tiles : ctx->sm_count;
  k_conv_tma<<<grid, kTmaThreads, smem, st>>>(mapA, mapB, P);
  RD_CHECK_LAUNCH(ctx, mode == 0 ? "conv_tma_fwd" : "conv_tma_dgrad");
  return RD_OK;
}

// =====================================================================================================
// TMA-fed wgrad for the stride-1 "same" convolutions:
//     dK[g][co][tap][ci] = sum_p dY[p, co] * X[p + shift(tap), ci]
// The reduction runs over PIXELS, so both operands are MN-major: a TMA box {channels, TW, TH, TN} lands in shared
// memory as [pixel][channel block] rows — exactly the MN-major swizzled layout (block of <= 64 channels, 8-pixel groups
// SBO apart, channel blocks LBO apart).  The dY box is loaded once per pixel tile, the X box once per (tap, channel
// block) with the tap's shift (zero fill = padding).  One tcgen05.mma covers 16 pixels.
//   normal     (Cout >= 128): D[co 128][n' <= 256]       A = 2 dY blocks,          B = the CTA's X boxes
//   transposed (Cout <  128): D[n' 128][co]  per M tile   A = 128/bi X boxes,       B = dY blocks
// n' = tap*Cin + ci enumerates X boxes in order, so dK[g][co][n'] is addressed directly.  Split-K over pixel-tile
// chunks; the epilogue adds the TMEM accumulators into dK with red.global.add.f32.
namespace {

constexpr int kWgTmaMaxStages = 6;

struct WgTmaParams {
  float* dK;
  float* dbias;                    // optional: bias gradient = dY^T * ones, one extra N=16 MMA per K-step
  int dbias_gpr;                   // weight groups per bias-gradient row (0: one row)
  uint32_t ones_off;               // smem offset of the all-ones [p_rows][16] block (after the stages)
  uint32_t bias_col;               // TMEM column of the bias accumulator
  int H, W, Cin, Cout, KW, pad;    // H, W: the OUTPUT image (= input for stride 1)
  int stride;                      // 1 or 2: the X boxes walk the input with TMA element strides
  int TW, TH, TN, p_rows;          // pixel tile (K-block): p_rows = TW*TH*TN, multiple of 16, <= 64
  int tiles_x, tiles_y, img_blocks_pg, ipg, ptiles_pg;
  int bi, bo;                      // channels per X / dY box
  int ci_blocks;                   // Cin / bi
  int xb_total, xb_per_cta, xsplits;
  int transposed, dy_blocks;       // dY boxes per stage
  int n_total;
  int chunk_tiles, chunks_pg;
  uint32_t dy_blk_bytes, x_blk_bytes, stage_bytes, tx_dy, tx_x;
  int stages;
  int dual;                        // two MMA-issuing warps (1 and 3): the accumulator column groups / M tiles alternate between them
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
  uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  uint64_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16);
  uint64_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
  return lo | (hi << 32);
}
__device__ __forceinline__ uint32_t make_idesc_mnmn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kTmaThreads, 1)
k_wgrad_tma(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDY, const WgTmaParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgTmaMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgTmaMaxStages];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = P.stages;
  const int g = blockIdx.z;
  const int chunk = blockIdx.x / P.xsplits;
  const int xs = blockIdx.x - chunk * P.xsplits;
  const int co0 = P.transposed ? 0 : blockIdx.y * 128;
  const int xb0 = xs * P.xb_per_cta;
  int nxb = P.xb_total - xb0;
  if (nxb > P.xb_per_cta) nxb = P.xb_per_cta;
  const int t0 = chunk * P.chunk_tiles;
  int t1 = t0 + P.chunk_tiles;
  if (t1 > P.ptiles_pg) t1 = P.ptiles_pg;
  const int k_blocks = t1 - t0;
  const uint32_t x_off = (uint32_t)P.dy_blocks * P.dy_blk_bytes;     // X boxes follow the dY boxes in a stage
  const bool do_bias = (P.dbias != nullptr) && (xs == 0);
  // two issuers when this CTA has at least two accumulator groups (N <= 256 column groups, or M tiles in the transposed form): group i
  // belongs to issuer i & 1, so no accumulator is touched by both (see k_conv_tma for why two issuers)
  const int grp_boxes = P.transposed ? 128 / P.bi : 256 / P.bi;
  const bool dualw = P.dual && nxb > grp_boxes;
  if (do_bias) {
    // [p_rows][16] bf16 block of ones (32-byte rows; all-ones is invariant under the 32B swizzle)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_raw + (smem_base - smem_u32(smem_raw)) + P.ones_off);
    for (int i = tid; i < P.p_rows * 8; i += kTmaThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapDY);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), dualw ? 2 : 1);
    }
    mbar_init(smem_u32(&acc_bar), dualw ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // producer: whole warp in warp-uniform control flow, one elected lane issues the TMA loads
    int stage = 0;
    uint32_t phase = 0;
    int pt0 = t0;
    int tx = pt0 % P.tiles_x; pt0 /= P.tiles_x;
    int ty = pt0 % P.tiles_y; pt0 /= P.tiles_y;
    int ib = pt0;                                  // image block within the group
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int img0 = g * P.ipg + ib * P.TN;
      const int x0 = tx * P.TW, y0 = ty * P.TH;
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
      const uint32_t st = smem_base + (uint32_t)stage * P.stage_bytes;
      const uint32_t fb = smem_u32(&full_bar[stage]);
      if (elect_one()) {
        mbar_arrive_expect_tx(fb, (uint32_t)P.dy_blocks * P.tx_dy + (uint32_t)nxb * P.tx_x);
        for (int b = 0; b < P.dy_blocks; ++b)
          tma_load_4d(st + (uint32_t)b * P.dy_blk_bytes, &mapDY, co0 + b * P.bo, x0, y0, img0, fb);
        int tap = xb0 / P.ci_blocks, cib = xb0 - tap * P.ci_blocks;
        int kh = tap / P.KW, kw = tap - kh * P.KW;
        for (int j = 0; j < nxb; ++j) {
          tma_load_4d(st + x_off + (uint32_t)j * P.x_blk_bytes, &mapX, cib * P.bi, x0 * P.stride + kw - P.pad, y0 * P.stride + kh - P.pad, img0, fb);
          if (++cib == P.ci_blocks) { cib = 0; if (++kw == P.KW) { kw = 0; ++kh; } }
        }
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
      if (++tx == P.tiles_x) { tx = 0; if (++ty == P.tiles_y) { ty = 0; ++ib; } }
    }
  } else if (warp == 1 || (warp == 3 && dualw)) {
    // MMA issuer(s): whole warp, elected lane issues; descriptors affine in the stage index
    const int iss = warp == 1 ? 0 : 1;
    int stage = 0;
    uint32_t phase = 0;
    const int ksteps = P.p_rows >> 4;
    const uint32_t xrow = (uint32_t)P.bi * 2u, drow = (uint32_t)P.bo * 2u;
    const uint64_t dydesc0 = make_desc_mn(smem_base, drow, P.dy_blk_bytes);
    const uint64_t xdesc0 = make_desc_mn(smem_base + x_off, xrow, P.x_blk_bytes);
    const uint64_t odesc0 = make_desc_mn(smem_base + P.ones_off, 32u, 0u);
    const uint32_t idesc_bias = make_idesc_mnmn(128, 16);
    const uint32_t idesc_t = make_idesc_mnmn(128, P.Cout);
    const uint32_t dk16 = (16u * drow) >> 4, xk16 = (16u * xrow) >> 4;
    const int per_n = 256 / P.bi, per_t = 128 / P.bi;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tc_fence_after();
      const uint64_t soff = (uint64_t)(((uint32_t)stage * P.stage_bytes) >> 4);
      const uint64_t dydesc = dydesc0 + soff, xdesc_s = xdesc0 + soff;
      if (elect_one()) {
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dk = (uint64_t)(dk16 * (uint32_t)k), xk = (uint64_t)(xk16 * (uint32_t)k);
          const uint32_t acc = (uint32_t)((kb | k) != 0);
          if (do_bias && iss == (dualw ? 1 : 0))   // bias gradient: D_bias[co][0..15] += dY^T (M = 128 channels) * ones (N = 16)
            umma_bf16(tmem_base + P.bias_col, dydesc + dk, odesc0 + (uint64_t)(32u * (uint32_t)k), idesc_bias, acc);
          if (!P.transposed) {
            // A = dY (M = 128 output channels), B = groups of X boxes (N <= 256 each)
            for (int j0 = 0, col = 0, gi = 0; j0 < nxb; j0 += per_n, ++gi) {
              int nb = nxb - j0 < per_n ? nxb - j0 : per_n;
              if (!dualw || (gi & 1) == iss)
                umma_bf16(tmem_base + (uint32_t)col, dydesc + dk, xdesc_s + (uint64_t)(((uint32_t)j0 * P.x_blk_bytes) >> 4) + xk,
                          make_idesc_mnmn(128, nb * P.bi), acc);
              col += nb * P.bi;
            }
          } else {
            // A = 128/bi X boxes (M = 128 rows of n'), B = dY (N = Cout)
            for (int j0 = 0, mt = 0; j0 < nxb; j0 += per_t, ++mt)
              if (!dualw || (mt & 1) == iss)
                umma_bf16(tmem_base + (uint32_t)(mt * P.Cout), xdesc_s + (uint64_t)(((uint32_t)j0 * P.x_blk_bytes) >> 4) + xk, dydesc + dk,
                          idesc_t, acc);
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&acc_bar));
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(smem_u32(&acc_bar), 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* dKg = P.dK + (int64_t)g * P.Cout * P.n_total;
    if (do_bias) {
      uint32_t r[16];
      tmem_ld16(taddr + P.bias_col, r);
      const int co = co0 + row;
      if (co < P.Cout) atomicAdd(P.dbias + (size_t)(P.dbias_gpr ? g / P.dbias_gpr : 0) * P.Cout + co, __uint_as_float(r[0]));
    }
    if (!P.transposed) {
      const int co = co0 + row;
      const int ncols = nxb * P.bi;
      const int np0 = xb0 * P.bi;
      for (int cb = 0; cb < ncols; cb += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)cb, r);
        if (co < P.Cout) {
          // 16-byte vector reductions (REDG.E.ADD.F32x4): a lane owns a row of dK, so scalar REDs were 16 separate 4-byte L2 transactions
          // per lane and block — the low-resolution layers (few pixels, large dK, split-K) spent most of their time in them
          float* dst = dKg + (int64_t)co * P.n_total + np0 + cb;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(r[i])), "f"(__uint_as_float(r[i + 1])),
                         "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3])) : "memory");
        }
      }
    } else {
      const int per = 128 / P.bi;
      const int mtiles = (nxb + per - 1) / per;
      for (int mt = 0; mt < mtiles; ++mt) {
        const int np = (xb0 + mt * per) * P.bi + row;
        const bool rvalid = (mt * per * P.bi + row) < nxb * P.bi;
        for (int cb = 0; cb < P.Cout; cb += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + (uint32_t)(mt * P.Cout + cb), r);
          if (rvalid) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cb + i < P.Cout) atomicAdd(dKg + (int64_t)(cb + i) * P.n_total + np, __uint_as_float(r[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols) : "memory");
  }
}

// pixel tile for the wgrad K-blocks: exact tiling, multiple of 16 pixels (UMMA K), at most 64
bool choose_ktile(int ipg, int H, int W, int& TN, int& TH, int& TW) {
  int best = 0;
  for (int tw = 1; tw <= W && tw <= 64; ++tw) {
    if (W % tw) continue;
    for (int th = 1; th <= H && tw * th <= 64; ++th) {
      if (H % th) continue;
      int tn = 1;
      if (tw == W && th == H)
        for (int c = 1; c <= ipg && tw * th * c <= 64; ++c)
          if (ipg % c == 0) tn = c;
      int px = tw * th * tn;
      if (px % 16) continue;
      if (px > best || (px == best && tw > TW)) { best = px; TN = tn; TH = th; TW = tw; }
    }
  }
  if (best >= 32) return true;
  // no exact tiling (10 x 12, 5 x 6 ... feature maps): full-width tiles whose last row block OVERHANGS the image — the TMA unit
  // zero-fills the rows below it in both operands, so they add nothing to the reduction (most useful pixels per tile wins)
  int best_useful = 0;
  for (int th = 1; th <= 16 && W * th <= 64; ++th) {
    if ((W * th) % 16) continue;
    const int tiles = (H + th - 1) / th;
    const int useful = W * H * 64 / (tiles * W * th);          // efficiency in 1/64
    if (useful > best_useful || (useful == best_useful && W * th > best)) { best_useful = useful; best = W * th; TN = 1; TH = th; TW = W; }
  }
  return best_useful >= 32;                                      // at least half of every tile is real pixels
}

int blk_of(int c) { return (c % 64 == 0) ? 64 : ((c % 32 == 0) ? 32 : ((c % 16 == 0) ? 16 : 0)); }
bool g_wg_attr_set = false;

// ---- split selection of k_wgrad_tma: wave quantisation ------------------------------------------------------------------------------
// One CTA is resident per SM (the operand ring takes the shared memory), CTAs are not persistent, and a launch is a few hundred CTAs of
// tens to hundreds of microseconds each: the round-2 `ncu --set full` of the largest launch of the step (SPADE sp4 gamma|beta: 320 CTAs =
// 2.16 waves of 148, the fifth X split of every chunk half as long as the others) shows sm__cycles_active avg / max = 0.70 — 30 % of
// the launch is SMs waiting for the last wave (profiles/r02_ncu_full_wgrad_tma_v39.txt).  The split therefore is chosen per launch from a
// small schedule model instead of "two waves": candidates are the balanced X-box splits the kernel already supports and every
// split-K chunk count; a CTA costs e0 + its accumulator read-out + tiles * (dY boxes + its X boxes) in units of one 64-pixel x 64-channel
// box; CTAs are dealt to the SMs in blockIdx order (x fastest), each SM takes the next CTA when it is free; the candidate with the shortest
// makespan wins, the "two waves" default is kept unless a candidate is at least 6 % shorter.  RD_B200_WGRAD_BALANCE=0 (or an explicit
// RD_B200_WGRAD_WAVES) restores the old rule.
struct WgSplit { int xb_per_cta, xsplits, chunk_tiles, chunks_pg; };
struct WgShape {
  int ptiles_pg, p_rows, bi, bo, xb_total, transposed, dy_blocks, grid_y, groups, cout, max_xb, sm_count;
  bool operator<(const WgShape& o) const {
    return std::tie(ptiles_pg, p_rows, bi, bo, xb_total, transposed, dy_blocks, grid_y, groups, cout, max_xb, sm_count) <
           std::tie(o.ptiles_pg, o.p_rows, o.bi, o.bo, o.xb_total, o.transposed, o.dy_blocks, o.grid_y, o.groups, o.cout, o.max_xb, o.sm_count);
  }
};

double wg_makespan(const WgShape& s, const WgSplit& c) {
  const double scale = s.p_rows / 64.0, e0 = 40.0;
  const int nx = c.chunks_pg * c.xsplits;
  std::vector<double> cost((size_t)nx);
  for (int x = 0; x < nx; ++x) {
    const int ch = x / c.xsplits, xi = x - ch * c.xsplits;
    const int nt = std::min(s.ptiles_pg, (ch + 1) * c.chunk_tiles) - ch * c.chunk_tiles;
    const int xm = std::min(s.xb_total, (xi + 1) * c.xb_per_cta) - xi * c.xb_per_cta;
    const int cols = s.transposed ? rd_div_up(xm * s.bi, 128) * s.cout : xm * s.bi;
    cost[(size_t)x] = e0 + cols / 12.0 + nt * scale * (s.dy_blocks * s.bo / 64.0 + xm * s.bi / 64.0);
  }
  std::priority_queue<double, std::vector<double>, std::greater<double>> sm;
  for (int i = 0; i < s.sm_count; ++i) sm.push(0.0);
  double end = 0.0;
  const int reps = s.grid_y * s.groups;
  for (int r = 0; r < reps; ++r)
    for (int x = 0; x < nx; ++x) {
      double t = sm.top() + cost[(size_t)x];
      sm.pop();
      sm.push(t);
      if (t > end) end = t;
    }
  return end;
}

WgSplit wg_split_of(const WgShape& s, int xb_per_cta, int64_t chunks) {
  WgSplit c;
  c.xb_per_cta = xb_per_cta;
  c.xsplits = rd_div_up(s.xb_total, xb_per_cta);
  const int64_t max_chunks = (s.ptiles_pg + 7) / 8;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  c.chunk_tiles = (int)((s.ptiles_pg + chunks - 1) / chunks);
  c.chunks_pg = rd_div_up(s.ptiles_pg, c.chunk_tiles);
  return c;
}

// balanced X boxes per CTA for a requested upper bound (the rule the launch always used)
int wg_balance_xb(const WgShape& s, int want) {
  if (want > s.xb_total) want = s.xb_total;
  int xsplits = rd_div_up(s.xb_total, want);
  int xb = rd_div_up(s.xb_total, xsplits);
  if (s.transposed) {                                                   // whole M tiles per CTA
    const int per = 128 / s.bi;
    xb = rd_div_up(xb, per) * per;
  }
  return xb;
}

WgSplit wg_choose_split(const WgShape& s) {
  static const int waves_env = getenv("RD_B200_WGRAD_WAVES") ? atoi(getenv("RD_B200_WGRAD_WAVES")) : 0;
  static const bool balance = !(getenv("RD_B200_WGRAD_BALANCE") && atoi(getenv("RD_B200_WGRAD_BALANCE")) == 0);
  const int xb_def = wg_balance_xb(s, s.max_xb);
  const int waves = waves_env > 0 ? waves_env : 2;
  const int64_t other = (int64_t)rd_div_up(s.xb_total, xb_def) * s.grid_y * s.groups;
  const WgSplit def = wg_split_of(s, xb_def, ((int64_t)s.sm_count * waves + other - 1) / other);
  if (!balance || waves_env > 0) return def;
  static std::mutex mu;
  static std::map<WgShape, WgSplit> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(s);
  if (it != cache.end()) return it->second;
  WgSplit best = def;
  const double t_def = wg_makespan(s, def);
  double t_best = t_def;
  const int step = s.transposed ? 128 / s.bi : 1;
  // The model has no variance: measured, ONE wave of long CTAs is slower than two (RD_B200_WGRAD_WAVES=1: +0.5 ms per step — SMs do not
  // run at equal speed and a single wave cannot rebalance), so a candidate never has fewer CTAs than the default unless it still has ~2 waves
  const int64_t n_def = (int64_t)def.chunks_pg * def.xsplits * s.grid_y * s.groups;
  const int64_t n_floor = std::min<int64_t>(n_def, 2 * s.sm_count - s.sm_count / 8);
  int last_xb = -1;
  for (int want = s.transposed ? step : 2; want <= s.max_xb; want += step) {      // (one-box CTAs, N = 64 MMAs only: never the better choice in the model)
    const int xb = wg_balance_xb(s, want);
    if (xb == last_xb || xb > s.max_xb) continue;
    last_xb = xb;
    const int64_t per_chunk = (int64_t)rd_div_up(s.xb_total, xb) * s.grid_y * s.groups;
    int last_tiles = -1;
    for (int64_t chunks = 1; chunks <= (s.ptiles_pg + 7) / 8 && chunks * per_chunk <= (int64_t)s.sm_count * 6; ++chunks) {
      const WgSplit c = wg_split_of(s, xb, chunks);
      if (c.chunk_tiles == last_tiles) continue;
      last_tiles = c.chunk_tiles;
      if ((int64_t)c.chunks_pg * per_chunk < n_floor) continue;
      const double t = wg_makespan(s, c);
      if (t < t_best * 0.995) { t_best = t; best = c; }
    }
  }
  if (t_best > 0.94 * t_def) { best = def; t_best = t_def; }
  if (getenv("RD_B200_WGRAD_TRACE"))
    fprintf(stderr, "wgrad_tma split: tiles/group %d x %d px, X boxes %d x %d ch, Cout %d, groups %d: default %d boxes/CTA x %d chunks (%d CTAs) -> "
            "%d boxes/CTA x %d chunks (%d CTAs), modelled makespan %.2f of the default\n", s.ptiles_pg, s.p_rows, s.xb_total, s.bi, s.cout, s.groups,
            def.xb_per_cta, def.chunks_pg, def.chunks_pg * def.xsplits * s.grid_y * s.groups, best.xb_per_cta, best.chunks_pg,
            best.chunks_pg * best.xsplits * s.grid_y * s.groups, t_best / t_def);
  cache[s] = best;
  return best;
}

// geometry of a k_wgrad_tma launch that does not depend on the split: pixel tiles, channel boxes, accumulator form.
// grid_y: CTAs along Cout; max_x_ch: X channels (n') one CTA can accumulate
bool wg_geometry(const rd_conv_desc* d, WgTmaParams& P, int& grid_y, int& max_x_ch) {
  P.H = d->oh; P.W = d->ow; P.Cin = d->cin; P.Cout = d->cout; P.KW = d->kw; P.pad = d->pad; P.stride = d->stride;
  P.ipg = d->n / d->groups;
  P.TW = 0;
  if (!choose_ktile(P.ipg, P.H, P.W, P.TN, P.TH, P.TW)) return false;
  P.p_rows = P.TW * P.TH * P.TN;
  P.tiles_x = P.W / P.TW; P.tiles_y = rd_div_up(P.H, P.TH); P.img_blocks_pg = P.ipg / P.TN;      // the last row block may overhang (choose_ktile)
  P.ptiles_pg = P.img_blocks_pg * P.tiles_y * P.tiles_x;
  P.bi = blk_of(P.Cin); P.bo = blk_of(P.Cout);
  P.ci_blocks = P.Cin / P.bi;
  const int taps = d->kh * d->kw;
  P.xb_total = taps * P.ci_blocks;
  P.n_total = taps * P.Cin;
  P.transposed = P.Cout < 128 ? 1 : 0;
  if (!P.transposed) {
    P.dy_blocks = 2;     // 128 output channels = two 64-channel boxes
    grid_y = P.Cout / 128;
    max_x_ch = 256;
  } else {
    P.dy_blocks = P.Cout / P.bo;
    grid_y = 1;
    // accumulators: (#M tiles) * Cout columns <= 512 ; smem: keep a stage under ~60 KB
    int max_mt = 512 / P.Cout;
    if (max_mt > 3) max_mt = 3;
    max_x_ch = max_mt * 128;
  }
  return true;
}

WgShape wg_shape_of(const rd_conv_desc* d, const WgTmaParams& P, int grid_y, int max_x_ch, int sm_count) {
  return WgShape{P.ptiles_pg, P.p_rows, P.bi, P.bo, P.xb_total, P.transposed, P.dy_blocks, grid_y, d->groups, P.Cout, max_x_ch / P.bi, sm_count};
}

}  // namespace

int rd_wgrad_tma_supported(const rd_conv_desc* d) {
  if (d->dtype != RD_BF16) return 0;
  static const bool s2 = getenv("RD_B200_NO_TMA_S2") == nullptr;
  const bool strided = s2 && d->stride == 2 && d->kh == d->kw && (d->kh == 3 || d->kh == 4) && d->pad == 1 && d->oh * 2 == d->h && d->ow * 2 == d->w;
  if (!strided) {
    if (d->stride != 1 || d->kh != d->kw || (d->kh != 1 && d->kh != 3) || d->pad != (d->kh - 1) / 2) return 0;
    if (d->oh != d->h || d->ow != d->w) return 0;
  }
  if (!blk_of(d->cin) || !blk_of(d->cout)) return 0;
  if (d->cout >= 128 && d->cout % 128) return 0;
  if (d->cout < 128 && (d->cout % 16 || d->cout > 64 * 2)) return 0;
  int TN, TH, TW = 0;
  if (!choose_ktile(d->n / d->groups, d->oh, d->ow, TN, TH, TW)) return 0;
  if (!get_encode()) return 0;
  return 1;
}

// Host helper of the ABI (no device work, no driver): the split the launch below would use on a device with sm_count SMs
extern "C" int rd_wgrad_tma_plan(const rd_conv_desc* d, int sm_count, int* out) {
  if (!d || !out || sm_count < 1 || d->dtype != RD_BF16 || d->groups < 1 || d->n % d->groups) return 0;
  if (!blk_of(d->cin) || !blk_of(d->cout)) return 0;
  if (d->cout >= 128 && d->cout % 128) return 0;
  if (d->cout < 128 && (d->cout % 16 || d->cout > 64 * 2)) return 0;
  WgTmaParams P;
  int grid_y, max_x_ch;
  if (!wg_geometry(d, P, grid_y, max_x_ch)) return 0;
  const WgShape shp = wg_shape_of(d, P, grid_y, max_x_ch, sm_count);
  const WgSplit c = wg_choose_split(shp);
  const int xb_def = wg_balance_xb(shp, shp.max_xb);
  const int64_t other = (int64_t)rd_div_up(shp.xb_total, xb_def) * grid_y * d->groups;
  const WgSplit def = wg_split_of(shp, xb_def, ((int64_t)sm_count * 2 + other - 1) / other);
  out[0] = P.xb_total; out[1] = c.xb_per_cta; out[2] = c.xsplits; out[3] = P.ptiles_pg; out[4] = c.chunk_tiles; out[5] = c.chunks_pg;
  out[6] = c.chunks_pg * c.xsplits * grid_y * d->groups;
  out[7] = def.chunks_pg * def.xsplits * grid_y * d->groups;
  return 1;
}

int rd_wgrad_tma_launch(rd_ctx* ctx, const rd_conv_desc* d, const void* x, const void* dy, float* dK, float* dbias,
                        cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  if (!enc) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  WgTmaParams P;
  P.dK = dK;
  P.dbias = dbias;
  P.dbias_gpr = d->bias_groups > 1 ? d->groups / d->bias_groups : 0;
  int grid_y, max_x_ch;
  if (!wg_geometry(d, P, grid_y, max_x_ch)) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "wgrad_tma: no pixel tiling");
  // X boxes per CTA (accumulator columns) and split-K chunks: see wg_choose_split
  const WgSplit split = wg_choose_split(wg_shape_of(d, P, grid_y, max_x_ch, ctx->sm_count));
  P.xb_per_cta = split.xb_per_cta;
  P.xsplits = split.xsplits;
  P.dy_blk_bytes = ((uint32_t)P.p_rows * P.bo * 2u + 1023u) & ~1023u;
  P.x_blk_bytes = ((uint32_t)P.p_rows * P.bi * 2u + 1023u) & ~1023u;
  P.tx_dy = (uint32_t)P.p_rows * P.bo * 2u;
  P.tx_x = (uint32_t)P.p_rows * P.bi * 2u;
  P.stage_bytes = (uint32_t)P.dy_blocks * P.dy_blk_bytes + (uint32_t)P.xb_per_cta * P.x_blk_bytes;
  int stages = (int)((186u * 1024u) / P.stage_bytes);
  if (stages > kWgTmaMaxStages) stages = kWgTmaMaxStages;
  if (stages < 2) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "wgrad_tma: stage too large");
  P.stages = stages;
  { const char* e_du = getenv("RD_B200_TMA_DUAL"); P.dual = (e_du && atoi(e_du) == 0) ? 0 : 1; }
  uint32_t need_cols = P.transposed ? (uint32_t)(rd_div_up(P.xb_per_cta * P.bi, 128) * P.Cout) : (uint32_t)(P.xb_per_cta * P.bi);
  P.bias_col = need_cols;
  if (dbias) need_cols += 16;
  uint32_t cols = 32;
  while (cols < need_cols) cols <<= 1;
  if (cols > 512) RD_FAIL(ctx, RD_ERR_UNSUPPORTED, "wgrad_tma: accumulator does not fit TMEM");
  P.tmem_cols = cols;
  // split-K: CTAs are not persistent (1 resident per SM), so the pixel tiles of a group are cut into chunks of at least 8 tiles
  P.chunk_tiles = split.chunk_tiles;
  P.chunks_pg = split.chunks_pg;

  alignas(64) CUtensorMap mapX, mapDY;
  auto sw_of = [](int b) { return b == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (b == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); };
  {
    const cuuint32_t sst = (cuuint32_t)P.stride;
    cuuint64_t dims[4] = {(cuuint64_t)P.Cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)P.Cin * 2, (cuuint64_t)d->w * P.Cin * 2, (cuuint64_t)d->h * d->w * P.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)P.bi, (cuuint32_t)((P.TW - 1) * P.stride + 1), (cuuint32_t)((P.TH - 1) * P.stride + 1), (cuuint32_t)P.TN};
    cuuint32_t es[4] = {1, sst, sst, 1};
    CUresult r = enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw_of(P.bi), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)P.Cout, (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)P.Cout * 2, (cuuint64_t)P.W * P.Cout * 2, (cuuint64_t)P.H * P.W * P.Cout * 2};
    cuuint32_t box[4] = {(cuuint32_t)P.bo, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&mapDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw_of(P.bo), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RD_FAIL(ctx, RD_ERR_CUDA, "cuTensorMapEncodeTiled(dY) failed: %d", (int)r);
  }
  // slack after the last stage: a partial last M tile (transposed) reads up to 128/bi boxes; then the ones block
  P.ones_off = (uint32_t)stages * P.stage_bytes + 16u * 1024u;
  size_t smem = (size_t)stages * P.stage_bytes + 1024 + 16 * 1024 + 4 * 1024;
  if (!g_wg_attr_set) {
    RD_CUDA(ctx, cudaFuncSetAttribute(k_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    g_wg_attr_set = true;
  }
  dim3 grid(P.chunks_pg * P.xsplits, grid_y, d->groups);
  k_wgrad_tma<<<grid, kTmaThreads, smem, st>>>(mapX, mapDY, P);
  RD_CHECK_LAUNCH(ctx, "wgrad_tma");
  return RD_OK;
}
