/*
 * rd_b200.h — C ABI of the B200-native hot path of representation-disentanglement.
 *
 * The reference (ouyangjiahong/representation-disentanglement) is pure PyTorch and exposes no
 * plugin / FFI interface; its hot path reaches cuDNN / cuBLAS / ATen through `torch` library
 * calls.  Each entry point below replaces one family of those ATen call sites (cited as
 * reference file:line under src/) with a hand-written sm_100a kernel.  The Python side
 * (rd_b200/lib.py, ctypes) binds exactly these symbols; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rd_status; rd_last_error() gives text.
 *   - all tensor arguments are raw DEVICE pointers owned by the caller (PyTorch's allocator);
 *     the library never allocates or frees device memory inside these calls, and never
 *     synchronises with the host: all work is ordered on `stream` (CUDA-graph capturable).
 *   - activations are NHWC ("channels last", C innermost) in `dtype` (RD_F32 or RD_BF16);
 *     statistics, losses, packed-weight gradients and parameters are fp32.
 *   - "groups": the leading image dimension is G*Ng; group g = n / Ng selects the g-th
 *     packed weight set / norm parameter set (one CondConv kernel per modality type,
 *     reference src/model.py:2108-2117 with inputs_type constant over the batch).
 */
#ifndef RD_B200_H_
#define RD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rd_ctx rd_ctx;
typedef void* rd_stream; /* cudaStream_t */

enum rd_dtype { RD_F32 = 0, RD_BF16 = 1 };
enum rd_status {
  RD_OK = 0, RD_ERR_ARG = -1, RD_ERR_CUDA = -2, RD_ERR_UNSUPPORTED = -3, RD_ERR_NO_DEVICE = -4
};
/* RD_ALGO_HALO forces the halo-tile tcgen05 kernel (3x3 stride-1, weights resident in shared memory); AUTO picks it
 * when the launch has enough tiles per SM. */
enum rd_conv_algo { RD_ALGO_AUTO = 0, RD_ALGO_DIRECT = 1, RD_ALGO_TCGEN05 = 2, RD_ALGO_HALO = 3 };
enum rd_act { RD_ACT_NONE = 0, RD_ACT_LRELU = 1 };

/* ---- context ---------------------------------------------------------------------------- */
int rd_abi_version(void);
int rd_ctx_create(rd_ctx** out, int device);          /* sets device limits / smem attributes   */
int rd_ctx_destroy(rd_ctx* ctx);
const char* rd_last_error(rd_ctx* ctx);
/* number of kernels this library has launched through `ctx` (bench.py's gpu_launches) */
int64_t rd_launch_count(rd_ctx* ctx);
/* 1 if the tcgen05 path was used by the last rd_conv2d_* call, else 0 */
int rd_last_conv_algo(rd_ctx* ctx);

/* ---- layout / cast ---------------------------------------------------------------------- */
/* src NCHW fp32 (n, c_total, h, w) channels [c0, c0+c) -> dst NHWC (n, h, w, c) in dtype.
 * replaces the per-modality slicing + implicit layout of src/main_missing.py:165-168 */
int rd_nchw_to_nhwc(rd_ctx*, const float* src, void* dst, int n, int c_total, int c0, int c, int h, int w,
                    int dtype, rd_stream);
/* all `mods` contrasts at once: src (n, mods * c, h, w) fp32 -> dst (mods * n, h, w, c), contrast-major (the stack the batched encoders
 * read; src/main_missing.py:165-168 slices the contrasts in a Python loop) */
int rd_stack_modalities(rd_ctx*, const float* src, void* dst, int n, int mods, int c, int c_pad /* dst channels, [c, c_pad) zero */, int h, int w, int dtype, rd_stream);
int rd_nhwc_to_nchw(rd_ctx*, const void* src, float* dst, int n, int c, int h, int w, int dtype, rd_stream);
int rd_cast(rd_ctx*, const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, rd_stream);
/* out[n, :, 0:ca] = a, out[n, :, ca:ca+cb] = b  (torch.cat(dim=1), src/model.py:2192) and its inverse */
int rd_concat_channels(rd_ctx*, const void* a, const void* b, void* out, int64_t pixels, int ca, int cb,
                       int dtype, rd_stream);
int rd_split_channels(rd_ctx*, const void* in, void* a, void* b, int64_t pixels, int ca, int cb,
                      int dtype, rd_stream);
/* out[p, 0:c] = in[p, 0:c], out[p, c:c_pad] = 0 (channel padding to a multiple of 8 for the tensor-core gathers) */
int rd_pad_channels(rd_ctx*, const void* in, void* out, int64_t pixels, int c, int c_pad, int dtype, rd_stream);
/* fan-out of s_i / z_j over the (i, j) decodes (src/model.py:3187-3224) in one launch: dst block k = src block
 * index[k] (a block = `block_pixels` consecutive pixels = the B images of one modality), channels [c, c_pad) of dst are
 * zero (the tensor-core kernels read 16-channel vectors).  index: HOST int32[nb], nb <= 32 (copied into the launch).
 * bwd: dsrc block i = sum over k with index[k] == i of dout block k (fp32 accumulation), padding channels dropped;
 * source blocks that are not referenced receive zeros. */
int rd_gather_blocks_fwd(rd_ctx*, const void* src, void* dst, const int32_t* index, int nb, int64_t block_pixels,
                         int c, int c_pad, int dtype, rd_stream);
int rd_gather_blocks_bwd(rd_ctx*, const void* dout, void* dsrc, const int32_t* index, int nb, int nsrc,
                         int64_t block_pixels, int c, int c_pad, int dtype, rd_stream);
/* dst block d (block_pixels pixels, c_pad channels, the padding zero) = block sblk[d] of source a (sel[d] == 0) or b (sel[d] == 1): the
 * gradients of the self- and cross-reconstruction stacks (src/model.py:3187-3224) scattered into the zero-padded dY of the decoder's
 * last convolution in one pass */
int rd_scatter_blocks2(rd_ctx*, const void* a, const void* b, void* dst, const int32_t* sel_host, const int32_t* sblk_host, int nb,
                       int64_t block_pixels, int c, int c_pad, int dtype, rd_stream);
/* y = x + a (grad accumulation of fan-out tensors), y may alias x */
/* y = xs[0] + ... + xs[k-1] (k <= 8 device pointers passed in a HOST array, 16-byte aligned tensors, fp32 accumulation): the summed
 * gradient of a tensor with several consumers in one pass */
int rd_add_n(rd_ctx*, const void* const* xs, int k, void* y, int64_t n, int dtype, rd_stream);
int rd_add(rd_ctx*, const void* x, const void* a, void* y, int64_t n, int dtype, rd_stream);

/* ---- CondConv expert mixing (src/model.py:2065-2113) ------------------------------------- */
/* r[g,e] = sigmoid(fc_w[e]*types[g] + fc_b[e]); K[g] = sum_e r[g,e] W[e].
 * W (E,O,I,kh,kw) fp32.  i_pad >= I is the channel count of the (zero-padded) activation tensor: the tensor-core
 * kernels need 16-byte channel vectors, so 4- / 7-channel inputs are stored with 8 channels.  Outputs (either NULL):
 *   packed  [G][o_total ][kh*kw][i_pad]  rows [o_off, o_off+O), columns >= I written as 0   (forward / wgrad, "OHWI")
 *   packedT [G][i_pad][kh*kw][oT_total]  cols [o_off, o_off+O), rows    >= I written as 0   (dgrad, "IHWO")
 * fc_w == NULL means a plain nn.Conv2d weight (E must be 1, r = 1).  `types` is a HOST array. */
int rd_condconv_mix_fwd(rd_ctx*, const float* W, const float* fc_w, const float* fc_b, const float* types,
                        int G, int E, int O, int I, int i_pad, int kh, int kw, int o_total, int oT_total, int o_off,
                        void* packed, void* packedT, float* r_out /* [G][E] or NULL */, int dtype, rd_stream);
/* dK [G][o_total][kh*kw][i_pad] fp32 (rows [o_off,o_off+O), columns < I used) -> dW (E,O,I,kh,kw) +=, dfc_w[e] +=, dfc_b[e] += */
int rd_condconv_mix_bwd(rd_ctx*, const float* dK, const float* W, const float* fc_w, const float* fc_b,
                        const float* types, int G, int E, int O, int I, int i_pad, int kh, int kw, int o_total, int o_off,
                        float* dW, float* dfc_w, float* dfc_b, rd_stream);

/* The same backward for MANY CondConv heads in one launch (the per-head launches are latency bound: ~20 us each,
 * ~85 per step).  `jobs` is a DEVICE array of rd_mix_job (the caller fills a host copy and uploads it); job j owns
 * blocks [block_begin[j], block_begin[j+1]) of the grid, block_begin ascending, total_blocks = the last job's end. */
typedef struct rd_mix_job {
  const float* dK; const float* W; const float* fc_w; const float* fc_b;
  float* dW; float* dfc_w; float* dfc_b;
  float types[16];
  int32_t G, E, O, I, i_pad, taps, o_total, o_off;
  int32_t block_begin, blocks;
  const float* bias_src; float* bias_dst;      /* optional: bias_dst[0..bias_n) += bias_src[..] (the head's slice of a fused launch's */
  int32_t bias_n, _pad;                        /* bias-gradient row added into bias.grad by the same launch) */
} rd_mix_job;
int rd_mix_job_blocks(int O, int I, int taps);          /* grid blocks one job needs (host helper) */
int rd_condconv_mix_bwd_batched(rd_ctx*, const rd_mix_job* jobs_dev, int njobs, int total_blocks, int max_groups /* largest G of the jobs (0 = unknown) */, rd_stream);

/* Forward mixing of MANY CondConv heads in one launch (the experts only change at the optimizer step, so a trainer mixes every
 * layer of the iteration up front instead of ~90 per-head launches): same arithmetic as rd_condconv_mix_fwd per job, plus an
 * optional copy of the head's bias into its slot of the fused launch's bias row.  Both packed layouts are always written. */
typedef struct rd_mixf_job {
  const float* W; const float* fc_w; const float* fc_b;
  void* packed; void* packedT;
  const float* bias_src; float* bias_dst;
  float types[16];
  int32_t G, E, O, I, i_pad, taps, o_total, oT_total, o_off, bias_n;
  int32_t block_begin, blocks;
} rd_mixf_job;
int rd_mixf_job_blocks(int O, int i_pad, int taps);      /* grid blocks one job needs (host helper) */
int rd_condconv_mix_fwd_batched(rd_ctx*, const rd_mixf_job* jobs_dev, int njobs, int total_blocks, int dtype, rd_stream);

/* ---- convolution (src/model.py:2104 F.conv2d and its autograd) ---------------------------- */
typedef struct rd_conv_desc {
  int n, h, w, cin;          /* input  NHWC                                   */
  int oh, ow, cout;          /* output NHWC                                   */
  int kh, kw, stride, pad;
  int groups;                /* G weight sets; images per group = n / groups  */
  int dtype;                 /* rd_dtype of x / y / packed weights            */
  int act;                   /* rd_act fused on the forward output            */
  float act_slope;           /* LeakyReLU slope (0.2, src/model.py:2227)      */
  int algo;                  /* rd_conv_algo                                  */
  int bias_groups;           /* 0 / 1: bias[cout] shared by all groups (one CondConv module, src/model.py:2104);
                                R > 1 (R divides groups): bias[R][cout], weight group g uses row g / (groups / R) —
                                several modules batched into one launch (the per-modality decoder halves
                                input_decoder_list[i]), each with its own bias and groups / R modality types        */
} rd_conv_desc;
/* y = act(conv(x, packed[g]) + bias); bias fp32 [cout] (or [bias_groups][cout]) or NULL (src/model.py:2104) */
int rd_conv2d_fwd(rd_ctx*, const rd_conv_desc*, const void* x, const void* packed, const float* bias,
                  void* y, rd_stream);
/* SPADE block (src/model.py:2444-2452): the gamma|beta convolution (cout = 2C: columns [0, C) gamma, [C, 2C) beta) with the modulation
 * in its epilogue — writes gamma [N,H,W,C] (kept for the backward) and mix = (z - mean) * invstd * (1 + gamma) + beta [N,H,W,C]; the
 * [N,H,W,2C] gamma|beta tensor and the rd_spade_modulate_fwd pass do not exist.  mean / invstd: InstanceNorm statistics of z, fp32
 * [N][C] (rd_norm_stats with one group per image).  rd_conv2d_fwd_spade_supported: 1 when the shape runs on the halo kernel
 * (bf16, 3x3 stride 1, weights resident in shared memory, C a multiple of 32 and <= 128), else 0 — the caller then uses
 * rd_conv2d_fwd + rd_spade_modulate_fwd. */
int rd_conv2d_fwd_spade_supported(rd_ctx*, const rd_conv_desc*);
int rd_conv2d_fwd_spade(rd_ctx*, const rd_conv_desc*, const void* x, const void* packed, const float* bias, const void* z,
                        const float* mean, const float* invstd, void* gamma, void* mix, rd_stream);
/* dx = conv_transpose(dy, packedT[g])   (input gradient) */
int rd_conv2d_dgrad(rd_ctx*, const rd_conv_desc*, const void* dy, const void* packedT, void* dx, rd_stream);
/* dK[g] (fp32, OHWI, zeroed by the call) = sum over the group's images; dbias[cout] (or [bias_groups][cout]) += sum dy (may be NULL) */
int rd_conv2d_wgrad(rd_ctx*, const rd_conv_desc*, const void* x, const void* dy, float* dK, float* dbias,
                    rd_stream);
/* Host helper (no device work, callable without a GPU): the launch split of the TMA weight-gradient kernel for `d` on a device with
 * sm_count SMs.  CTAs are one per SM and not persistent, so the accumulator columns per CTA and the split-K chunk count are chosen per
 * launch against wave quantisation.  out[0..7] = {X boxes in total, X boxes per CTA, X splits, pixel tiles per weight group,
 * tiles per split-K chunk, chunks per group, CTAs of the launch, CTAs of the plain "two waves" rule}.  Returns 1, or 0 when the shape
 * does not run on that kernel.  (The backward of CondConv2d's F.conv2d, src/model.py:2104.) */
int rd_wgrad_tma_plan(const rd_conv_desc* d, int sm_count, int* out);

/* ---- composed decoder tail (src/model.py:2606-2612): SPADEBlockNew sp6 `out` (3x3, Cin -> OA) followed directly by the CondConv 1x1
 * `out` (OA -> OB) is ONE 3x3 convolution with per-group weights W_eff[g] = pB[g] . pA[g], b_eff[g] = pB[g] . bA[m] + bB[m]
 * (m = g / (G / modules)).  pA fp32 [G][OA][taps][Cin], pB fp32 [G][OB][OA] (the mixed experts, rd_condconv_mix_fwd), bA [modules][OA]
 * / bB [modules][OB] or NULL.  Writes the convolution's packed weights [G][OB][taps][Cin] and transposed [G][Cin][taps][o_pad]
 * (columns >= OB zero) in `dtype`, and b_eff fp32 [G][OB].  Replaces torch.einsum + casts of the first implementation. */
int rd_compose_tail_fwd(rd_ctx*, const float* pA, const float* pB, const float* bA, const float* bB, int G, int modules,
                        int OA, int OB, int taps, int Cin, int o_pad, int dtype, void* packed, void* packedT, float* b_eff,
                        rd_stream);
/* chain rule through the composition: dK fp32 [G][o_pad][taps][Cin] and db fp32 [G][o_pad] of the composed convolution ->
 * dpA [G][OA][taps][Cin], dpB [G][OB][OA] (both overwritten), dbA [modules][OA] += , dbB [modules][OB] += (NULL = skip). */
int rd_compose_tail_bwd(rd_ctx*, const float* dK, const float* db, const float* pA, const float* pB, const float* bA,
                        int G, int modules, int OA, int OB, int taps, int Cin, int o_pad, float* dpA, float* dpB,
                        float* dbA, float* dbB, rd_stream);

/* ---- runtime services: CUDA-graph capture of a launch sequence, NCCL gradient averaging --------------
 * rd_graph_*: everything launched on `stream` between begin and end (rd_* entry points never synchronise or allocate) becomes one
 * CUDA graph; rd_graph_launch replays it.  This is what rd_b200.trainer does through torch.cuda.CUDAGraph (reference loop body
 * src/main_missing.py:165-284 = one graph launch per iteration). */
typedef struct rd_graph rd_graph;
/* zero `bytes` bytes at p (cudaMemsetAsync: a memset node, not a kernel) */
int rd_zero(rd_ctx*, void* p, int64_t bytes, rd_stream);
int rd_graph_begin(rd_ctx*, rd_stream stream);
int rd_graph_end(rd_ctx*, rd_stream stream, rd_graph** out);
int rd_graph_launch(rd_ctx*, rd_graph*, rd_stream stream);
int rd_graph_node_count(rd_ctx*, rd_graph*, int64_t* kernel_nodes, int64_t* total_nodes);
int rd_graph_destroy(rd_ctx*, rd_graph*);
/* rd_ddp_*: one process per GPU.  Rank 0 calls rd_ddp_unique_id and hands the 128 bytes to the other ranks by the caller's own means;
 * every rank calls rd_ddp_init; rd_ddp_bucket_allreduce averages (or sums) a contiguous fp32 range of the flat gradient buffer in
 * place, stream-ordered and graph-capturable; rd_ddp_broadcast copies rank `root`'s bytes to all ranks (parameters, Adam state).
 * The reference has no data parallelism: this is "what DDP would do" to src/main_missing.py:272-284.
 * rd_ddp_available: 1 when libnccl.so.2 could be loaded (version = NCCL_VERSION_CODE), else 0. */
int rd_ddp_available(rd_ctx*, int* version);
int rd_ddp_unique_id(rd_ctx*, void* id128);
int rd_ddp_init(rd_ctx*, int world, int rank, const void* id128);
int rd_ddp_bucket_allreduce(rd_ctx*, float* grad, int64_t n, int average, rd_stream);
int rd_ddp_broadcast(rd_ctx*, void* buf, int64_t bytes, int root, rd_stream);
int rd_ddp_finalize(rd_ctx*);

/* ---- normalisation: BatchNorm2d train/eval (src/model.py:2132,2179) and InstanceNorm2d (:2431) -- */
/* per (group, channel) mean / inverse std over the group's images and all pixels (biased variance,
 * eps inside the sqrt).  InstanceNorm = one group per image.  partial: workspace fp32
 * [2 * G * C * (rd_norm_partial_chunks(pixels_per_group, C) + 1)].  If running_mean != NULL the G group statistics are
 * folded into the running buffers sequentially in group order with `momentum` and the unbiased
 * variance, and *num_batches_tracked += G (torch.nn.BatchNorm2d semantics, one module called G times). */
int rd_norm_partial_chunks(int64_t pixels_per_group, int C);
int rd_norm_stats(rd_ctx*, const void* x, int G, int64_t pixels_per_group, int C, int dtype, float eps,
                  float* partial, float* mean, float* invstd,
                  float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                  rd_stream);
/* invstd = 1/sqrt(var+eps), mean = running_mean broadcast over G (eval-mode BN) */
int rd_norm_eval_stats(rd_ctx*, const float* running_mean, const float* running_var, int G, int C, float eps,
                       float* mean, float* invstd, rd_stream);
/* y = (x-mean)*invstd*weight + bias  (weight/bias NULL = no affine) */
int rd_norm_apply(rd_ctx*, const void* x, const float* mean, const float* invstd, const float* weight,
                  const float* bias, void* y, int G, int64_t pixels_per_group, int C, int dtype, rd_stream);
/* backward of train-mode normalisation: dx, dweight[C] +=, dbias[C] += (NULL = no affine) */
int rd_norm_bwd(rd_ctx*, const void* x, const void* dy, const float* mean, const float* invstd,
                const float* weight, void* dx, float* dweight, float* dbias, float* partial,
                int G, int64_t pixels_per_group, int C, int dtype, rd_stream);

/* ---- SPADE modulation (src/model.py:2438-2452): mix = IN(z)*(1+gamma)+beta, gb = [gamma | beta] (2C) -- */
int rd_spade_modulate_fwd(rd_ctx*, const void* z, const float* mean, const float* invstd, const void* gb,
                          void* mix, int N, int64_t hw, int C, int dtype, rd_stream);
/* dgb = [dmix*zhat | dmix]; dz = IN-backward of dmix*(1+gamma).  rd_spade_modulate_bwd reads gamma from gb [.., 2C];
 * rd_spade_modulate_bwd_g takes the gamma tensor [.., C] that rd_conv2d_fwd_spade saved. */
int rd_spade_modulate_bwd_g(rd_ctx*, const void* z, const float* mean, const float* invstd, const void* gamma,
                            const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                            int dtype, rd_stream);
int rd_spade_modulate_bwd(rd_ctx*, const void* z, const float* mean, const float* invstd, const void* gb,
                          const void* dmix, void* dz, void* dgb, float* partial, int N, int64_t hw, int C,
                          int dtype, rd_stream);
/* floats of `partial` the two calls above need (per-chunk partial sums + per-image sums).  With 16-byte aligned rows and C / (8 bf16 |
 * 4 fp32) a power of two <= 32 they run as ONE pass (k_spade_bwd_fused: z, gamma, dmix read once and kept in registers across a
 * per-image counter barrier); RD_B200_SPADE_BWD_FUSED=0 selects the two-pass form. */
int64_t rd_spade_bwd_workspace(int N, int64_t hw, int C, int dtype);

/* ---- bilinear resize (src/model.py:2175 align_corners=True x2; :2432,:2501 align_corners=False) ---- */
int rd_bilinear_fwd(rd_ctx*, const void* x, void* y, int n, int h, int w, int c, int oh, int ow,
                    int align_corners, int dtype, rd_stream);
int rd_bilinear_bwd(rd_ctx*, const void* dy, void* dx, int n, int h, int w, int c, int oh, int ow,
                    int align_corners, int dtype, rd_stream);

/* ---- activations ------------------------------------------------------------------------- */
int rd_lrelu_fwd(rd_ctx*, const void* x, void* y, int64_t n, float slope, int dtype, rd_stream);
/* dx = dy * (y > 0 ? 1 : slope), y = forward OUTPUT */
int rd_lrelu_bwd(rd_ctx*, const void* dy, const void* y, void* dx, int64_t n, float slope, int dtype, rd_stream);
/* F.softplus (beta 1, threshold 20): input_output_act / target_output_act / ana_dec_act 'softplus'
 * (src/main_missing.py:75-86, src/model.py:2631, 3145-3146); bwd takes the forward INPUT x */
int rd_softplus_fwd(rd_ctx*, const void* x, void* y, int64_t n, int dtype, rd_stream);
int rd_softplus_bwd(rd_ctx*, const void* dy, const void* x, void* dx, int64_t n, int dtype, rd_stream);
/* masked softmax of compute_anatomy_encoding (src/model.py:3149-3153):
 * p = softmax_c([100*mask_img, s])[1:]  (mask_img NULL -> plain softmax over c) */
int rd_masked_softmax_fwd(rd_ctx*, const void* s, const float* mask_img, int64_t mask_pixels /* mask index = pixel % mask_pixels */,
                          void* p, int64_t pixels, int C, int dtype, rd_stream);
int rd_masked_softmax_bwd(rd_ctx*, const void* p, const void* dp, void* ds, int64_t pixels, int C, int dtype,
                          rd_stream);

/* attention gate of the U+SA output decoder (SpatialAttentionLayer, src/model.py:1316-1327) */
int rd_add_relu_fwd(rd_ctx*, const void* a, const void* b, void* y, int64_t n, int dtype, rd_stream);   /* relu(a+b) */
int rd_relu_bwd(rd_ctx*, const void* dy, const void* y, void* dx, int64_t n, int dtype, rd_stream);
int rd_sigmoid_fwd(rd_ctx*, const void* x, void* y, int64_t n, int dtype, rd_stream);
int rd_sigmoid_bwd(rd_ctx*, const void* dy, const void* y, void* dx, int64_t n, int dtype, rd_stream);
/* y[p,c] = alpha[p] * x[p,c] ; backward: dx = alpha*dy, dalpha[p] = sum_c dy*x */
int rd_mul_bcast_fwd(rd_ctx*, const void* alpha, const void* x, void* y, int64_t pixels, int C, float off, int dtype, rd_stream);   /* (off + alpha[p]) * x[p, c] */
int rd_mul_bcast_bwd(rd_ctx*, const void* alpha, const void* x, const void* dy, void* dx, void* dalpha,
                     int64_t pixels, int C, float off, int dtype, rd_stream);
/* output-decoder variants U+SA+CA / U+SSA+CA: ChannelAttentionLayer (src/model.py:1417-1433): y = (1 + a[n, c]) * x with a fp32 [N][C];
 * backward dx = (1 + a) dy, da[n, c] = sum_p dy * x.  rd_chan_bcast: dx[n, p, c] = v[n, c] * scale (backward of the global average pool
 * torch.mean(x, (2, 3)); the forward is rd_norm_stats with one group per image).  rd_flip_absdiff: |g - flip_H(g)| of
 * SymmetryGateResidualSpatialAttentionLayer (:1408-1409) and its backward sign(g - flip g) * (dout + flip dout). */
int rd_chan_scale_fwd(rd_ctx*, const void* x, const float* a, void* y, int N, int64_t hw, int C, int dtype, rd_stream);
int rd_chan_scale_bwd(rd_ctx*, const void* x, const float* a, const void* dy, void* dx, float* da, int N, int64_t hw, int C,
                      int dtype, rd_stream);
int rd_chan_bcast(rd_ctx*, const float* v, void* dx, int N, int64_t hw, int C, float scale, int dtype, rd_stream);
int rd_flip_absdiff_fwd(rd_ctx*, const void* g, void* out, int N, int H, int W, int C, int dtype, rd_stream);
int rd_flip_absdiff_bwd(rd_ctx*, const void* g, const void* dout, void* dg, int N, int H, int W, int C, int dtype, rd_stream);

/* ---- small dense layers (nn.Linear: src/model.py:2359-2364, 2499) — fp32 -------------------- */
int rd_linear_fwd(rd_ctx*, const float* x, const float* W, const float* b, float* y, int rows, int in_f,
                  int out_f, int act /* rd_act */, float slope, rd_stream);
int rd_linear_bwd(rd_ctx*, const float* x, const float* W, const float* dy, float* dx, float* dW, float* db,
                  int rows, int in_f, int out_f, rd_stream);   /* dW, db are += ; dx may be NULL */
/* z = mu + eps*exp(0.5*logvar)  (src/model.py:3159-3162) */
int rd_sample_fwd(rd_ctx*, const float* mu, const float* logvar, const float* eps, float* z, int64_t n, rd_stream);
int rd_sample_bwd(rd_ctx*, const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar,
                  int64_t n, rd_stream);

/* ---- missing-modality fusion (src/model.py:3239-3246): boolean gather, row-major over (b, m) -- */
/* si: [M][B] images of `row_elems` elements (modality-major stack); mask [B][M] fp32 (==1 selects).
 * out rows [0,K) receive image (b,m) in (b,m) row-major order; idx_out[k] = b*M+m; *count_out = K. */
int rd_fuse_gather_fwd(rd_ctx*, const void* si, const float* mask, void* out, int32_t* idx_out,
                       int32_t* count_out, int B, int M, int64_t row_elems, int dtype, rd_stream);
int rd_fuse_gather_bwd(rd_ctx*, const void* dout, const float* mask, void* dsi, int B, int M, int64_t row_elems,
                       int dtype, rd_stream);  /* dsi fully written (zeros where not selected) */

/* ---- losses ------------------------------------------------------------------------------ */
/* per-row mean of |gt - x|^p over row_elems (compute_recon_loss, src/model.py:3260-3266).
 * x rows r in [0,R); gt row for r = gt_index[r] (device int32, <0 = unused row -> loss 0).  */
int rd_recon_rows_fwd(rd_ctx*, const void* x, const void* gt, int gt_dtype, const int32_t* gt_index,
                      float* row_loss, float* partial, int R, int64_t row_elems, int p, int dtype, rd_stream);
/* dx[r] = coef[r] * d/dx mean|gt-x|^p  (coef device fp32, already includes the upstream gradient) */
int rd_recon_rows_bwd(rd_ctx*, const void* x, const void* gt, int gt_dtype, const int32_t* gt_index,
                      const float* coef, void* dx, int R, int64_t row_elems, int p, int dtype, rd_stream);
/* w[i] = [mask[:, i].sum() != 0] / #such i — the mean over non-skipped contrasts of compute_segmentation_loss_y_list
 * (src/model.py:3299-3313) as device-side weights (no host read of the mask: stage 2 stays graph-capturable). */
int rd_modality_weights(rd_ctx*, const float* mask, float* w, int B, int M, rd_stream);
/* masked combination of the per-row losses, exactly the python loops of
 * compute_recon_loss_x_list (:3315) [kind 0], compute_recon_loss_x_mix_list (:3327, index lag Q4) [kind 1].
 * row_loss rows: kind 0 -> [M][B]; kind 1 -> [M(M-1)][B] where row t belongs to the t-th NON-skipped
 * pair; plan (rd_xmix_plan) gives gt modality per t.  Writes loss[0] and coef[r] = dloss/drow_loss[r]. */
int rd_xmix_plan(rd_ctx*, const float* mask, int32_t* gt_index /* [M(M-1)*B] */, int B, int M, rd_stream);
int rd_masked_combine(rd_ctx*, const float* row_loss, const float* mask, float* loss, float* coef,
                      int B, int M, int kind, rd_stream);
/* latent / similarity / KL terms on (M,B,Z) fp32 stacks: loss[0] and gradients in one launch */
int rd_latent_z_loss(rd_ctx*, const float* mu, const float* mu_new, const float* mask, float* loss,
                     float* dmu, float* dmu_new, int B, int M, int Z, rd_stream);           /* :3384 */
int rd_sim_z_loss(rd_ctx*, const float* z, const float* mask, float margin, float* loss, float* dz,
                  int B, int M, int Z, rd_stream);                                           /* :3537 */
int rd_kl_loss(rd_ctx*, const float* mu, const float* logvar, const float* mask, float* loss, float* dmu,
               float* dlogvar, int B, int M, int Z, rd_stream);                              /* :3343 */
/* compute_compact_s_max (:3448): 16x16 max-pool of NHWC s -> pooled [N][C*(H/16)*(W/16)] fp32 (+argmax) */
int rd_maxpool16_fwd(rd_ctx*, const void* s, float* pooled, int32_t* argmax, int N, int H, int W, int C,
                     int dtype, rd_stream);
int rd_maxpool16_bwd(rd_ctx*, const float* dpooled, const int32_t* argmax, void* ds, int N, int H, int W, int C,
                     int dtype, rd_stream);   /* ds fully written */
/* compute_compact_s_mean (:3453): 16x16 average pool, same output order */
int rd_avgpool16_fwd(rd_ctx*, const void* s, float* pooled, int N, int H, int W, int C, int dtype, rd_stream);
int rd_avgpool16_bwd(rd_ctx*, const float* dpooled, void* ds, int N, int H, int W, int C, int dtype, rd_stream);
/* compute_similarity_s_loss (:3478) on pooled vectors of modalities i, j: hinge(margin - cos(si,sj) + cos(roll si, si)) */
int rd_sim_s_loss(rd_ctx*, const float* pooled /* [M][B][D] */, const float* mask,
                  const int32_t* pair /* DEVICE int32[2] = (i, j), drawn on the host (np.random.choice, :3485) */,
                  float margin, float* loss, float* dpooled, int B, int M, int D, rd_stream);
/* compute_segmentation_loss_y (:3287): weighted CE [1,5,5,5] + soft Dice over classes 1..3; y NHWC C=4 */
int rd_seg_loss_fwd(rd_ctx*, const void* y, const float* target /* [N][HW] labels */, float* loss, float* partial,
                    int N, int64_t hw, int dtype, rd_stream);
int rd_seg_loss_bwd(rd_ctx*, const void* y, const float* target, const float* partial, const float* upstream,
                    void* dy, int N, int64_t hw, int dtype, rd_stream);

/* ---- optimizer: clip_grad_norm_(1.0) + Adam(amsgrad, wd) (src/main_missing.py:118,272,283) ---- */
/* segments: device int64 [nseg][2] = (offset, length) into the flat fp32 buffers (params that receive grads) */
int rd_grad_norm(rd_ctx*, const float* grad, const int64_t* segments, int nseg, float* partial, float* scalars,
                 float max_norm, rd_stream);   /* scalars[0]=total norm, [1]=clip coef, [2]=finite flag */
int rd_grad_scale(rd_ctx*, float* grad, const int64_t* segments, int nseg, const float* scalars, rd_stream);
/* hyper: device fp32 [8] = {lr, beta1, beta2, eps, weight_decay, step, 1 - beta1, 1 - beta2}; step is incremented here.  The last two
 * are the complements rounded from DOUBLE (torch.optim.Adam passes `1 - beta2` as a Python float); 0 = derive them in fp32. */
int rd_adam_amsgrad(rd_ctx*, float* param, const float* grad, float* m, float* v, float* vmax,
                    const int64_t* segments, int nseg, float* hyper, rd_stream);
/* The three steps of main_missing.py:272-284 in one pass: g = grad * scalars[1] (what rd_grad_scale would have stored; scalars may
 * be NULL = no clipping), the Adam(amsgrad) update above, and (zero_grad != 0) optimizer.zero_grad() of the same segments. */
int rd_clip_adam_amsgrad(rd_ctx*, float* param, float* grad, float* m, float* v, float* vmax, const int64_t* segments, int nseg,
                         float* hyper, const float* scalars, int zero_grad, rd_stream);
/* The same with torch.optim.Adam's per-parameter "grad is None -> skip" rule and per-parameter step counters (src/main_missing.py:118,
 * 282-284: modules the masked loss terms never reach in an accumulation window keep grad None).  seg_param: device int32 [nseg] =
 * parameter index of each segment; partial: the per-segment squared sums rd_grad_norm just wrote; param_flags (int32 [nparams], out):
 * 1 = the parameter received a gradient; param_steps (fp32 [nparams], in/out): Adam step count per parameter. */
int rd_clip_adam_amsgrad_gated(rd_ctx*, float* param, float* grad, float* m, float* v, float* vmax, const int64_t* segments,
                               const int32_t* seg_param, int nseg, const float* partial, int32_t* param_flags, float* param_steps,
                               int nparams, float* hyper, const float* scalars, int zero_grad, rd_stream);

/* ---- evaluation metrics on the device (src/util.py:935-992; callers src/main_missing.py:519-533) ---- */
/* compute_reconstruction_metrics: for image n the pair (target[t_index ? t_index[n] : n, :, :, ct0], pred[n, :, :, cp0]) of NHWC tensors
 * with Ct / Cp channels (the reference compares channel 0 only): both shifted to min 0, data_range R = max(target - min), then
 * skimage.metrics mean_squared_error / peak_signal_noise_ratio(data_range=R) / structural_similarity(data_range=R) (7x7 uniform
 * window, sample covariance, K1 0.01, K2 0.03, mean over the interior).  out fp32 [N][3] = {ssim, psnr, mse};
 * workspaces: stats fp32 [N][3], partial fp64 [N][rd_metrics_recon_tiles(H, W)][2]. */
int rd_metrics_recon_tiles(int H, int W);
int rd_metrics_recon(rd_ctx*, const void* target, int t_dtype, int Ct, int ct0, const int32_t* t_index, const void* pred, int p_dtype,
                     int Cp, int cp0, int N, int H, int W, float* stats, double* partial, float* out, rd_stream);
/* compute_segmentation_metrics: target fp32 (N, H*W) labels, pred NHWC (N, H*W, Cp >= 3) logits; class i in {0,1,2}: (target == i+1)
 * against (pred[..., i] > 0.5) exactly as src/util.py:984-990 indexes them; out fp32 [N][2] = {mean dice, mean iou} with the +1 smoothing. */
int rd_metrics_seg(rd_ctx*, const float* target, const void* pred, int p_dtype, int Cp, int N, int64_t hw, float* out, rd_stream);

/* ---- slab assembly from a device-resident volume store (ZeroDoseDataset.__getitem__, src/util.py:471-566) ---- */
/* vols fp32 (S, M, D, H, W), present uint8 (S, M); tvols fp32 (S, D, H, W) or NULL, has_target uint8 (S); brain_mask fp32 (D, H, W) or
 * NULL (skull_strip); per sample b: subj[b], slice_idx[b] (clamped to [block, clamp_hi - block] like :476-483), drop[b] = contrast
 * index removed by the random dropoff (:538-542, drawn on the host with the reference's NumPy calls) or -1.
 * Writes inputs fp32 (B, M*(2*block+1), H, W), targets (B, 1, H, W) (label 4 -> 3 when remap4, :527), mask (B, M),
 * mask_img (B, H, W) = (inputs[:, 0] == 0) (:563-564). */
int rd_assemble_slabs(rd_ctx*, const float* vols, const uint8_t* present, const float* tvols, const uint8_t* has_target,
                      const float* brain_mask, const int32_t* subj, const int32_t* slice_idx, const int32_t* drop, float* inputs,
                      float* targets, float* mask, float* mask_img, int B, int M, int block, int D, int H, int W, int remap4,
                      int clamp_hi, rd_stream);

#ifdef __cplusplus
}
#endif
#endif /* RD_B200_H_ */
