/*
 * hicdiff_b200 -- C ABI of the B200-native HiCDiff reverse-diffusion sampling path.
 *
 * The reference (BioinfoMachineLearning/hicdiff) has no FFI / plugin layer: its boundary for this path is the
 * Python class surface of `Unet` / `hicedrn_Diff` / `GaussianDiffusion`.  This header is the C boundary that sits
 * directly UNDER that surface; every entry point below cites the reference code it replaces.  All pointers are
 * plain device (or, where stated, host) pointers, all sizes are explicit, no torch / C++ types cross the boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero code on failure; `hd_last_error()` then returns a
 *     thread-local, human readable message (the Python layer raises RuntimeError with it);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the device
 *     unless stated;
 *   - tiles are fp32 [B, 1, 64, 64] contiguous (the reference's NCHW with C == 1);
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef HICDIFF_B200_H_
#define HICDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HD_ABI_VERSION 1

#if defined(__GNUC__)
#define HD_API __attribute__((visibility("default")))
#else
#define HD_API
#endif

typedef struct hd_plan hd_plan;

/* eps-predictor variants. */
enum hd_variant {
    HD_UNET = 0,        /* Unet of src/hicdiff_condition.py:255-384 (self_condition=True) and src/hicdiff.py (False) */
    HD_UNET_SR3 = 1,    /* Unet(noise_level_emb=True) of src/hicdiff_sr3.py:310-445                                   */
    HD_HICEDRN = 2,     /* hicedrn_Diff of src/model/hicedrn_Diff.py:210-297                                          */
    HD_HICEDRN_SR3 = 3  /* hicedrn_Diff of src/model/hicedrn_sr3_Diff.py                                              */
};

/* Mirrors the constructor arguments the reference scripts actually use. */
typedef struct hd_config {
    int32_t abi_version;      /* must be HD_ABI_VERSION                                                              */
    int32_t variant;          /* enum hd_variant                                                                     */
    int32_t self_condition;   /* 1: the conditional (noisy) tile is concatenated as input channel 0                  */
    int32_t dim;              /* Unet `dim` (64); ignored for HiCEDRN (n_feat = 256)                                 */
    int32_t num_mults;        /* len(dim_mults), <= 8                                                                */
    int32_t dim_mults[8];     /* (1, 2, 4, 8)                                                                        */
    int32_t image_size;       /* 64                                                                                  */
    int32_t timesteps;        /* T (GaussianDiffusion.num_timesteps)                                                 */
    int32_t num_blocks;       /* HiCEDRN number_resnet (32); ignored for Unet                                        */
    int32_t debug_keep;       /* 1: keep every intermediate activation addressable via hd_debug_read               */
    int32_t reserved[8];      /* reserved[0]: option bits, 0 = defaults.  bit 0: GroupNorm applied in the conv epilogue;
                               * bits 1-2: padded-slab conv form; bit 3: CTA pairs (tcgen05 cta_group::2);
                               * bits 4-5: GEMM precision -- 0 = bf16 operands (default), 1 = "bf16w2": bf16 activations,
                               * conv weights as hi + lo bf16 pairs (two MMAs per product; the reference computes in fp32,
                               * src/hicdiff_condition.py:90,105); bit 6: the 3x3, Cout = 64 convs issue one MMA per filter tap
                               * instead of the default dx-stacked form (three taps per N = 192 MMA); bit 7: ResnetBlocks with
                               * a res_conv run res_conv + block2's GroupNorm apply as ONE launch (opt-in, measured not
                               * faster; hicdiff_condition.py:191-197).  Other words: 0.                                 */
} hd_config;

/* -------------------------------------------------------------------------------------------------------------
 * Plan life cycle.  Replaces nn.Module construction + load_state_dict + .to(device) for the sampling path
 * (inference.py:59-94).  hd_plan_set_weight is called once per state_dict entry under `model.` (key WITHOUT the
 * `model.` prefix, e.g. "downs.0.0.block1.proj.weight"); the data is copied, the caller keeps ownership.
 * ------------------------------------------------------------------------------------------------------------- */
HD_API int hd_plan_create(const hd_config* cfg, hd_plan** out);
HD_API int hd_plan_set_weight(hd_plan* plan, const char* key, const float* dev_ptr, const int64_t* shape, int32_t ndim,
                       void* stream);
/* Schedule tables of GaussianDiffusion.__init__ (hicdiff_condition.py:491-519), all fp32 [T] DEVICE pointers:
 * sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2 and
 * sigma = exp(0.5 * posterior_log_variance_clipped) (p_sample :597).  `time_values` is what the eps-net receives as
 * `time` at step t: float(t) for Unet/HiCEDRN (:594), sqrt_alphas_cumprod_prev[t+1] for SR3 (hicdiff_sr3.py:636). */
HD_API int hd_plan_set_schedule(hd_plan* plan, const float* sqrt_recip, const float* sqrt_recipm1, const float* coef1,
                         const float* coef2, const float* sigma, const float* time_values, int32_t T, void* stream);
/* Standardises + re-lays-out weights (bf16 GEMM layout), builds the [T, .] time-embedding table.  Must be called
 * after all weights and the schedule are set and again after any weight changes.  Synchronises `stream`. */
HD_API int hd_plan_finalize(hd_plan* plan, void* stream);
HD_API void hd_plan_destroy(hd_plan* plan);

/* -------------------------------------------------------------------------------------------------------------
 * eps-predictor forward.  Replaces Unet.forward / hicedrn_Diff.forward (hicdiff_condition.py:345-384,
 * hicedrn_Diff.py:267-289).  `time` is fp32 [B] (integer timesteps converted to float, or SR3 noise levels).
 * `cond` may be NULL iff self_condition == 0.  x, cond, eps: fp32 [B,1,64,64].
 * ------------------------------------------------------------------------------------------------------------- */
HD_API int hd_eps_forward(hd_plan* plan, const float* x, const float* cond, const float* time, float* eps, int32_t B,
                   void* stream);

/* One reverse step on a caller-owned state: replaces p_sample (hicdiff_condition.py:591-598) given eps.
 * x is updated in place; `noise` is the z tensor for this step (ignored when t == 0) or NULL for Philox
 * (seed, tile_offset).  x0_out (optional) receives the clipped x_start. */
HD_API int hd_ddpm_step(hd_plan* plan, float* x, const float* eps, const float* noise, float* x0_out, int32_t t, int32_t B,
                 uint64_t seed, uint64_t tile_offset, void* stream);

/* Full reverse process.  Replaces p_sample_loop (hicdiff_condition.py:600-623, hicdiff.py:603-620,
 * hicdiff_sr3.py:654-677): x_T ~ N(0, I), then t = T-1 .. 0.  One CUDA graph per step, replayed T times with a
 * device-side step counter.
 *   cond   fp32 [B,1,64,64] or NULL (unconditional)
 *   noise  NULL -> Philox(seed) with per-tile streams keyed by (tile_offset + b), so results do not depend on
 *          how tiles are sharded; else fp32 [T, B,1,64,64]: noise[0] = x_T, noise[i] = z of step t = T - i
 *          (the exact draw order of the reference: :605 then :596, no draw at t == 0)
 *   out    fp32 [B,1,64,64] final x_0
 *   trace  optional fp32 [T, B,1,64,64]: x after each step (return_all_timesteps), or NULL
 *   t_start / t_end: run steps t = t_start .. t_end (inclusive, descending); pass T-1, 0 for the full chain.
 *          When t_start < T-1 the chain starts from `out`'s current content instead of fresh noise. */
HD_API int hd_sample(hd_plan* plan, const float* cond, const float* noise, float* out, float* trace, int32_t B,
              uint64_t seed, uint64_t tile_offset, int32_t t_start, int32_t t_end, void* stream);

/* One step of the DDRM sampler for the denoising operator (every singular value 1), given eps = model(x_t, t): replaces the
 * body of the loop of efficient_generalized_steps (src/functions/denoising.py:49-104; H = Denoising,
 * src/functions/svd_replacement.py:148-168).  With all singular values equal the three masked cases collapse to one per
 * step, chosen by the caller: mode 0 (sigma_next > sigma_0): c0 = etaB, c1 = 1 - etaB, c2 = sqrt(sigma_next^2 - sigma_0^2 etaB^2);
 * mode 1 (sigma_next < sigma_0): c0 = sqrt(sigma_next^2 - (sigma_next etaA)^2), c1 = sigma_next etaA; mode 2 (equal):
 * c0 = sqrt(sigma_next^2 - (sigma_next etaC)^2), c1 = sigma_next etaC.  x is updated in place to x_{t_next}; x0_out (optional)
 * receives x0_t; noise is this step's z or NULL for Philox(seed, tile_offset + tile, step_id).  n = elements (B*4096). */
HD_API int hd_ddrm_step(float* x, const float* eps, const float* y, const float* noise, float* x0_out, int32_t mode, float sqrt_at,
                 float sqrt_1m_at, float sqrt_at_next, float c0, float c1, float c2, float sigma_0, int64_t n, uint64_t seed,
                 uint64_t tile_offset, uint32_t step_id, void* stream);

/* One step of DDIM sampling given eps = model(x_t, t): replaces the loop body of ddim_sample (src/hicdiff.py:636-659,
 * src/hicdiff_condition.py:639-664): x_start = clamp(sqrt_recip * x - sqrt_recipm1 * eps, -1, 1); last != 0 (time_next < 0):
 * x = x_start; else x = x_start * sqrt_a_next + c * eps + sigma * z.  The five scalars are the caller's fp32 evaluations of the
 * reference expressions (:648-657).  noise: this step's z or NULL for Philox(seed, tile_offset + tile, step_id); n = B * 4096. */
HD_API int hd_ddim_step(float* x, const float* eps, const float* noise, float* x0_out, float sqrt_recip, float sqrt_recipm1,
                 float sqrt_a_next, float c, float sigma, int32_t last, int64_t n, uint64_t seed, uint64_t tile_offset,
                 uint32_t step_id, void* stream);

/* Per-tile SSIM and MSE of fp32 [B,1,64,64] tiles: _ssim of src/Utils/loss/SSIM.py:17-36 with the caller's 11x11 window
 * (create_window :10-14, 121 floats on the device) and zero "same" padding; rescale = 1 first applies
 * inverse_data_transform('rescaled') (src/datasets/__init__.py:214-223).  ssim_out / mse_out: fp32 [B] on the device; the
 * reference's scalar SSIM is their mean, its PSNR 10 log10(1 / mean(mse_out)). */
HD_API int hd_ssim_mse_tiles(const float* a, const float* b, const float* window121, float* ssim_out, float* mse_out, int32_t B,
                      int32_t rescale, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * Training step.  Replaces the forward + loss + loss.backward() of one iteration (train.py:120-129,
 * pretrain/train_*.py) for every eps-net variant: eps = model(x_t, t, cond) (src/model/hicedrn_Diff.py:267-289,
 * src/hicdiff_condition.py:345-384, src/hicdiff_sr3.py:410-445), loss = mean(|eps - target|^p * weight[b]) (p_losses, hicdiff_condition.py:741-746,
 * loss_fn :706-713; weight = p2_loss_weight[t]), and d loss / d parameter for every parameter of the net.
 *   hd_trainer_create    any hd_variant (the SR3 variants take the continuous noise level as `time`; Unet: dim = 64,
 *                        dim_mults in {1,2,4,8}); `batch` tiles per step (fixed per trainer)
 *   hd_trainer_bind      once per state_dict entry (key without the `model.` prefix): `param` is READ IN PLACE at every
 *                        step (so any optimizer may update it between steps), `grad` (same shape, fp32) is OVERWRITTEN by
 *                        every step; both are device pointers the caller keeps alive
 *   hd_trainer_step      x_t = q_sample(x_start, t, noise) (:698-704), cond (NULL iff self_condition == 0), time fp32 [B],
 *                        target fp32 [B,1,64,64] (the noise), weight fp32 [B]; loss_type 0 = l1, 1 = l2.
 *                        eps_out (optional) fp32 [B,1,64,64]; loss_out: device fp32 scalar.
 * Gradients are deterministic (fixed-order two-pass reductions, no atomics).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct hd_trainer hd_trainer;
HD_API int hd_trainer_create(const hd_config* cfg, int32_t batch, hd_trainer** out);
HD_API int hd_trainer_bind(hd_trainer* trainer, const char* key, const float* param, float* grad, const int64_t* shape,
                    int32_t ndim);
HD_API int hd_trainer_finalize(hd_trainer* trainer, void* stream);
HD_API int hd_trainer_step(hd_trainer* trainer, const float* x_t, const float* cond, const float* time, const float* target,
                    const float* weight, int32_t loss_type, float* eps_out, float* loss_out, void* stream);
/* Launch groups of one step / per-kernel-family timing of one step as JSON {"family": {"ms", "flops", "ops"}} / bytes held */
/* Data-parallel replicas (BASELINE config 5; DistributedDataParallel's bucketed, overlapped gradient all-reduce -- the reference
 * itself has no distributed code).  Call BEFORE hd_trainer_finalize with n module-name prefixes in BACKWARD order
 * (e.g. "ups.2.", "ups.0.", "downs.2."): bucket k holds the gradients of the modules the backward pass finishes before it reaches
 * the first module whose name starts with prefix k (time-embedding and init_conv gradients are only final at the end of the
 * step and belong to no bucket).  Every step then records one event per bucket inside its CUDA graph;
 * hd_trainer_wait_grad_bucket makes `stream` (the caller's communication stream) wait for bucket k of the step most recently
 * enqueued, so its all-reduce overlaps the rest of the backward.  Unet trainers only. */
HD_API int hd_trainer_set_grad_buckets(hd_trainer* trainer, const char* const* first_module_prefix, int32_t n);
HD_API int hd_trainer_wait_grad_bucket(hd_trainer* trainer, int32_t bucket, void* stream);
HD_API int hd_trainer_num_launches(hd_trainer* trainer);
HD_API int hd_trainer_profile(hd_trainer* trainer, int32_t reps, char* buf, int64_t buflen, void* stream);
HD_API int64_t hd_trainer_device_bytes(hd_trainer* trainer);
HD_API void hd_trainer_destroy(hd_trainer* trainer);
/* The optimiser of the training loop (train.py:111 `torch.optim.Adam(diffusion.parameters(), lr=2e-5)`, stepped at :129) as one
 * launch over all parameter tensors; the arithmetic follows torch's multi-tensor Adam operation by operation (L2-style
 * weight_decay added to the gradient; no amsgrad / maximize).  params: nparams device pointers (fp32, contiguous) with numels
 * elements each, updated in place; the moments are owned by the handle (zero-initialised).  hd_adam_step takes this step's
 * gradient pointers (same order, fp32, contiguous; the table is re-uploaded only when a pointer changed), counts the step
 * itself and is asynchronous on `stream`.  hd_adam_state exposes parameter `index`'s moments (device pointers into the flat
 * state) and the step count for checkpointing; hd_adam_set_step restores the count. */
typedef struct hd_adam hd_adam;
HD_API int hd_adam_create(const void* const* params, const int64_t* numels, int32_t nparams, hd_adam** out);
HD_API int hd_adam_step(hd_adam* adam, const void* const* grads, double lr, double beta1, double beta2, double eps, double weight_decay,
                 void* stream);
HD_API int hd_adam_state(hd_adam* adam, int32_t index, float** exp_avg, float** exp_avg_sq, int64_t* step);
HD_API int hd_adam_set_step(hd_adam* adam, int64_t step);
HD_API void hd_adam_destroy(hd_adam* adam);
/* The two conv gradients as single operators (256 -> 256, 3x3 "same", 64x64 tiles; activations bf16 NHWC as uint16):
 * dw fp32 [256,256,3,3] = d/dW of conv2d(x, W) given dy; dx = d/dx given dy and W (fp32, reference layout).  Synchronise. */
HD_API int hd_op_conv3x3_wgrad(const uint16_t* x, const uint16_t* dy, float* dw, int32_t B, void* stream);
/* General conv weight gradient (the Unet's shapes): x [B,H,W,Cin], dy [B,H,W,Cout] bf16 NHWC, Cin / Cout multiples of 64, W in
 * {8,16,32,64}, ksize 1 or 3 ("same").  Writes columns [ci0, ci0 + Cin) of dw fp32 [Cout, cin_total, k, k] -- the slice that
 * belongs to one operand of a channel concat (cin_total = Cin, ci0 = 0 for a plain conv).  Synchronises. */
HD_API int hd_op_conv_wgrad(const uint16_t* x, const uint16_t* dy, float* dw, int32_t B, int32_t H, int32_t W, int32_t Cin,
                     int32_t Cout, int32_t ksize, int32_t cin_total, int32_t ci0, void* stream);
/* Backward of Block's tail (hicdiff_condition.py:159-171): s = SiLU(GroupNorm_8(y) * (scale + 1) + shift).  y, ds, dy bf16
 * [B,P,C] (dy may alias ds); gamma / beta [C]; scale / shift [B,C] or both NULL (block2); outputs dgamma / dbeta [C] and,
 * when FiLM is present, dscale / dshift [B,C]; dconv_bias [C] (optional) = sum over samples and pixels of dy, i.e. the bias
 * gradient of the conv that produced y, derived from the same reductions.  eps = 1e-5.  Synchronises. */
HD_API int hd_op_groupnorm_silu_bwd(const uint16_t* y, const uint16_t* ds, const float* gamma, const float* beta, const float* scale,
                             const float* shift, uint16_t* dy, float* dgamma, float* dbeta, float* dscale, float* dshift,
                             float* dconv_bias, int32_t B, int32_t P, int32_t C, void* stream);
/* Backward of the channel LayerNorm (hicdiff_condition.py:99-108): z = (x - mean_c) * rsqrt(var_c + 1e-5) * g; x, dz, dx bf16
 * [M,C] (dx must not alias x), dg fp32 [C].  Synchronises. */
HD_API int hd_op_channel_layernorm_bwd(const uint16_t* x, const uint16_t* dz, const float* g, uint16_t* dx, float* dg, int64_t M,
                                int32_t C, void* stream);
/* Backward of WeightStandardizedConv2d's weight transform (:89-95): dw [Cout,K] from the gradient w.r.t. the standardised
 * weight (dw may alias dwt); K = Cin * k * k. */
HD_API int hd_op_weight_standardize_bwd(const float* w, const float* dwt, float* dw, int32_t Cout, int32_t K, void* stream);
/* Backward of the attention cores given d_out [B,n,128]: linear = 1 LinearAttention (hicdiff_condition.py:212-227),
 * linear = 0 Attention at 8x8 (:239-251, n = 64).  qkv / dqkv bf16 [B,n,384].  Synchronises. */
HD_API int hd_op_attention_bwd(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, int32_t B, int32_t n, int32_t linear,
                        void* stream);
HD_API int hd_op_conv3x3_dgrad(const uint16_t* dy, const float* w, uint16_t* dx, int32_t B, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * Data preparation: loadBothConstraints (processdata/PrepareData_linear.py:48-103) and the noise
 * injection of split_numpy (:183-213) on the device.  All results are bit-exact (indexing, integer histograms, IEEE fp32).
 *   hd_coo_to_dense        the loop :67-72: mat[r - smallbin, c - smallbin] = mat[c - smallbin, r - smallbin] = v for every triple,
 *                          LATER triples overwriting earlier ones; rows / cols int64 bin indices, vals fp32, mat fp32 [n, n]
 *                          (fully overwritten).  Synchronises.
 *   hd_remove_empty_bins   :77-85: delete the rows and columns whose diagonal entry is 0 or NaN; out fp32 [m, m] (capacity
 *                          n * n), kept_idx int64 [n] (first m valid), *n_kept_host = m.  Synchronises.
 *   hd_select_ranks        exact order statistics (0-based ranks, ascending) of x[0..n) -- the two neighbours np.percentile
 *                          (:88) interpolates between; ranks / results are HOST arrays.  Synchronises.
 *   hd_normalize_contacts  :90-92 in place: x = 2 * (clip(x, 0, per) / per) - 1
 *   hd_add_noise           :203-204 for the 'deno' operator: out = x + sigma * noise (noise: caller's N(0,1) draws)
 * ------------------------------------------------------------------------------------------------------------- */
HD_API int hd_coo_to_dense(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz, int64_t smallbin, int64_t n,
                    float* mat, void* stream);
HD_API int hd_remove_empty_bins(const float* mat, int64_t n, float* out, int64_t* kept_idx, int64_t* n_kept_host, void* stream);
HD_API int hd_select_ranks(const float* x, int64_t n, const int64_t* ranks_host, int32_t nranks, float* out_host, void* stream);
HD_API int hd_normalize_contacts(float* x, int64_t n, float per, void* stream);
HD_API int hd_add_noise(const float* x, const float* noise, float sigma, int64_t n, float* out, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * Tiling.  hd_tile_extract replaces splitPieces (processdata/PrepareData_linear.py:25-46): zero-pad the n x n
 * matrix to a multiple of `piece`, enumerate block rows i and block columns j >= i with (j - i) <= band_blocks
 * (= 4 * int(40000 / res)), row-major.  hd_tile_scatter is its exact inverse (the reference has none): tile k is
 * written at block (i, j) and, for i != j, transposed at (j, i); padding is cropped; cells outside the band are
 * left untouched (zero them first).  Pure indexing: bit-exact.
 * ------------------------------------------------------------------------------------------------------------- */
HD_API int64_t hd_tile_count(int64_t n, int32_t piece, int32_t band_blocks);
HD_API int hd_tile_extract(const float* mat, int64_t n, float* tiles, int32_t piece, int32_t band_blocks, void* stream);
HD_API int hd_tile_scatter(const float* tiles, float* mat, int64_t n, int32_t piece, int32_t band_blocks, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * Single-operator entry points (same kernels the plan launches).  Used by the parity tests to pin each kernel
 * against torch fp32 ops; layouts: activations bf16 NHWC [B,H,W,C] as raw uint16 device buffers.
 * ------------------------------------------------------------------------------------------------------------- */
/* conv: w fp32 [Cout,Cin,k,k] in reference layout; standardize=1 applies WeightStandardizedConv2d's transform.
 * x1 (second concat operand, C1 channels) may be NULL.  mode 0: k x k "same" conv; mode 1: Downsample (pixel
 * unshuffle + 1x1, w is [Cout, 4*C0, 1, 1]; x0 is [B,2H,2W,C0], output [B,H,W,Cout]).  res (optional) is added.
 * Bits above bit 0 of `standardize` select opt-in kernel forms for the parity tests: bits 1-2 padded slab, bit 3 CTA pairs,
 * bits 4-5 form of the 3x3, Cout = 64 conv (0 default: dx-stacked with two epilogue groups, 1: one MMA per tap, 2: dx-stacked, one group). */
HD_API int hd_op_conv2d(const uint16_t* x0, int32_t C0, const uint16_t* x1, int32_t C1, const float* w, const float* bias,
                 const uint16_t* res, uint16_t* out, int32_t B, int32_t H, int32_t W, int32_t Cout, int32_t ksize,
                 int32_t mode, int32_t standardize, void* stream);
/* Block (hicdiff_condition.py:155-171) as ONE launch: 3x3 conv whose epilogue applies GroupNorm(8) + FiLM + SiLU (+ res).
 * Only shapes on the slab path (H*W >= 128, W >= 16, Cout in {64, 128}); scale/shift/res optional. */
HD_API int hd_op_conv_gn(const uint16_t* x0, int32_t C0, const uint16_t* x1, int32_t C1, const float* w, const float* bias,
                  const float* gamma, const float* beta, const float* scale, const float* shift, const uint16_t* res,
                  uint16_t* out, int32_t B, int32_t H, int32_t W, int32_t Cout, int32_t standardize, void* stream);
HD_API int hd_op_groupnorm_silu(const uint16_t* x, uint16_t* y, const float* gamma, const float* beta, const float* scale,
                         const float* shift, const uint16_t* res, int32_t B, int32_t P, int32_t C, void* stream);
HD_API int hd_op_channel_layernorm(const uint16_t* x, uint16_t* y, const float* g, const uint16_t* res, int32_t B, int32_t H,
                            int32_t W, int32_t C, int32_t upsample2x, void* stream);
HD_API int hd_op_linear_attention(const uint16_t* qkv, uint16_t* out, int32_t B, int32_t n, void* stream);
/* Whole Residual(PreNorm(LinearAttention)) block (hicdiff_condition.py:64-70,99-118,199-227) as the fused three-launch
 * path the plan uses for C in {64, 128}, n >= 128: g1 [C] PreNorm gain, wqkv [384, C], wo [C, 128], bo [C], g2 [C].
 * bound_out (host, optional) receives the analytic softmax bound; fails if it exceeds the fused path's limit. */
HD_API int hd_op_linattn_block(const uint16_t* x, const float* g1, const float* wqkv, const float* wo, const float* bo,
                        const float* g2, uint16_t* y, int32_t B, int32_t n, int32_t C, float* bound_out, void* stream);
HD_API int hd_op_full_attention(const uint16_t* qkv, uint16_t* out, int32_t B, int32_t n, void* stream);
HD_API int hd_op_stem_conv(const float* x0, const float* x1, const float* w, const float* bias, uint16_t* y, int32_t B,
                    int32_t Cout, int32_t Cin, int32_t ksize, void* stream);
HD_API int hd_op_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t tile_offset, void* stream);

/* Debug: copy a named intermediate of the last hd_eps_forward(B) as fp32 NCHW into `out` (device), needs
 * cfg.debug_keep.  Returns its element count through *numel; out may be NULL to query. */
HD_API int hd_debug_read(hd_plan* plan, int32_t B, const char* name, float* out, int64_t* numel, int32_t* shape4,
                  void* stream);
HD_API int hd_debug_names(hd_plan* plan, int32_t B, char* buf, int64_t buflen);

/* Number of kernels the plan launches per hd_eps_forward / per sampling step at batch B (for gpu_launches). */
HD_API int hd_plan_launches_per_step(hd_plan* plan, int32_t B, int32_t* eps_launches, int32_t* step_launches);
/* Built-in profiler: runs every kernel launch of one sampling step at batch B `reps` times between CUDA events on
 * `stream` and writes a JSON array [{"tag","kernel","ms","flops","bytes"}, ...] (average ms per launch, algorithmic
 * FLOPs / HBM bytes per launch) into buf.  Synchronises. */
HD_API int hd_plan_profile_step(hd_plan* plan, int32_t B, int32_t reps, char* buf, int64_t buflen, void* stream);
/* Device bytes held by the plan (weights + tables + workspaces). */
HD_API int64_t hd_plan_device_bytes(hd_plan* plan);

HD_API const char* hd_last_error(void);
HD_API int hd_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HICDIFF_B200_H_ */
