// Internal launcher interface between the plan/graph builder (plan.cu) and the kernels.
// Everything here is device-pointer + stream based; no torch types anywhere in csrc/.
//
// Activation layout in HBM: NHWC, bf16, i.e. a [B*H*W, C] row-major matrix whose rows are
// pixels.  The sample state x_t, eps and everything the posterior update touches stay fp32
// in the reference's NCHW layout (C == 1, so NCHW == NHWC there).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: the calls are no-ops unless a tool (nsys / ncu --nvtx) is attached

namespace hd {

// Launch with programmatic stream serialization (PDL): the kernel's CTAs may be scheduled while the previous kernel of the stream
// drains; the kernel itself calls ptx::grid_dep_wait() before it touches anything an earlier kernel wrote (ptx.cuh).  Inside a
// stream capture this becomes a programmatic dependency edge of the graph.
bool pdl_enabled();   // plan.cu: HD_PDL = 0 | 1; default: only inside a PdlScope (the training step), see there
struct PdlScope {
    PdlScope();
    ~PdlScope();
    PdlScope(const PdlScope&) = delete;
    PdlScope& operator=(const PdlScope&) = delete;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// NVTX range around a host-side phase (plan finalize, graph capture, one sampling chain, one training step, ...), so a
// timeline shows which C-ABI call a launch belongs to (the reference has no profiler hooks at all).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};


typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// conv_gemm.cu -- tcgen05/TMEM implicit-GEMM convolution (3x3 pad 1, 1x1, pixel-unshuffle 1x1)
// ---------------------------------------------------------------------------------------------
struct ConvSrc {
    const bf16* ptr;  // [B, H, W, C] NHWC
    int C;            // channels (multiple of 64)
};

enum ConvMode : int {
    CONV_TAPS = 0,       // kh x kw taps, stride 1, "same" zero padding (kh==kw in {1,3})
    CONV_UNSHUFFLE = 1,  // 'b c (h p1) (w p2) -> b (c p1 p2) h w' followed by a 1x1 conv
    CONV_UPSAMPLE = 2,   // nearest x2 upsample followed by a 3x3 "same" conv, as four 2x2 phase convs over the low-res input
};

struct ConvEpilogue {
    const float* bias = nullptr;      // [N]
    // per-step feature-wise modulation (HiCEDRN): v = v * (1 + scale[c]) + shift[c]
    const float* film = nullptr;      // table base, row-major [rows, film_ld]
    const int* film_row = nullptr;    // device row index (step counter) or per-sample rows
    int film_row_stride = 0;          // 0: one row for the whole batch, 1: film_row[b]
    int film_ld = 0;
    int film_off = 0;                 // column offset of this layer's slice inside a row
    int film_has_scale = 0;           // 1: [scale(N) | shift(N)], 0: [shift(N)] only
    int silu = 0;                     // 1: v = v * sigmoid(v) with one tanh.approx; 2: the same through ex2 + rcp ("bf16w2" precision)
    float out_scale = 1.0f;           // v *= out_scale
    const bf16* res = nullptr;        // v += res[m, c]   (row stride ldr)
    int ldr = 0;
    // ResnetBlock tail in one pass (hicdiff_condition.py:191-197): when set, `res` is the RAW output of block2's conv and
    // gnres[b * N + c] = (mul, add) folds its GroupNorm(8) affine (groupnorm_finalize_run): v += SiLU(res * mul + add).  The
    // launch is then res_conv(x) + block2's norm / activation / skip add: the separate groupnorm_apply pass and the
    // res_conv output tensor disappear (one 1x1 conv launch instead of conv + a 3-stream elementwise pass).
    const float2* gnres = nullptr;
    float* out_f32 = nullptr;         // if set: write fp32 [M, n_valid] instead of bf16
    int n_valid = 0;
    // optional second bf16 output [M, N]: lo = bf16(v - bf16(v)), so out + out_lo carries ~16 mantissa bits.  Used where the
    // consumer cancels a large common component (to_out conv -> channel LayerNorm; scripts/precision_study.py: rounding
    // that one tensor to bf16 shifts the T = 1000 PSNR by 1.3e-2 dB, every other activation rounding together by 3e-4 dB).
    bf16* out_lo = nullptr;
    // GroupNorm statistics of THIS conv's output, from the fp32 accumulators (+bias): one (sum, M2) pair per
    // (32-row warp block, 8-channel piece), M2 = sum of squared deviations from the piece-block's own mean (merged
    // stably, in a fixed order, by groupnorm_apply_kernel).
    float2* gn_part = nullptr;        // [M / 32][N / 8]
    // Fused GroupNorm(8) apply (conv_gemm_can_fuse_gn): when gn_gamma is set the epilogue itself waits for the image's
    // statistics and writes silu(GN(acc + bias) * (scale + 1) + shift) [+ post-add] [+ res]; `film` then modulates the
    // NORMALISED value (film_has_scale = 0: the row is an SR3 additive embedding applied after the activation).
    const float* gn_gamma = nullptr;  // [N]
    const float* gn_beta = nullptr;   // [N]
    float gn_eps = 1e-5f;
    int* gn_counter = nullptr;        // [B] per-image arrival counters, zero when the kernel starts
};

struct ConvGemmDesc {
    ConvSrc src0, src1;       // src1.ptr == nullptr when there is no channel concat
    int B, H, W;              // OUTPUT spatial size (== input size for CONV_TAPS; input is 2H x 2W for UNSHUFFLE,
                              // H/2 x W/2 for UPSAMPLE)
    int ksize;                // 1 or 3 (CONV_TAPS); ignored for UNSHUFFLE
    ConvMode mode;
    const bf16* weight;       // [Npad, Ktot] K-major; K order = (tap, concat channel); UPSAMPLE: 4 stacked phase matrices
    int N;                    // rows of `weight` (multiple of the chosen N tile)
    bf16* out;                // [B*H*W, N]
    ConvEpilogue epi;
    int cg2_mode = 0;         // 1: run on CTA pairs (tcgen05 cta_group::2) where the kind supports it
    int pad_mode = 0;         // 3x3, N == 64: 1 = padded-slab form where >= 2 stages fit, 2 = wherever it fits, 0 = never
    int dx3_mode = 2;         // 3x3, N == 64, resident weights: dx-stacked form (N = 192 MMAs, shifts in the epilogue) with 2 = two epilogue
                              // groups on alternating tiles, 1 = one epilogue group; 0 = one MMA per tap
    int static_weights = 0;   // 1: `weight` is not written by any kernel of the graph this launch belongs to (the sampling plan's
                              // weights, prepared at finalize): the resident-weight load may then precede the PDL wait (ptx.cuh).
                              // The trainer re-derives its bf16 layouts every step -> 0.
    int wsplit = 0;           // 1: `weight` holds hi + lo bf16 pairs (rows of 2 * Ktot, per tap [hi | lo]): the K loop walks the
                              // input channels twice per tap (x * hi + x * lo), "bf16w2" precision
};

// Opaque prepared launch (tensor maps encoded once, replayed inside CUDA graphs).
struct ConvGemmLaunch {
    CUtensorMap tmA0, tmA1, tmB, tmD, tmD31, tmD30;   // tmD31 / tmD30: padded-slab kind, 31- / 30-row output boxes
    int bn;             // N tile (16, 64, 128 or 256)
    int gnf;            // 1: GroupNorm-fused epilogue
    int cg;             // 1, or 2: CTA pairs (clusters of two, tcgen05 cta_group::2); num_m_tiles then counts tile PAIRS
    int m_tiles_real;   // real 128-row M tiles
    int kind;           // 0 general, 1 slab (3x3, one A box per (chunk, dx)), 2 slab + shared-memory resident weights,
                        // 3 padded slab (one A box per chunk serves all nine taps; GroupNorm partials use the padded layout)
                        // 4 dx-stacked resident slab (N == 64: MMAs of 192 columns = 3 dx taps, shifted and summed in the epilogue)
                        // 5 the same with two epilogue groups of 8 warps on alternating tiles (576 threads; where shared memory allows)
    int grid;
    int smem_bytes;
    // kernel scalar arguments
    int M, N, num_m_tiles, num_n_tiles, num_tiles, nkb, chunks0, chunks1, mode, W, P, kh, kw, pad, stages;
    int passes;         // passes over the sources' channel chunks per tap: 1, or 2 with split (hi + lo) weights
    int static_weights; // see ConvGemmDesc
    int Wl_box, rows_box;
    int PW, tiles_per_img, Hh;   // padded-slab kind: row pitch W + 2, 128-position tiles per image, image height
    uint32_t slab_bytes, slab_dy_bytes, res_b_bytes;
    ConvEpilogue epi;
    bf16* out;
    int ldo;
};

// True when a 3x3 conv of this shape runs on the slab path, whose epilogue can apply the following GroupNorm itself.
bool conv_gemm_can_fuse_gn(int B, int H, int W, int N, int ksize, ConvMode mode);
// Returns 0 on success; on failure fills `err` (size errlen).
int conv_gemm_prepare(const ConvGemmDesc& d, int num_sms, ConvGemmLaunch* out, char* err, int errlen);
cudaError_t conv_gemm_run(const ConvGemmLaunch& l, cudaStream_t s);

// bf16 tiled tensor map with 128-byte swizzle (driver entry point resolved at run time; no libcuda link dependency)
int encode_tmap_bf16(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, char* err, int errlen, int swizzle_bytes = 128);

// ---------------------------------------------------------------------------------------------
// norm.cu -- GroupNorm(8)+FiLM+SiLU (cluster/DSMEM two-pass), channel LayerNorm
// ---------------------------------------------------------------------------------------------
struct GroupNormArgs {
    const bf16* x;        // [B, P, C]
    bf16* y;              // [B, P, C]
    int B, P, C;          // P = H*W pixels
    const float2* part;   // [B * P / 32][C / 8] (sum, M2) partials written by the producing conv (or groupnorm_stats_run)
    int part_tpi = 0;     // > 0: partials come from a padded-slab conv: part_tpi * 4 warp blocks per image, block k covers
    int part_W = 0;       //      padded positions [32k, 32k + 32) of a (W + 2)-pitch image (halo columns carry no data)
    const float* gamma;   // [C]
    const float* beta;    // [C]
    float eps;
    // FiLM: y = gn * (scale + 1) + shift, row picked like ConvEpilogue
    const float* film = nullptr;
    const int* film_row = nullptr;
    int film_row_stride = 0;
    int film_ld = 0;
    int film_off = 0;           // scale at film_off + c, shift at film_off + C + c
    // SR3: additive per-channel noise embedding applied AFTER the activation
    const float* postadd = nullptr;   // same row addressing as film; value at postadd_off + c
    int postadd_off = 0;
    const bf16* res = nullptr;        // + res[b, p, c] after the activation (ResnetBlock skip)
    float2* stats_out = nullptr;      // [B, 8] (optional) the (mean, rstd) this pass normalised with, kept for the training backward
    int exact_act = 0;                // 1: SiLU through ex2 + rcp instead of one tanh.approx ("bf16w2" precision)
};
cudaError_t groupnorm_film_silu_run(const GroupNormArgs& a, cudaStream_t s);
// Statistics only: merges the conv's partials exactly like groupnorm_film_silu_run (same order, same arithmetic) and writes the folded
// per-(image, channel) affine ma[b * C + c] = (gamma * rstd, beta - mean * gamma * rstd); dense partial layout, no FiLM.
cudaError_t groupnorm_finalize_run(const float2* part, const float* gamma, const float* beta, float eps, float2* ma, int B, int P, int C,
                                   cudaStream_t s);
// Stand-alone producer of the same partials from a bf16 tensor (used when the input does not come from conv_gemm).
cudaError_t groupnorm_stats_run(const bf16* x, float2* part, int B, int P, int C, cudaStream_t s);

struct LayerNormArgs {
    const bf16* x;    // [M, C]
    bf16* y;          // [M, C] (or upsampled [B, 2H, 2W, C] when upsample2x)
    int M, C;
    const float* g;   // [C]
    float eps;
    const bf16* res = nullptr;  // + res[m, c] (Residual wrapper)
    const bf16* x_lo = nullptr; // optional low half of x (x = x + x_lo, ConvEpilogue::out_lo): "bf16w2" precision
    int upsample2x = 0;         // nearest x2: write each pixel to its 2x2 block
    int H = 0, W = 0;           // needed for upsample2x
};
cudaError_t channel_layernorm_run(const LayerNormArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// attention.cu
// ---------------------------------------------------------------------------------------------
struct LinAttnArgs {
    const bf16* qkv;   // [B, n, 384]: q = [0,128), k = [128,256), v = [256,384); head h owns 32 channels
    bf16* out;         // [B, n, 128]
    float* ctx;        // scratch, >= B*4*32*32*2 bytes: normalised ctx^T[b][head][e][d] in bf16
    int B, n;
};
cudaError_t linear_attention_run(const LinAttnArgs& a, cudaStream_t s);

// linattn_fused.cu -- the whole Residual(PreNorm(LinearAttention)) block in three launches (C <= 128, n >= 128)
struct LinAttnFusedDesc {
    const bf16* x;          // [B, n, C] block input (also the residual)
    bf16* y;                // [B, n, C] block output
    int B, n, C;
    const bf16* wqkv;       // [384, C] to_qkv weight with the PreNorm gain folded in (linattn_prep_run)
    const float* rowsum;    // [384] row sums of wqkv (LayerNorm mean correction)
    const float* kshift;    // [128] softmax_n shift per k channel
    const float* wo;        // [C, 128] to_out.0 weight (reference layout, fp32)
    const float* bo;        // [C]
    const float* g2;        // [C] to_out.1 gain
    float* ctx_part;        // scratch [B * max_parts][128][32]
    float* s_part;          // scratch [B * max_parts][128]
    bf16* mb;               // scratch [B][C][128]
    int max_parts;
    float eps;
};
struct LinAttnFusedLaunch {
    LinAttnFusedDesc d;
    CUtensorMap tmX, tmXk, tmW, tmM, tmY;
    int parts, tiles_per_unit, num_tiles, out_grid, num_sms;
};
cudaError_t linattn_prep_run(const float* wqkv, const float* g1, bf16* out, float* rowsum, float* kshift,
                             int* max_bound_bits, int C, cudaStream_t s);
int linattn_fused_prepare(const LinAttnFusedDesc& d, int num_sms, LinAttnFusedLaunch* out, char* err, int errlen);
cudaError_t linattn_fused_run(const LinAttnFusedLaunch& l, cudaStream_t s);

struct FullAttnArgs {
    const bf16* qkv;   // [B, n, 384], n <= 64
    bf16* out;         // [B, n, 128]
    int B, n;
};
cudaError_t full_attention_run(const FullAttnArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// pointwise.cu -- stem conv (tiny Cin), 1x1 head to eps, DDPM posterior step, Philox noise
// ---------------------------------------------------------------------------------------------
struct StemConvArgs {
    const float* x0;   // first input channel plane [B, H, W] fp32 (cond when self_condition, else x)
    const float* x1;   // second plane or nullptr
    const float* w;    // [Cout, Cin, k, k] fp32 (reference layout)
    const float* bias; // [Cout]
    bf16* y;           // [B, H, W, Cout]
    int B, H, W, Cout, Cin, ksize;
    int precise = 0;   // 1: fp32 inputs and weights on the FMA pipe (stem_conv_kernel) instead of bf16 mma.sync fragments
};
cudaError_t stem_conv_run(const StemConvArgs& a, cudaStream_t s);

struct HeadConvArgs {
    const bf16* x;     // [M, C]
    const float* w;    // [C]
    const float* bias; // [1]
    float* eps;        // [M]
    int M, C;
};
cudaError_t head_conv1x1_run(const HeadConvArgs& a, cudaStream_t s);

// Device-resident control block of a sampling run; written once per hd_sample / hd_ddpm_step call, advanced on
// the device, so a single captured CUDA graph can be replayed for every step.
struct SampleCtl {
    int step;                        // current t (FiLM table row and coefficient row)
    int noise_single;                // 1: `noise` is the z tensor of this very step; 0: base of [T, n], row T - t
    const float* noise;              // nullptr -> Philox
    unsigned long long seed;
    unsigned long long tile_offset;  // first global tile id of this batch (world-size independent streams)
};

// coefficient table row: {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1,
//                         posterior_mean_coef2, exp(0.5 * posterior_log_variance_clipped), 0, 0, 0}, T rows
struct PosteriorArgs {
    float* x;               // [n] in/out sample state
    const float* eps;       // [n]
    const float* coef;      // [T, 8]
    const SampleCtl* ctl;   // device
    int T;
    long long n;            // elements (B*H*W)
    int tile_elems;         // H*W
    float* x0_out;          // optional: clipped x_start
};
cudaError_t posterior_step_run(const PosteriorArgs& a, cudaStream_t s);

// One DDRM step for the denoising operator (all singular values 1): mode 0 "noisier than y", 1 "less noisy than y",
// 2 "sigma_next == sigma_0"; c0..c2 are the mode's coefficients (see ddrm_step_kernel).
struct DdrmArgs {
    float* x;               // [n] in: x_t, out: x_{t_next}
    const float* eps;       // [n]
    const float* y;         // [n] observation y_0
    const float* noise;     // [n] z of this step, or nullptr -> Philox(seed, tile, step_id)
    float* x0_out;          // optional: x0_t
    int mode;
    float sqrt_at, sqrt_1m_at, sqrt_at_next, c0, c1, c2, sigma_0;
    long long n;
    int tile_elems;
    unsigned long long seed, tile_offset;
    unsigned int step_id;
};
cudaError_t ddrm_step_run(const DdrmArgs& a, cudaStream_t s);

// One DDIM step (ddim_sample, hicdiff.py:623-664): x0 = clamp(sr * x - srm1 * eps, -1, 1); last step: x = x0; else
// x = x0 * sqrt_a_next + c * eps + sigma * z
struct DdimArgs {
    float* x;               // [n] in: x_t, out: x_{t_next}
    const float* eps;       // [n]
    const float* noise;     // [n] z of this step, or nullptr -> Philox(seed, tile, step_id)
    float* x0_out;          // optional clipped x_start
    float sr, srm1, sqrt_a_next, c, sigma;
    int last;               // 1: time_next < 0
    long long n;
    int tile_elems;
    unsigned long long seed, tile_offset;
    unsigned int step_id;
};
cudaError_t ddim_step_run(const DdimArgs& a, cudaStream_t s);

cudaError_t philox_normal_run(float* out, long long n, unsigned long long seed, unsigned long long tile_offset,
                              int tile_elems, unsigned long long stream_id, cudaStream_t s);
cudaError_t step_advance_run(SampleCtl* ctl, int delta, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// prep.cu -- one-off weight preparation and the time-embedding tables
// ---------------------------------------------------------------------------------------------
// [Cout, Cin, k, k] fp32 -> [Npad, k*k*Cin] bf16 with K order (tap, cin); optional weight standardisation
// training: every conv's forward (qf) and dgrad (qd) bf16 layouts + weight-standardisation statistics in two launches
struct PrepSlot {
    const float* w;
    bf16* qf;
    bf16* qd;
    float2* stats;
    int Cout, Cin, ksize, ws;
};
cudaError_t prep_weights_batched_run(const PrepSlot* slots, const int2* fwd_rows, int nfwd, const int2* bwd_rows, int nbwd, float eps,
                                     cudaStream_t s);
// split = 1: hi + lo bf16 pairs, rows of 2K with per-tap [hi | lo] halves (the "bf16w2" precision mode, prep.cu)
cudaError_t prep_conv_weight_run(const float* w, bf16* out, int Cout, int Cin, int ksize, int standardize, float eps,
                                 int Npad, cudaStream_t s, int split = 0);
// Downsample 1x1 weight [Cout, 4*C] with K order (c, p1, p2) -> (p1, p2, c)
cudaError_t prep_unshuffle_weight_run(const float* w, bf16* out, int Cout, int C, cudaStream_t s, int split = 0);
// Upsample's 3x3 weight [Cout, Cin, 3, 3] -> four phase matrices [4][Cout][(u, v, cin)] of the equivalent 2x2 convs
// over the low-resolution input (taps that land on the same low-res pixel are summed in fp32)
cudaError_t prep_upsample_weight_run(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s, int split = 0);

// y[r, ldy*r + off + j] = bias[j] + sum_k act(x[r, k]) * W[j, k];  in_act: 0 none, 1 SiLU, 2 GELU(erf) (on the input), out_act: 0 none, 1 GELU(erf)
cudaError_t linear_rows_run(const float* x, int ldx, const float* W, const float* bias, float* y, int ldy, int off,
                            int rows, int in_f, int out_f, int in_act, int out_act, cudaStream_t s);
// mode 0: SinusoidalPosEmb(dim) of integer timesteps t[r] (as float); mode 1: SR3 PositionalEncoding of level[r]
cudaError_t posenc_rows_run(const float* t, float* y, int rows, int dim, int mode, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// tiles.cu -- splitPieces-order tile extraction / reassembly
// ---------------------------------------------------------------------------------------------
cudaError_t tile_extract_run(const float* mat, int n, float* tiles, int piece, int band_blocks, cudaStream_t s);
cudaError_t tile_scatter_run(const float* tiles, float* mat, int n, int piece, int band_blocks, cudaStream_t s);
int tile_count(int n, int piece, int band_blocks);

// ---------------------------------------------------------------------------------------------
// metrics.cu -- per-tile SSIM (reference window, zero padding) and MSE of 64x64 tiles
// ---------------------------------------------------------------------------------------------
cudaError_t ssim_mse_tiles_run(const float* a, const float* b, const float* window, float* ssim_out, float* mse_out, int B,
                               int rescale, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// wgrad.cu / train_kernels.cu -- the HiCEDRN training step (backward of conv / FiLM / SiLU / time MLPs, loss)
// ---------------------------------------------------------------------------------------------
struct WgradLaunch {
    CUtensorMap tmG, tmX;   // 5-D views {64 ch, W, H, B, C / 64} of the NHWC dY and conv input
    int kb_total, H, nsplit;
    float* part;            // [nsplit][9][256][256] fp32
};
int wgrad_prepare(const bf16* g, const bf16* x, int B, int H, int W, int C, int nsplit, float* part, WgradLaunch* out, char* err,
                  int errlen);
cudaError_t wgrad_run(const WgradLaunch& l, cudaStream_t s);
// general form: dY [B,H,W,Cout], X [B,H,W,Cin] NHWC bf16 (Cout, Cin multiples of 64; W in {8,16,32,64}; ksize 1 or 3)
struct WgradGenLaunch {
    CUtensorMap tmG, tmX, tmXs;      // tmXs: the halo'd 66-pixel slab box of the filter-row form
    int B, H, W, rows_kb, Cout, Cin, Nt, taps, mtiles, ntiles, nsplit, kb_total;
    int rowmode = 0;                 // 1: one CTA per (filter row, K split) with three accumulators (3x3, W = 64, Cin tile <= 128)
    float* part;            // [nsplit][taps][Cout][Cin] fp32, wgrad_general_part_bytes(); set by the caller after prepare
};
int wgrad_general_prepare(const bf16* g, const bf16* x, int B, int H, int W, int Cout, int Cin, int ksize, int num_sms,
                          WgradGenLaunch* out, char* err, int errlen);
size_t wgrad_general_part_bytes(const WgradGenLaunch& l);
cudaError_t wgrad_general_run(const WgradGenLaunch& l, cudaStream_t s);
// dw [Cout, cin_total, k, k] columns [ci0, ci0 + Cin) (+)= scale * sum over the K splits (the slice of a channel concat)
cudaError_t wgrad_general_reduce_run(const WgradGenLaunch& l, int cin_total, int ci0, float scale, int accumulate, float* dw,
                                     cudaStream_t s);
// dW [256, 256, 3, 3] (+)= scale * sum over the K splits
cudaError_t wgrad_reduce_run(const float* part, int nsplit, float scale, int accumulate, float* dw, cudaStream_t s);
size_t wgrad_part_bytes(int nsplit);

cudaError_t prep_dgrad_weight_run(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s);
cudaError_t flip_tail_weight_run(const float* w, float* out, int C, cudaStream_t s);
// has_scale = 1: row = [scale(C) | shift(C)] (hicedrn_Diff); 0: row = [shift(C)] only (SR3 FeatureWiseAffine)
cudaError_t film_silu_fwd_run(const bf16* a, bf16* sout, const float* film, int ld, int off, int B, int P, int C, int has_scale,
                              cudaStream_t s);
int film_bwd_part_floats(int B, int C);
// da (may alias ds) = ds * SiLU'(a * (scale + 1) + shift) * (scale + 1); dfilm[b, off + c] = d scale, [off + C + c] = d shift
cudaError_t film_silu_bwd_run(const bf16* ds, const bf16* a, bf16* da, const float* film, float* dfilm, int ld, int off, int B,
                              int P, int C, float* part, int has_scale, cudaStream_t s);
cudaError_t edrn_bias_grad_run(const float* film, const float* dfilm, int ld, int off, int B, const float* colsum_g, float g_scale,
                               float* dbias, int C, int has_scale, cudaStream_t s);
int colsum_parts(long long M);
// out[c] (+)= scale * sum_m x[m, c]
cudaError_t colsum_run(const bf16* x, long long M, int C, float* part, float scale, int accumulate, float* out, cudaStream_t s);
cudaError_t sum_parts_run(const float* part, int nparts, int n, float scale, int accumulate, float* out, cudaStream_t s);
cudaError_t add_bf16_run(const bf16* a, const bf16* b, bf16* y, long long n, cudaStream_t s);
// dw[(c * nk + k) * 9 + tap] = sum_{b,y,x} G[b,y,x,c] * u_k[b, y + sgn*(ky-1), x + sgn*(kx-1)];  part: [B * 8][nk][9][256]
cudaError_t thin_wgrad_run(const bf16* G, const float* u0, const float* u1, int sgn, int B, float* part, float* dw, cudaStream_t s);
int loss_parts();
cudaError_t loss_grad_run(const float* eps, const float* target, const float* w, int loss_type, int B, int tile_elems, float* d_eps,
                          float* part, float* loss, cudaStream_t s);
cudaError_t sum_f32_run(const float* x, long long n, float* part, float* out, cudaStream_t s);
cudaError_t linear_bwd_weight_run(const float* dY, int ldy, int off, const float* X, int ldx, int rows, int in_f, int out_f,
                                  int in_act, float* dW, float* db, cudaStream_t s);
cudaError_t linear_bwd_input_run(const float* dY, int ldy, int off, const float* W, int rows, int in_f, int out_f, int accumulate,
                                 float* dX, int ldx, cudaStream_t s);
struct LinSlot {            // one per-block time-embedding Linear inside the FiLM row
    const float* W;         // [width, in_f]
    float* gW;              // its gradient
    float* gb;              // bias gradient [width]
    int off, width;         // column slice of the FiLM row
};
cudaError_t linear_bwd_weight_batched_run(const float* dY, int ldy, const float* X, int ldx, int rows, int in_f, const LinSlot* slots,
                                          int nslots, int max_width, cudaStream_t s);
cudaError_t linear_bwd_input_batched_run(const float* dY, int ldy, int rows, int in_f, const LinSlot* slots, int nslots, float* part,
                                         float* dX, cudaStream_t s);
cudaError_t act_apply_run(const float* x, float* y, long long n, int act, cudaStream_t s);   // act: 1 SiLU, 2 GELU(erf)
cudaError_t act_grad_run(float* d, const float* x, long long n, int act, cudaStream_t s);
// fused Adam: one block per chunk (<= ADAM_CHUNK elements of one parameter tensor); exp_avg / exp_avg_sq are flat state buffers
struct AdamChunk {
    float* param;
    const float* grad;
    long long state_off;
    int n;
    int pad;
};
constexpr int ADAM_CHUNK = 4096;
cudaError_t adam_step_run(const AdamChunk* chunks, int nchunks, float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2,
                          double eps, double weight_decay, long long step, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// norm_bwd.cu -- backward of GroupNorm(8)+FiLM+SiLU, channel LayerNorm, weight standardisation (Unet training blocks)
// ---------------------------------------------------------------------------------------------
struct GroupNormBwdArgs {
    const bf16* y;        // [B, P, C] the GroupNorm INPUT (conv output)
    const bf16* ds;       // [B, P, C] gradient w.r.t. the block output SiLU(...)
    bf16* dy;             // [B, P, C] gradient w.r.t. y (may alias ds)
    int B, P, C;
    const float* gamma;   // [C]
    const float* beta;    // [C]
    float eps;
    const float* scale = nullptr;   // [B, C] FiLM scale (nullptr: no FiLM, e.g. block2)
    const float* shift = nullptr;   // [B, C]
    int ld = 0;                     // row stride of scale / shift / dscale / dshift (0: C)
    float* dgamma;        // [C]
    float* dbeta;         // [C]
    float* dscale = nullptr;        // [B, C] (optional)
    float* dshift = nullptr;        // [B, C]
    float* dconv_bias = nullptr;    // [C] (optional) sum_{b,p} dy: the bias gradient of the conv that produced y, from the same sums
    float* dpost = nullptr;         // [B, C] (optional) sum_p ds: gradient of an SR3 embedding added AFTER the activation
    const float2* stats_in = nullptr;   // [B, 8] (optional) the forward pass's (mean, rstd) (GroupNormArgs::stats_out); recomputed from y when absent
};
size_t gn_bwd_scratch_floats(int B, int P, int C);
cudaError_t groupnorm_silu_bwd_run(const GroupNormBwdArgs& a, float* scratch, cudaStream_t s);
int ln_bwd_blocks(long long M);
// z = LayerNorm_C(x) * gain: dx (may NOT alias x), dgain [C]; part: ln_bwd_blocks(M) * C floats
cudaError_t channel_layernorm_bwd_run(const bf16* x, const bf16* dz, const float* gain, long long M, int C, float eps, bf16* dx,
                                      float* dgain, float* part, cudaStream_t s);
// dw [Cout, K] from the gradient w.r.t. the standardised weight (dw may alias dwt)
cudaError_t weight_standardize_bwd_run(const float* w, const float* dwt, int Cout, int K, float eps, float* dw, cudaStream_t s);

// attention_bwd.cu -- backward of the linear-attention core and of the 8x8 softmax attention; qkv / dqkv [B, n, 384], dout [B, n, 128]
size_t linattn_bwd_scratch_floats(int B);
cudaError_t linear_attention_bwd_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, float* scratch, cudaStream_t s);
// the same through mma.sync tiles (attention_bwd_mma.cu); kmax / ksum [B*4][32], cd [B*4][2][32][32] fp32 scratch
cudaError_t linear_attention_bwd_mma_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, float* kmax, float* ksum,
                                         float* cd, cudaStream_t s);
cudaError_t full_attention_bwd_run(const bf16* qkv, const bf16* dout, bf16* dqkv, int B, int n, cudaStream_t s);

// unet_train_kernels.cu -- resampling copies, thin convs at the ends of the Unet, standardised dgrad weights
cudaError_t unshuffle_run(const bf16* in, bf16* out, int B, int H, int W, int C, int inverse, cudaStream_t s);   // H, W: LOW-res size
cudaError_t upsample2x_run(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s);               // in [B,H,W,C]
cudaError_t sumpool2x_run(const bf16* in, bf16* out, int B, int H, int W, int C, cudaStream_t s);                // out [B,H,W,C]
cudaError_t stem_wgrad_run(const bf16* G, const float* u0, const float* u1, int B, int C, int ksize, float* part, float* dw, cudaStream_t s);
int head_bwd_parts(long long M);
cudaError_t head_bwd_run(const bf16* x, const float* d_eps, const float* w, long long M, int C, bf16* dx, float* part, float* dw,
                         cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// dataprep.cu -- contact triples -> dense matrix, empty-bin removal, exact percentile, normalisation, noise injection
// ---------------------------------------------------------------------------------------------
size_t coo_scratch_bytes(long long n);
// synchronises; *bad_host != 0 when a triple fell outside [smallbin, smallbin + n)
cudaError_t coo_to_dense_run(const long long* rows, const long long* cols, const float* vals, long long nnz, long long smallbin,
                             long long n, float* mat, void* scratch, int* bad_host, cudaStream_t s);
cudaError_t keep_map_run(const float* mat, long long n, long long* map, long long* n_kept_dev, cudaStream_t s);
cudaError_t compact_run(const float* mat, long long n, const long long* map, long long m, float* out, cudaStream_t s);
size_t select_scratch_bytes();
cudaError_t select_rank_run(const float* x, long long n, unsigned long long rank, float* out, void* scratch, cudaStream_t s);
cudaError_t normalize_contacts_run(float* x, long long n, float per, cudaStream_t s);
cudaError_t axpy_noise_run(const float* x, const float* z, float sigma, long long n, float* y, cudaStream_t s);

}  // namespace hd
