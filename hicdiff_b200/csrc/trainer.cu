// hd_trainer: one training step (forward, loss, backward) of GaussianDiffusion.p_losses over the hicedrn_Diff eps-net, the
// model the reference's train.py trains (train.py:84-107 builds hicedrn_Diff + GaussianDiffusion(loss_type='l2'),
// :109-136 runs loss = diffusion(x); loss.backward(); optimizer.step()).  Reference math:
//   forward   hicedrn_Diff.forward              /root/reference/src/model/hicedrn_Diff.py:267-289 (ResnetBlock :194-208)
//   loss      p_losses                          /root/reference/src/hicdiff_condition.py:715-746 (loss_fn :706-713)
//   backward  torch.autograd over the above     (no reference source: the chain rule of the forward, restated in
//                                                oracle/hicdiff_oracle.py::hicedrn_loss_and_grads with torch.autograd)
// Parameters are bound by pointer (the caller's fp32 tensors are read in place every step, so a stock torch optimizer
// keeps working) and gradients are written into caller-provided fp32 buffers.  Activations are bf16 NHWC; the convs run on
// tcgen05 (conv_gemm.cu forward and data gradient, wgrad.cu weight gradient); statistics-free glue is in train_kernels.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "trainer_internal.h"

using namespace hd;

namespace {

#define T_TRY(expr)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) return tfail("%s failed: %s", #expr, cudaGetErrorString(e__));   \
    } while (0)

constexpr int F = 256;          // n_feat
constexpr int S = 64;           // tile edge
constexpr int P = S * S;
constexpr int TIME_DIM = 1024;
constexpr int WG_SPLITS = 16;   // 9 taps x 16 K splits = 144 CTAs

}  // namespace

namespace {

struct TB {   // builder state
    hd_trainer* t;
    bool ok = true;
    void push(const char* kernel, const std::string& tag, std::function<cudaError_t(cudaStream_t)> fn, double flops = 0) {
        TOp o;
        o.fn = std::move(fn); o.tag = tag; o.kernel = kernel; o.flops = flops;
        t->ops.push_back(std::move(o));
    }
    // 3x3 256 -> N conv over [B, 64, 64, 256] on conv_gemm
    bool conv(const std::string& tag, const bf16* in, const bf16* wq, int N, bf16* out, ConvEpilogue epi) {
        ConvGemmDesc d;
        d.src0 = ConvSrc{in, F};
        d.src1 = ConvSrc{nullptr, 0};
        d.B = t->B; d.H = S; d.W = S; d.ksize = 3; d.mode = CONV_TAPS;
        d.weight = wq; d.N = N; d.out = out; d.epi = epi;
        ConvGemmLaunch l;
        char e[256];
        if (conv_gemm_prepare(d, t->num_sms, &l, e, sizeof(e))) { tfail("%s: %s", tag.c_str(), e); return ok = false; }
        push("conv_gemm", tag, [l](cudaStream_t s) { return conv_gemm_run(l, s); }, 2.0 * t->B * P * (N < F ? 1 : N) * 9.0 * F);
        return true;
    }
};

int build(hd_trainer* t) {
    const int B = t->B, nb = t->nb;
    const int cin = t->cfg.self_condition ? 2 : 1;
    const bool sr3 = t->cfg.variant == HD_HICEDRN_SR3;   // FeatureWiseAffine: h = conv(x) + Linear(t) (no scale, no SiLU on t)
    const int fw = sr3 ? F : 2 * F;                       // FiLM row width per block
    const int hs = sr3 ? 0 : 1;
    const int ld = nb * fw;
    const size_t act_elems = static_cast<size_t>(B) * P * F;
    const long long M = static_cast<long long>(B) * P;
    TB b{t};

    // ---------------------------------------------------------------- parameters
    const TParam *head_w = find_p(t, "head.weight", {F, cin, 3, 3}), *head_b = find_p(t, "head.bias", {F});
    const TParam *w1 = find_p(t, "time_mlp.1.weight", {TIME_DIM, F}), *b1 = find_p(t, "time_mlp.1.bias", {TIME_DIM});
    const TParam *w3 = find_p(t, "time_mlp.3.weight", {TIME_DIM, TIME_DIM}), *b3 = find_p(t, "time_mlp.3.bias", {TIME_DIM});
    const TParam *bt_w = find_p(t, "body_tail.weight", {F, F, 3, 3}), *bt_b = find_p(t, "body_tail.bias", {F});
    const TParam *tail_w = find_p(t, "tail.weight", {1, F, 3, 3}), *tail_b = find_p(t, "tail.bias", {1});
    if (!head_w || !head_b || !w1 || !b1 || !w3 || !b3 || !bt_w || !bt_b || !tail_w || !tail_b) return 1;
    std::vector<const TParam*> cw(nb), cb(nb), mw(nb), mb(nb);
    for (int i = 0; i < nb; ++i) {
        const std::string pre = "body." + std::to_string(i);
        cw[i] = find_p(t, pre + ".conv.proj.weight", {F, F, 3, 3});
        cb[i] = find_p(t, pre + ".conv.proj.bias", {F});
        mw[i] = find_p(t, pre + (sr3 ? ".noise_func.noise_func.0.weight" : ".mlp.1.weight"), {fw, TIME_DIM});
        mb[i] = find_p(t, pre + (sr3 ? ".noise_func.noise_func.0.bias" : ".mlp.1.bias"), {fw});
        if (!cw[i] || !cb[i] || !mw[i] || !mb[i]) return 1;
    }

    // ---------------------------------------------------------------- buffers
    if (dalloc(t, &t->x, M * 4, true) || dalloc(t, &t->cond, M * 4, true) || dalloc(t, &t->target, M * 4, true) ||
        dalloc(t, &t->eps, M * 4) || dalloc(t, &t->d_eps, M * 4) || dalloc(t, &t->time, B * 4, true) ||
        dalloc(t, &t->weight, B * 4, true) || dalloc(t, &t->loss, 16, true))
        return 1;
    std::vector<bf16*> wqf(nb + 1), wqd(nb + 1);   // forward / data-gradient GEMM layouts; index nb = body_tail
    for (int i = 0; i <= nb; ++i)
        if (dalloc(t, &wqf[i], static_cast<size_t>(F) * 9 * F * 2) || dalloc(t, &wqd[i], static_cast<size_t>(F) * 9 * F * 2)) return 1;
    bf16* wq_tail = nullptr;
    float *tail_bias16 = nullptr, *tail_flip = nullptr, *zero_bias = nullptr;
    if (dalloc(t, &wq_tail, static_cast<size_t>(16) * 9 * F * 2) || dalloc(t, &tail_bias16, 16 * 4, true) ||
        dalloc(t, &tail_flip, F * 9 * 4) || dalloc(t, &zero_bias, F * 4, true))
        return 1;
    float *posenc = nullptr, *z1 = nullptr, *g1 = nullptr, *temb = nullptr, *stemb = nullptr, *film = nullptr, *dfilm = nullptr, *d_act = nullptr, *d_g1 = nullptr;
    if (dalloc(t, &posenc, static_cast<size_t>(B) * F * 4) || dalloc(t, &z1, static_cast<size_t>(B) * TIME_DIM * 4) ||
        dalloc(t, &temb, static_cast<size_t>(B) * TIME_DIM * 4) || dalloc(t, &g1, static_cast<size_t>(B) * TIME_DIM * 4) ||
        dalloc(t, &stemb, static_cast<size_t>(B) * TIME_DIM * 4) || dalloc(t, &film, static_cast<size_t>(B) * ld * 4) ||
        dalloc(t, &dfilm, static_cast<size_t>(B) * ld * 4, true) || dalloc(t, &d_act, static_cast<size_t>(B) * TIME_DIM * 4) ||
        dalloc(t, &d_g1, static_cast<size_t>(B) * TIME_DIM * 4))
        return 1;
    std::vector<bf16*> X(nb + 1), A(nb), Sx(nb);
    for (int i = 0; i <= nb; ++i) if (dalloc(t, &X[i], act_elems * 2)) return 1;
    for (int i = 0; i < nb; ++i) if (dalloc(t, &A[i], act_elems * 2) || dalloc(t, &Sx[i], act_elems * 2)) return 1;
    bf16 *BT = nullptr, *DBT = nullptr, *GX[2] = {nullptr, nullptr}, *DS = nullptr;
    if (dalloc(t, &BT, act_elems * 2) || dalloc(t, &DBT, act_elems * 2) || dalloc(t, &GX[0], act_elems * 2) ||
        dalloc(t, &GX[1], act_elems * 2) || dalloc(t, &DS, act_elems * 2))
        return 1;
    float *ws1 = nullptr, *ws2 = nullptr, *film_part = nullptr, *cs_part = nullptr, *cs_g = nullptr, *thin_part = nullptr,
          *loss_part = nullptr;
    if (dalloc(t, &ws1, wgrad_part_bytes(WG_SPLITS)) || dalloc(t, &ws2, wgrad_part_bytes(WG_SPLITS)) ||
        dalloc(t, &film_part, static_cast<size_t>(film_bwd_part_floats(B, F)) * 4) ||
        dalloc(t, &cs_part, static_cast<size_t>(colsum_parts(M)) * F * 4) || dalloc(t, &cs_g, F * 4) ||
        dalloc(t, &thin_part, static_cast<size_t>(B) * 8 * 2 * 9 * F * 4) || dalloc(t, &loss_part, static_cast<size_t>(loss_parts()) * 4))
        return 1;
    // weight gradient of one conv use: dW partials of conv(xin) given its output gradient g, into workspace ws
    auto wgrad = [&](const std::string& tag, const bf16* g, const bf16* xin, float* ws) -> bool {
        WgradLaunch wl;
        char e[256];
        if (wgrad_prepare(g, xin, B, S, S, F, WG_SPLITS, ws, &wl, e, sizeof(e))) { tfail("%s: %s", tag.c_str(), e); return false; }
        b.push("wgrad", tag, [wl](cudaStream_t s) { return wgrad_run(wl, s); }, 2.0 * M * F * 9.0 * F);
        return true;
    };

    // ---------------------------------------------------------------- per-step weight preparation (parameters move every step)
    for (int i = 0; i <= nb; ++i) {
        const float* w = i < nb ? cw[i]->w : bt_w->w;
        bf16 *qf = wqf[i], *qd = wqd[i];
        b.push("prep", "prep.conv." + std::to_string(i), [=](cudaStream_t s) {
            cudaError_t e = prep_conv_weight_run(w, qf, F, F, 3, 0, 1e-5f, F, s);
            return e != cudaSuccess ? e : prep_dgrad_weight_run(w, qd, F, F, s);
        });
    }
    {
        const float *w = tail_w->w, *bias = tail_b->w;
        b.push("prep", "prep.tail", [=](cudaStream_t s) {
            cudaError_t e = prep_conv_weight_run(w, wq_tail, 1, F, 3, 0, 1e-5f, 16, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(tail_bias16, bias, 4, cudaMemcpyDeviceToDevice, s);
            return e != cudaSuccess ? e : flip_tail_weight_run(w, tail_flip, F, s);
        });
    }

    // ---------------------------------------------------------------- forward: time embedding -> FiLM rows
    {
        float* tv = t->time;
        b.push("time", "posenc", [=](cudaStream_t s) { return posenc_rows_run(tv, posenc, B, F, sr3 ? 1 : 0, s); });
        const float *w1d = w1->w, *b1d = b1->w, *w3d = w3->w, *b3d = b3->w;
        b.push("time", "time_mlp.1", [=](cudaStream_t s) { return linear_rows_run(posenc, F, w1d, b1d, z1, TIME_DIM, 0, B, F, TIME_DIM, 0, 0, s); });
        // activations are applied ONCE into their own buffers (g1 = GELU(z1), stemb = SiLU(temb)): the linears and their
        // backward would otherwise re-evaluate them once per output feature
        b.push("time", "gelu+time_mlp.3+silu", [=](cudaStream_t s) {
            cudaError_t e = act_apply_run(z1, g1, static_cast<long long>(B) * TIME_DIM, 2, s);
            if (e == cudaSuccess) e = linear_rows_run(g1, TIME_DIM, w3d, b3d, temb, TIME_DIM, 0, B, TIME_DIM, TIME_DIM, 0, 0, s);
            return e != cudaSuccess ? e : act_apply_run(temb, stemb, static_cast<long long>(B) * TIME_DIM, 1, s);
        });
        for (int i = 0; i < nb; ++i) {
            const float *w = mw[i]->w, *bias = mb[i]->w;
            const int off = i * fw;
            const float* tin = sr3 ? temb : stemb;
            b.push("time", "body." + std::to_string(i) + ".mlp", [=](cudaStream_t s) {
                return linear_rows_run(tin, TIME_DIM, w, bias, film, ld, off, B, TIME_DIM, fw, 0, 0, s);
            });
        }
    }
    // ---------------------------------------------------------------- forward: eps-net
    {
        StemConvArgs a;
        a.x0 = t->cfg.self_condition ? t->cond : t->x;
        a.x1 = t->cfg.self_condition ? t->x : nullptr;
        a.w = head_w->w; a.bias = head_b->w; a.y = X[0];
        a.B = B; a.H = S; a.W = S; a.Cout = F; a.Cin = cin; a.ksize = 3;
        b.push("stem_conv", "head", [a](cudaStream_t s) { return stem_conv_run(a, s); });
    }
    for (int i = 0; i < nb; ++i) {
        const std::string pre = "body." + std::to_string(i);
        ConvEpilogue e1;
        e1.bias = cb[i]->w;
        if (!b.conv(pre + ".conv#1", X[i], wqf[i], F, A[i], e1)) return 1;
        const bf16* ai = A[i];
        bf16* si = Sx[i];
        const int off = i * fw;
        b.push("film_silu_fwd", pre + ".film_silu", [=](cudaStream_t s) { return film_silu_fwd_run(ai, si, film, ld, off, B, P, F, hs, s); });
        ConvEpilogue e2;
        e2.bias = cb[i]->w; e2.out_scale = 0.1f; e2.res = X[i]; e2.ldr = F;
        if (!b.conv(pre + ".conv#2", Sx[i], wqf[i], F, X[i + 1], e2)) return 1;
    }
    {
        ConvEpilogue et;
        et.bias = bt_b->w; et.res = X[0]; et.ldr = F;
        if (!b.conv("body_tail", X[nb], wqf[nb], F, BT, et)) return 1;
        ConvEpilogue eo;
        eo.bias = tail_bias16; eo.out_f32 = t->eps; eo.n_valid = 1;
        if (!b.conv("tail", BT, wq_tail, 16, nullptr, eo)) return 1;
    }
    // ---------------------------------------------------------------- loss and d loss / d eps
    {
        hd_trainer* tt = t;
        b.push("loss", "loss", [=](cudaStream_t s) {
            return loss_grad_run(tt->eps, tt->target, tt->weight, tt->loss_kind, B, P, tt->d_eps, loss_part, tt->loss, s);
        });
    }
    // ---------------------------------------------------------------- backward: tail (256 -> 1)
    {
        const float* de = t->d_eps;
        float *gw = tail_w->g, *gb = tail_b->g;
        b.push("thin_wgrad", "tail.wgrad", [=](cudaStream_t s) { return thin_wgrad_run(BT, de, nullptr, -1, B, thin_part, gw, s); });
        b.push("reduce", "tail.bias_grad", [=](cudaStream_t s) {
            return sum_f32_run(de, M, loss_part, gb, s);
        });
        StemConvArgs a;
        a.x0 = de; a.x1 = nullptr; a.w = tail_flip; a.bias = zero_bias; a.y = DBT;
        a.B = B; a.H = S; a.W = S; a.Cout = F; a.Cin = 1; a.ksize = 3;
        b.push("stem_conv", "tail.dgrad", [a](cudaStream_t s) { return stem_conv_run(a, s); });
    }
    // ---------------------------------------------------------------- backward: body_tail
    {
        float *gw = bt_w->g, *gb = bt_b->g;
        b.push("colsum", "body_tail.bias_grad", [=](cudaStream_t s) { return colsum_run(DBT, M, F, cs_part, 1.0f, 0, gb, s); });
        if (!wgrad("body_tail.wgrad", DBT, X[nb], ws1)) return 1;
        b.push("reduce", "body_tail.wgrad_reduce", [=](cudaStream_t s) { return wgrad_reduce_run(ws1, WG_SPLITS, 1.0f, 0, gw, s); });
        if (!b.conv("body_tail.dgrad", DBT, wqd[nb], F, GX[0], ConvEpilogue())) return 1;
    }
    // ---------------------------------------------------------------- backward: residual blocks
    int cur = 0;
    for (int i = nb - 1; i >= 0; --i) {
        const std::string pre = "body." + std::to_string(i);
        const bf16* g = GX[cur];
        bf16* gnext = GX[cur ^ 1];
        const bf16 *si = Sx[i], *ai = A[i], *xi = X[i];
        const int off = i * fw;
        float *gw = cw[i]->g, *gb = cb[i]->g;
        // graph schedule: the HBM-bound reductions of a block (column sums of g, bias gradient, split-K reduction of dW) run on
        // the side stream next to the tensor-bound GEMMs; the block's first GEMM waits for the previous block's side work (it
        // overwrites ws2, and the next dgrad#1 overwrites g)
        b.push("colsum", pre + ".colsum_g", [=](cudaStream_t s) { return colsum_run(g, M, F, cs_part, 1.0f, 0, cs_g, s); });
        t->ops.back().side = 1;
        if (!wgrad(pre + ".wgrad#2", g, si, ws2)) return 1;
        t->ops.back().join = 1;
        ConvEpilogue e2;
        e2.out_scale = 0.1f;
        if (!b.conv(pre + ".dgrad#2", g, wqd[i], F, DS, e2)) return 1;
        b.push("film_silu_bwd", pre + ".film_silu_bwd", [=](cudaStream_t s) {
            return film_silu_bwd_run(DS, ai, DS, film, dfilm, ld, off, B, P, F, film_part, hs, s);
        });
        b.push("reduce", pre + ".bias_grad", [=](cudaStream_t s) { return edrn_bias_grad_run(film, dfilm, ld, off, B, cs_g, 0.1f, gb, F, hs, s); });
        t->ops.back().side = 1;
        if (!wgrad(pre + ".wgrad#1", DS, xi, ws1)) return 1;
        b.push("reduce", pre + ".wgrad_reduce", [=](cudaStream_t s) {
            cudaError_t e = wgrad_reduce_run(ws1, WG_SPLITS, 1.0f, 0, gw, s);
            return e != cudaSuccess ? e : wgrad_reduce_run(ws2, WG_SPLITS, 0.1f, 1, gw, s);
        });
        t->ops.back().side = 1;
        ConvEpilogue e1;
        e1.res = g; e1.ldr = F;
        if (!b.conv(pre + ".dgrad#1", DS, wqd[i], F, gnext, e1)) return 1;
        cur ^= 1;
    }
    // ---------------------------------------------------------------- backward: head (r = head output feeds block 0 AND body_tail's skip)
    {
        const bf16* g = GX[cur];
        bf16* g0 = GX[cur ^ 1];
        float *gw = head_w->g, *gb = head_b->g;
        const float* u0 = t->cfg.self_condition ? t->cond : t->x;
        const float* u1 = t->cfg.self_condition ? t->x : nullptr;
        b.push("pointwise", "head.grad_sum", [=](cudaStream_t s) { return add_bf16_run(g, DBT, g0, static_cast<long long>(act_elems), s); });
        t->ops.back().join = 1;     // the last block's side work shares cs_part / g with what follows
        b.push("colsum", "head.bias_grad", [=](cudaStream_t s) { return colsum_run(g0, M, F, cs_part, 1.0f, 0, gb, s); });
        b.push("thin_wgrad", "head.wgrad", [=](cudaStream_t s) { return thin_wgrad_run(g0, u0, u1, +1, B, thin_part, gw, s); });
    }
    // ---------------------------------------------------------------- backward: time-embedding MLPs
    {
        std::vector<LinSlot> hs_slots(nb);
        for (int i = 0; i < nb; ++i) hs_slots[i] = LinSlot{mw[i]->w, mw[i]->g, mb[i]->g, i * fw, fw};
        LinSlot* slots = nullptr;
        float* lpart = nullptr;
        if (dalloc(t, &slots, sizeof(LinSlot) * nb) || dalloc(t, &lpart, static_cast<size_t>(nb) * B * TIME_DIM * 4)) return 1;
        if (cudaMemcpy(slots, hs_slots.data(), sizeof(LinSlot) * nb, cudaMemcpyHostToDevice) != cudaSuccess) return tfail("slot upload failed");
        const float* tin = sr3 ? temb : stemb;
        b.push("time_bwd", "body.*.mlp.bwd", [=](cudaStream_t s) {
            cudaError_t e = linear_bwd_weight_batched_run(dfilm, ld, tin, TIME_DIM, B, TIME_DIM, slots, nb, fw, s);
            return e != cudaSuccess ? e : linear_bwd_input_batched_run(dfilm, ld, B, TIME_DIM, slots, nb, lpart, d_act, s);
        });
    }
    {
        float *gw3 = w3->g, *gb3 = b3->g, *gw1 = w1->g, *gb1 = b1->g;
        const float* w3d = w3->w;
        b.push("time_bwd", "time_mlp.bwd", [=](cudaStream_t s) {
            cudaError_t e = sr3 ? cudaSuccess : act_grad_run(d_act, temb, static_cast<long long>(B) * TIME_DIM, 1, s);   // through SiLU(temb)
            if (e == cudaSuccess) e = linear_bwd_weight_run(d_act, TIME_DIM, 0, g1, TIME_DIM, B, TIME_DIM, TIME_DIM, 0, gw3, gb3, s);
            if (e == cudaSuccess) e = linear_bwd_input_run(d_act, TIME_DIM, 0, w3d, B, TIME_DIM, TIME_DIM, 0, d_g1, TIME_DIM, s);
            if (e == cudaSuccess) e = act_grad_run(d_g1, z1, static_cast<long long>(B) * TIME_DIM, 2, s);              // through GELU(z1)
            if (e == cudaSuccess) e = linear_bwd_weight_run(d_g1, TIME_DIM, 0, posenc, F, B, F, TIME_DIM, 0, gw1, gb1, s);
            return e;
        });
    }
    return b.ok ? 0 : 1;
}

}  // namespace

extern "C" {

int hd_trainer_create(const hd_config* cfg, int32_t batch, hd_trainer** out) {
    if (!cfg || !out) return tfail("hd_trainer_create: null argument");
    if (cfg->abi_version != HD_ABI_VERSION) return tfail("ABI version mismatch: header %d, library %d", cfg->abi_version, HD_ABI_VERSION);
    if (cfg->variant < HD_UNET || cfg->variant > HD_HICEDRN_SR3) return tfail("unknown variant %d", cfg->variant);
    if (cfg->image_size != 64) return tfail("image_size must be 64");
    if (cfg->variant == HD_UNET || cfg->variant == HD_UNET_SR3) {
        if (cfg->dim != 64) return tfail("the Unet training step is built for dim = 64 (got %d)", cfg->dim);
        if (cfg->num_mults < 1 || cfg->num_mults > 4) return tfail("len(dim_mults) must be 1..4 for 64x64 tiles (got %d)", cfg->num_mults);
        for (int i = 0; i < cfg->num_mults; ++i)
            if (cfg->dim_mults[i] != 1 && cfg->dim_mults[i] != 2 && cfg->dim_mults[i] != 4 && cfg->dim_mults[i] != 8)
                return tfail("dim_mults[%d] = %d: the training kernels cover 64 / 128 / 256 / 512 channels", i, cfg->dim_mults[i]);
    } else if (cfg->num_blocks < 1) {
        return tfail("HiCEDRN needs num_blocks >= 1");
    }
    if (batch < 1) return tfail("batch must be positive (got %d)", batch);
    int dev = 0, cc_major = 0, sms = 0;
    T_TRY(cudaGetDevice(&dev));
    T_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    T_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major != 10) return tfail("hicdiff_b200 needs an sm_100a GPU (compute capability 10.x); device %d is %d.x", dev, cc_major);
    hd_trainer* t = new hd_trainer();
    t->cfg = *cfg;
    t->device = dev;
    t->num_sms = sms;
    t->B = batch;
    t->nb = cfg->num_blocks;
    *out = t;
    return 0;
}

int hd_trainer_bind(hd_trainer* t, const char* key, const float* param, float* grad, const int64_t* shape, int32_t ndim) {
    if (!t || !key || !param || !grad || (ndim > 0 && !shape)) return tfail("hd_trainer_bind: null argument");
    if (t->finalized) return tfail("hd_trainer_bind: the trainer is already finalized");
    TParam p;
    p.w = param; p.g = grad; p.numel = 1;
    for (int i = 0; i < ndim; ++i) { p.shape.push_back(shape[i]); p.numel *= static_cast<size_t>(shape[i]); }
    t->p[key] = p;
    return 0;
}

int hd_trainer_finalize(hd_trainer* t, void* stream) {
    NvtxRange range("hd_trainer_finalize");
    if (!t) return tfail("hd_trainer_finalize: null trainer");
    if (t->finalized) return 0;
    T_TRY(cudaSetDevice(t->device));
    if ((t->cfg.variant == HD_UNET || t->cfg.variant == HD_UNET_SR3) ? build_unet_trainer(t) : build(t)) return 1;
    T_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    t->finalized = true;
    return 0;
}

int hd_trainer_step(hd_trainer* t, const float* x_t, const float* cond, const float* time, const float* target,
                    const float* weight, int32_t loss_type, float* eps_out, float* loss_out, void* stream) {
    NvtxRange range("hd_trainer_step (forward + loss + backward)");
    if (!t || !x_t || !time || !target || !weight || !loss_out) return tfail("hd_trainer_step: null argument");
    if (!t->finalized) return tfail("hd_trainer_finalize has not been called");
    if (t->cfg.self_condition && !cond) return tfail("hd_trainer_step: the net is self-conditioned, cond must not be NULL");
    if (loss_type != 0 && loss_type != 1) return tfail("loss_type must be 0 (l1) or 1 (l2)");
    T_TRY(cudaSetDevice(t->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t tb = static_cast<size_t>(t->B) * P * 4;
    T_TRY(cudaMemcpyAsync(t->x, x_t, tb, cudaMemcpyDeviceToDevice, s));
    if (cond) T_TRY(cudaMemcpyAsync(t->cond, cond, tb, cudaMemcpyDeviceToDevice, s));
    T_TRY(cudaMemcpyAsync(t->target, target, tb, cudaMemcpyDeviceToDevice, s));
    T_TRY(cudaMemcpyAsync(t->time, time, t->B * 4, cudaMemcpyDeviceToDevice, s));
    T_TRY(cudaMemcpyAsync(t->weight, weight, t->B * 4, cudaMemcpyDeviceToDevice, s));
    t->loss_kind = loss_type;
    PdlScope pdl;     // the forward's kernels (conv_gemm, GroupNorm, attention ...) become programmatic dependents inside the step
    static const bool use_graph = [] { const char* v = getenv("HD_TRAIN_GRAPH"); return !(v && v[0] == '0'); }();
    if (use_graph && t->graph[loss_type] == nullptr && t->eager_steps[loss_type] >= 1) {
        // capture (nothing executes) on the private stream: the caller's stream may be the legacy default stream
        if (!t->cap_stream) {
            T_TRY(cudaStreamCreateWithFlags(&t->cap_stream, cudaStreamNonBlocking));
            T_TRY(cudaStreamCreateWithFlags(&t->side_stream, cudaStreamNonBlocking));
            T_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
            T_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
        }
        T_TRY(cudaStreamBeginCapture(t->cap_stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t le = cudaSuccess;
        const char* bad = "";
        bool side_dirty = false;
        auto join_side = [&]() -> cudaError_t {
            if (!side_dirty) return cudaSuccess;
            side_dirty = false;
            cudaError_t e = cudaEventRecord(t->ev_join, t->side_stream);
            return e != cudaSuccess ? e : cudaStreamWaitEvent(t->cap_stream, t->ev_join, 0);
        };
        for (const TOp& op : t->ops) {
            if (op.side) {
                le = cudaEventRecord(t->ev_fork, t->cap_stream);
                if (le == cudaSuccess) le = cudaStreamWaitEvent(t->side_stream, t->ev_fork, 0);
                if (le == cudaSuccess) le = op.fn(t->side_stream);
                side_dirty = true;
            } else {
                if (op.join) le = join_side();
                if (le == cudaSuccess) le = op.fn(t->cap_stream);
                if (le == cudaSuccess && op.bucket >= 0)      // visible to streams outside the graph: an external record node
                    le = cudaEventRecordWithFlags(t->bucket_ev[op.bucket], t->cap_stream, cudaEventRecordExternal);
            }
            if (le != cudaSuccess) { bad = op.tag.c_str(); break; }
        }
        if (le == cudaSuccess) le = join_side();
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(t->cap_stream, &g);
        if (le == cudaSuccess && ce == cudaSuccess) ce = cudaGraphInstantiate(&t->graph[loss_type], g, 0);
        if (g) cudaGraphDestroy(g);
        if (le != cudaSuccess || ce != cudaSuccess) {
            t->graph[loss_type] = nullptr;
            return tfail("capturing the training step failed at '%s': %s", bad, cudaGetErrorString(le != cudaSuccess ? le : ce));
        }
    }
    if (use_graph && t->graph[loss_type] != nullptr) {
        T_TRY(cudaGraphLaunch(t->graph[loss_type], s));
    } else {
        for (const TOp& op : t->ops) {
            cudaError_t e = op.fn(s);
            if (e == cudaSuccess && op.bucket >= 0) e = cudaEventRecord(t->bucket_ev[op.bucket], s);
            if (e != cudaSuccess) return tfail("launch of '%s' failed: %s", op.tag.c_str(), cudaGetErrorString(e));
        }
        t->eager_steps[loss_type] += 1;
    }
    if (eps_out) T_TRY(cudaMemcpyAsync(eps_out, t->eps, tb, cudaMemcpyDeviceToDevice, s));
    T_TRY(cudaMemcpyAsync(loss_out, t->loss, 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int hd_trainer_set_grad_buckets(hd_trainer* t, const char* const* first_module_prefix, int32_t n) {
    if (!t || n < 0 || (n > 0 && !first_module_prefix)) return tfail("hd_trainer_set_grad_buckets: bad argument");
    if (t->finalized) return tfail("hd_trainer_set_grad_buckets must be called before hd_trainer_finalize");
    if (t->cfg.variant != HD_UNET && t->cfg.variant != HD_UNET_SR3) return tfail("gradient buckets are built for the Unet trainers");
    T_TRY(cudaSetDevice(t->device));
    for (cudaEvent_t e : t->bucket_ev) cudaEventDestroy(e);
    t->bucket_ev.clear();
    t->bucket_prefix.clear();
    for (int i = 0; i < n; ++i) {
        if (!first_module_prefix[i] || !first_module_prefix[i][0]) return tfail("hd_trainer_set_grad_buckets: empty prefix %d", i);
        cudaEvent_t e = nullptr;
        T_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        t->bucket_ev.push_back(e);
        t->bucket_prefix.push_back(first_module_prefix[i]);
    }
    return 0;
}

int hd_trainer_wait_grad_bucket(hd_trainer* t, int32_t bucket, void* stream) {
    if (!t || bucket < 0 || bucket >= static_cast<int32_t>(t->bucket_ev.size())) return tfail("hd_trainer_wait_grad_bucket: no bucket %d", bucket);
    T_TRY(cudaSetDevice(t->device));
    T_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), t->bucket_ev[bucket], 0));
    return 0;
}

int hd_trainer_num_launches(hd_trainer* t) { return t ? static_cast<int>(t->ops.size()) : 0; }

int hd_trainer_profile(hd_trainer* t, int32_t reps, char* buf, int64_t buflen, void* stream) {
    if (!t || !buf || buflen <= 0 || reps <= 0) return tfail("hd_trainer_profile: bad argument");
    if (!t->finalized) return tfail("hd_trainer_finalize has not been called");
    T_TRY(cudaSetDevice(t->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaEvent_t e0, e1;
    T_TRY(cudaEventCreate(&e0));
    T_TRY(cudaEventCreate(&e1));
    std::map<std::string, std::pair<double, double>> fam;   // kernel family -> (ms, flops)
    std::map<std::string, int> cnt;
    std::vector<std::pair<double, std::string>> slow;        // (ms, tag) of every op, for the "top" list
    for (const TOp& op : t->ops) {
        cudaError_t e = op.fn(s);
        if (e == cudaSuccess) e = cudaEventRecord(e0, s);
        for (int r = 0; r < reps && e == cudaSuccess; ++r) e = op.fn(s);
        if (e == cudaSuccess) e = cudaEventRecord(e1, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return tfail("profiling '%s' failed: %s", op.tag.c_str(), cudaGetErrorString(e));
        }
        slow.push_back({ms / reps, op.tag});
        fam[op.kernel].first += ms / reps;
        fam[op.kernel].second += op.flops;
        cnt[op.kernel] += 1;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    std::string out = "{";
    bool first = true;
    for (auto& kv : fam) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\":{\"ms\":%.6f,\"flops\":%.6e,\"ops\":%d}", first ? "" : ",", kv.first.c_str(),
                 kv.second.first, kv.second.second, cnt[kv.first]);
        out += line;
        first = false;
    }
    std::sort(slow.begin(), slow.end(), [](const std::pair<double, std::string>& a, const std::pair<double, std::string>& b) { return a.first > b.first; });
    out += ",\"_top\":{";
    for (size_t i = 0; i < slow.size() && i < 24; ++i) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\":{\"ms\":%.6f,\"flops\":0,\"ops\":1}", i ? "," : "", slow[i].second.c_str(), slow[i].first);
        out += line;
    }
    out += "}}";
    if (static_cast<int64_t>(out.size()) + 1 > buflen) return tfail("hd_trainer_profile: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

int64_t hd_trainer_device_bytes(hd_trainer* t) { return t ? static_cast<int64_t>(t->bytes) : 0; }

void hd_trainer_destroy(hd_trainer* t) {
    if (!t) return;
    for (cudaEvent_t e : t->bucket_ev) cudaEventDestroy(e);
    cudaSetDevice(t->device);
    for (int k = 0; k < 2; ++k) if (t->graph[k]) cudaGraphExecDestroy(t->graph[k]);
    if (t->cap_stream) cudaStreamDestroy(t->cap_stream);
    if (t->side_stream) cudaStreamDestroy(t->side_stream);
    if (t->ev_fork) cudaEventDestroy(t->ev_fork);
    if (t->ev_join) cudaEventDestroy(t->ev_join);
    for (void* q : t->allocs) cudaFree(q);
    delete t;
}

// ---- single-operator entry points of the training path (parity tests pin each kernel against torch.autograd on the CPU)
int hd_op_conv3x3_wgrad(const uint16_t* x, const uint16_t* dy, float* dw, int32_t B, void* stream) {
    if (!x || !dy || !dw || B < 1) return tfail("hd_op_conv3x3_wgrad: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* ws = nullptr;
    T_TRY(cudaMalloc(&ws, wgrad_part_bytes(WG_SPLITS)));
    WgradLaunch wl;
    char e[256];
    int rc = wgrad_prepare(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), B, S, S, F, WG_SPLITS, ws, &wl, e, sizeof(e));
    cudaError_t ce = cudaSuccess;
    if (!rc) ce = wgrad_run(wl, s);
    if (!rc && ce == cudaSuccess) ce = wgrad_reduce_run(ws, WG_SPLITS, 1.0f, 0, dw, s);
    if (!rc && ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(ws);
    if (rc) return tfail("%s", e);
    if (ce != cudaSuccess) return tfail("hd_op_conv3x3_wgrad failed: %s", cudaGetErrorString(ce));
    return 0;
}

int hd_op_groupnorm_silu_bwd(const uint16_t* y, const uint16_t* ds, const float* gamma, const float* beta, const float* scale,
                             const float* shift, uint16_t* dy, float* dgamma, float* dbeta, float* dscale, float* dshift,
                             float* dconv_bias, int32_t B, int32_t P, int32_t C, void* stream) {
    if (!y || !ds || !gamma || !beta || !dy || !dgamma || !dbeta || B < 1) return tfail("hd_op_groupnorm_silu_bwd: bad argument");
    if ((scale == nullptr) != (shift == nullptr)) return tfail("hd_op_groupnorm_silu_bwd: scale and shift come together");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* scratch = nullptr;
    T_TRY(cudaMalloc(&scratch, gn_bwd_scratch_floats(B, P, C) * sizeof(float)));
    GroupNormBwdArgs a;
    a.y = reinterpret_cast<const bf16*>(y); a.ds = reinterpret_cast<const bf16*>(ds); a.dy = reinterpret_cast<bf16*>(dy);
    a.B = B; a.P = P; a.C = C; a.gamma = gamma; a.beta = beta; a.eps = 1e-5f; a.scale = scale; a.shift = shift;
    a.dgamma = dgamma; a.dbeta = dbeta; a.dscale = dscale; a.dshift = dshift; a.dconv_bias = dconv_bias;
    cudaError_t ce = groupnorm_silu_bwd_run(a, scratch, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (ce != cudaSuccess) return tfail("hd_op_groupnorm_silu_bwd failed: %s", cudaGetErrorString(ce));
    return 0;
}

int hd_op_channel_layernorm_bwd(const uint16_t* x, const uint16_t* dz, const float* g, uint16_t* dx, float* dg, int64_t M, int32_t C,
                                void* stream) {
    if (!x || !dz || !g || !dx || !dg || M < 1) return tfail("hd_op_channel_layernorm_bwd: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* part = nullptr;
    T_TRY(cudaMalloc(&part, static_cast<size_t>(ln_bwd_blocks(M)) * C * sizeof(float)));
    cudaError_t ce = channel_layernorm_bwd_run(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(dz), g, M, C, 1e-5f,
                                               reinterpret_cast<bf16*>(dx), dg, part, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(part);
    if (ce != cudaSuccess) return tfail("hd_op_channel_layernorm_bwd failed: %s", cudaGetErrorString(ce));
    return 0;
}

int hd_op_weight_standardize_bwd(const float* w, const float* dwt, float* dw, int32_t Cout, int32_t K, void* stream) {
    if (!w || !dwt || !dw || Cout < 1 || K < 1) return tfail("hd_op_weight_standardize_bwd: bad argument");
    T_TRY(weight_standardize_bwd_run(w, dwt, Cout, K, 1e-5f, dw, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_op_attention_bwd(const uint16_t* qkv, const uint16_t* dout, uint16_t* dqkv, int32_t B, int32_t n, int32_t linear, void* stream) {
    if (!qkv || !dout || !dqkv || B < 1 || n < 1) return tfail("hd_op_attention_bwd: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t ce;
    if (linear) {
        float* scratch = nullptr;
        T_TRY(cudaMalloc(&scratch, linattn_bwd_scratch_floats(B) * sizeof(float)));
        ce = linear_attention_bwd_run(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dqkv), B, n, scratch, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        cudaFree(scratch);
    } else {
        if (n != 64) return tfail("hd_op_attention_bwd: the softmax attention runs at 8x8 only (n = 64), got n = %d", n);
        ce = full_attention_bwd_run(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(dout), reinterpret_cast<bf16*>(dqkv), B, n, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    }
    if (ce != cudaSuccess) return tfail("hd_op_attention_bwd failed: %s", cudaGetErrorString(ce));
    return 0;
}

int hd_op_conv_wgrad(const uint16_t* x, const uint16_t* dy, float* dw, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                     int32_t ksize, int32_t cin_total, int32_t ci0, void* stream) {
    if (!x || !dy || !dw || B < 1) return tfail("hd_op_conv_wgrad: bad argument");
    if (ci0 < 0 || ci0 + Cin > cin_total) return tfail("hd_op_conv_wgrad: channel slice [%d, %d) outside [0, %d)", ci0, ci0 + Cin, cin_total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgradGenLaunch wl;
    char e[256];
    if (wgrad_general_prepare(reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), B, H, W, Cout, Cin, ksize, sms, &wl, e, sizeof(e)))
        return tfail("%s", e);
    float* ws = nullptr;
    T_TRY(cudaMalloc(&ws, wgrad_general_part_bytes(wl)));
    wl.part = ws;
    cudaError_t ce = wgrad_general_run(wl, s);
    if (ce == cudaSuccess) ce = wgrad_general_reduce_run(wl, cin_total, ci0, 1.0f, 0, dw, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(ws);
    if (ce != cudaSuccess) return tfail("hd_op_conv_wgrad failed: %s", cudaGetErrorString(ce));
    return 0;
}

int hd_op_conv3x3_dgrad(const uint16_t* dy, const float* w, uint16_t* dx, int32_t B, void* stream) {
    if (!dy || !w || !dx || B < 1) return tfail("hd_op_conv3x3_dgrad: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bf16* wq = nullptr;
    T_TRY(cudaMalloc(&wq, static_cast<size_t>(F) * 9 * F * 2));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ConvGemmDesc d;
    d.src0 = ConvSrc{reinterpret_cast<const bf16*>(dy), F};
    d.src1 = ConvSrc{nullptr, 0};
    d.B = B; d.H = S; d.W = S; d.ksize = 3; d.mode = CONV_TAPS;
    d.weight = wq; d.N = F; d.out = reinterpret_cast<bf16*>(dx);
    ConvGemmLaunch l;
    char e[256];
    int rc = conv_gemm_prepare(d, sms, &l, e, sizeof(e));
    cudaError_t ce = cudaSuccess;
    if (!rc) ce = prep_dgrad_weight_run(w, wq, F, F, s);
    if (!rc && ce == cudaSuccess) ce = conv_gemm_run(l, s);
    if (!rc && ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(wq);
    if (rc) return tfail("%s", e);
    if (ce != cudaSuccess) return tfail("hd_op_conv3x3_dgrad failed: %s", cudaGetErrorString(ce));
    return 0;
}

}  // extern "C"
