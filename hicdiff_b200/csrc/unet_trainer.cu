// Training step of the Unet eps-net (BASELINE config 5: conditional Unet p_losses + backward).
//   forward   Unet.forward                 /root/reference/src/hicdiff_condition.py:345-384 (ResnetBlock :173-197, Block :155-171,
//                                           LinearAttention :199-227, Attention :229-251, PreNorm / LayerNorm / Residual :64-118,
//                                           Downsample / Upsample :72-82, WeightStandardizedConv2d :84-97, time_mlp :300-305)
//   backward  torch.autograd over the above (pretrain/train_unet_Diff*.py: loss = diffusion(x); loss.backward())
// The forward is the sampling path's op list with every intermediate kept (bf16 NHWC, no arena reuse) and the two resampling
// layers materialised; the backward walks the same list in reverse.  Convs: conv_gemm.cu (forward and data gradient over the
// flipped / transposed, standardised filter) and wgrad.cu (general kernel); everything else: norm_bwd.cu, attention_bwd.cu,
// train_kernels.cu, unet_train_kernels.cu.  Every gradient tensor has its own buffer; a tensor with several consumers
// (skip connections, residual branches) accumulates through add_bf16 in a fixed order.
#include <cstdlib>
#include <cstring>
#include <memory>

#include "trainer_internal.h"

namespace hd {
namespace {

constexpr int S = 64;
constexpr float EPS = 1e-5f;

struct Ten {           // an activation and (lazily) its gradient
    bf16* p = nullptr;
    bf16* g = nullptr;
    int H = 0, C = 0;  // square H x H, C channels
    bool gset = false; // the gradient buffer already holds a contribution
};
using TenP = std::shared_ptr<Ten>;

struct ConvW {         // one conv's parameters and per-step prepared layouts
    const TParam *w = nullptr, *b = nullptr;
    int Cout = 0, Cin = 0, k = 1;
    bool ws = false;   // weight-standardised
    bf16* qf = nullptr;    // forward GEMM layout [Cout][taps * Cin]
    bf16* qd = nullptr;    // dgrad layout [Cin][taps * Cout]
    float2* stats = nullptr;
};

struct UB {
    hd_trainer* t;
    int B;
    bool ok = true;
    bool sr3 = false;                 // hicdiff_sr3.Unet: additive noise-level embedding after block1 instead of FiLM
    int ld = 0;                       // FiLM row width
    float *film = nullptr, *dfilm = nullptr;
    int* iota = nullptr;
    float2* gn_part = nullptr;
    float *gn_scratch = nullptr, *la_scratch = nullptr, *ln_part = nullptr, *cs_part = nullptr, *la_ctx = nullptr, *zero_bias = nullptr;
    bf16 *T1 = nullptr, *T2 = nullptr, *T3 = nullptr;    // gradient temporaries (largest activation)
    std::vector<std::function<void()>> bwd;              // backward emitters, run in reverse
    std::vector<std::string> bwd_name;                   // module prefix of each emitter (gradient-bucket boundaries)
    std::vector<std::unique_ptr<ConvW>> convs;           // build-time only: every launch closure captures values, not these
    std::vector<PrepSlot> prep_slots;                    // every conv's weight preparation, uploaded by finish_prep()
    struct PrepTables { PrepSlot* slots = nullptr; int2 *fwd = nullptr, *bwd = nullptr; int nfwd = 0, nbwd = 0; };
    std::shared_ptr<PrepTables> prep = std::make_shared<PrepTables>();

    // Graph schedule (see TOp::side / join): weight-gradient work (wgrad + split reduction, standardisation backward, bias column
    // sums) is a leaf of the backward graph, so it runs on the side stream next to the data-gradient chain; it reads the gradient
    // temporaries T1 / T2 of its own layer, which the NEXT layer's backward overwrites -> the first main-stream op of every
    // backward emitter joins.  On the low-resolution levels neither kind of kernel fills the GPU, so the two overlap.
    bool side_mode = false, pending_join = false;
    void push(const char* kernel, const std::string& tag, std::function<cudaError_t(cudaStream_t)> fn, double flops = 0) {
        TOp o;
        o.fn = std::move(fn); o.tag = tag; o.kernel = kernel; o.flops = flops;
        if (side_mode) {
            o.side = 1;
        } else if (pending_join) {
            o.join = 1;
            pending_join = false;
        }
        t->ops.push_back(std::move(o));
    }
    // the step's first op: all weight layouts in two launches (tables are filled by finish_prep() once every conv is declared)
    void push_prep() {
        std::shared_ptr<PrepTables> pt = prep;
        push("prep", "prep.all", [pt](cudaStream_t s) { return prep_weights_batched_run(pt->slots, pt->fwd, pt->nfwd, pt->bwd, pt->nbwd, EPS, s); });
    }
    bool finish_prep() {
        std::vector<int2> fwd, bwd;
        for (size_t i = 0; i < prep_slots.size(); ++i) {
            for (int r = 0; r < prep_slots[i].Cout; ++r) fwd.push_back(make_int2(static_cast<int>(i), r));
            for (int r = 0; r < prep_slots[i].Cin; ++r) bwd.push_back(make_int2(static_cast<int>(i), r));
        }
        if (prep_slots.empty()) return true;
        if (dalloc(t, &prep->slots, prep_slots.size() * sizeof(PrepSlot)) || dalloc(t, &prep->fwd, fwd.size() * sizeof(int2)) ||
            dalloc(t, &prep->bwd, bwd.size() * sizeof(int2)))
            return false;
        if (cudaMemcpy(prep->slots, prep_slots.data(), prep_slots.size() * sizeof(PrepSlot), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(prep->fwd, fwd.data(), fwd.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(prep->bwd, bwd.data(), bwd.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess)
            return fail("weight preparation table upload failed");
        prep->nfwd = static_cast<int>(fwd.size());
        prep->nbwd = static_cast<int>(bwd.size());
        return true;
    }
    bool fail(const std::string& m) { if (ok) { ok = false; tfail("%s", m.c_str()); } return false; }
    size_t elems(int H, int C) const { return static_cast<size_t>(B) * H * H * C; }
    long long rows(int H) const { return static_cast<long long>(B) * H * H; }

    TenP ten(int H, int C) {
        auto x = std::make_shared<Ten>();
        x->H = H; x->C = C;
        if (dalloc(t, &x->p, elems(H, C) * 2) || dalloc(t, &x->g, elems(H, C) * 2)) ok = false;
        return x;
    }
    ConvW* convw(const std::string& wkey, const std::string& bkey, int Cout, int Cin, int k, bool ws) {
        convs.emplace_back(new ConvW());
        ConvW* c = convs.back().get();
        c->Cout = Cout; c->Cin = Cin; c->k = k; c->ws = ws;
        c->w = find_p(t, wkey, {Cout, Cin, k, k});
        c->b = bkey.empty() ? nullptr : find_p(t, bkey, {Cout});
        if (!c->w || (!bkey.empty() && !c->b)) { ok = false; return c; }
        const size_t n = static_cast<size_t>(Cout) * Cin * k * k;
        if (dalloc(t, &c->qf, n * 2) || dalloc(t, &c->qd, n * 2) || dalloc(t, &c->stats, static_cast<size_t>(Cout) * sizeof(float2))) ok = false;
        const float* w = c->w->w;
        bf16 *qf = c->qf, *qd = c->qd;
        float2* st = c->stats;
        const int wsi = ws ? 1 : 0;
        prep_slots.push_back(PrepSlot{w, qf, qd, st, Cout, Cin, k, wsi});   // prepared by the one batched "prep" op at the head of the step
        return c;
    }

    // ---------------------------------------------------------------- primitive emitters
    // forward conv over one or two sources
    bool conv_fwd(const std::string& tag, const ConvW* c, const Ten& x0, const Ten* x1, bf16* out, ConvEpilogue epi, bool gn_stats) {
        ConvGemmDesc d;
        d.src0 = ConvSrc{x0.p, x0.C};
        d.src1 = ConvSrc{x1 ? x1->p : nullptr, x1 ? x1->C : 0};
        d.B = B; d.H = x0.H; d.W = x0.H; d.ksize = c->k; d.mode = CONV_TAPS;
        d.weight = c->qf; d.N = c->Cout; d.out = out;
        epi.bias = c->b ? c->b->w : zero_bias;
        if (gn_stats) epi.gn_part = gn_part;
        d.epi = epi;
        ConvGemmLaunch l;
        char e[256];
        if (conv_gemm_prepare(d, t->num_sms, &l, e, sizeof(e))) return fail(tag + ": " + e);
        push("conv_gemm", tag, [l](cudaStream_t s) { return conv_gemm_run(l, s); }, 2.0 * rows(x0.H) * c->Cout * c->k * c->k * c->Cin);
        return true;
    }
    // data gradient of conv `c` w.r.t. the source occupying input channels [ci0, ci0 + Cs): out [rows, Cs] (+ res)
    bool conv_dgrad(const std::string& tag, const ConvW* c, const bf16* dy, int H, int ci0, int Cs, bf16* out, const bf16* res) {
        ConvGemmDesc d;
        d.src0 = ConvSrc{dy, c->Cout};
        d.src1 = ConvSrc{nullptr, 0};
        d.B = B; d.H = H; d.W = H; d.ksize = c->k; d.mode = CONV_TAPS;
        d.weight = c->qd + static_cast<size_t>(ci0) * c->k * c->k * c->Cout;
        d.N = Cs; d.out = out;
        ConvEpilogue epi;
        epi.bias = zero_bias;
        epi.res = res; epi.ldr = Cs;
        d.epi = epi;
        ConvGemmLaunch l;
        char e[256];
        if (conv_gemm_prepare(d, t->num_sms, &l, e, sizeof(e))) return fail(tag + ": " + e);
        push("conv_gemm", tag, [l](cudaStream_t s) { return conv_gemm_run(l, s); }, 2.0 * rows(H) * Cs * c->k * c->k * c->Cout);
        return true;
    }
    // weight (+ bias) gradient of conv `c` given dy and its source tensors; finishes with the standardisation backward
    bool conv_wgrad(const std::string& tag, const ConvW* c, const bf16* dy, int H, const Ten& x0, const Ten* x1, bool bias_done = false) {
        struct SideScope { UB* u; explicit SideScope(UB* p) : u(p) { u->side_mode = true; } ~SideScope() { u->side_mode = false; } } scope(this);
        int ci0 = 0;
        const Ten* src[2] = {&x0, x1};
        for (int i = 0; i < 2; ++i) {
            if (!src[i]) continue;
            WgradGenLaunch wl;
            char e[256];
            if (wgrad_general_prepare(dy, src[i]->p, B, H, H, c->Cout, src[i]->C, c->k, t->num_sms, &wl, e, sizeof(e))) return fail(tag + ": " + e);
            if (dalloc(t, &wl.part, wgrad_general_part_bytes(wl))) return ok = false;
            float* gw = c->w->g;
            const int ct = c->Cin, c0 = ci0;
            push("wgrad", tag, [wl, gw, ct, c0](cudaStream_t s) {
                cudaError_t er = wgrad_general_run(wl, s);
                return er != cudaSuccess ? er : wgrad_general_reduce_run(wl, ct, c0, 1.0f, 0, gw, s);
            }, 2.0 * rows(H) * c->Cout * c->k * c->k * src[i]->C);
            ci0 += src[i]->C;
        }
        if (c->ws) {
            const float* w = c->w->w;
            float* gw = c->w->g;
            const int Cout = c->Cout, K = c->Cin * c->k * c->k;
            push("ws_bwd", tag + ".ws", [=](cudaStream_t s) { return weight_standardize_bwd_run(w, gw, Cout, K, EPS, gw, s); });
        }
        if (c->b && !bias_done) {      // (the GroupNorm backward already produced the bias gradient of the convs it follows)
            float* gb = c->b->g;
            float* part = cs_part;
            const long long M = rows(H);
            const int Cout = c->Cout;
            push("colsum", tag + ".bias", [=](cudaStream_t s) { return colsum_run(dy, M, Cout, part, 1.0f, 0, gb, s); });
        }
        return true;
    }
    // fold `src` into x's gradient
    void accumulate(const std::string& tag, Ten& x, const bf16* src) {
        bf16* g = x.g;
        const size_t n = elems(x.H, x.C);
        if (x.gset) push("pointwise", tag + ".acc", [=](cudaStream_t s) { return add_bf16_run(g, src, g, static_cast<long long>(n), s); });
        else push("pointwise", tag + ".set", [=](cudaStream_t s) { return cudaMemcpyAsync(g, src, n * 2, cudaMemcpyDeviceToDevice, s); });
        x.gset = true;
    }
    // where a conv-type producer should write x's gradient: straight into x.g the first time, else into `tmp` + accumulate
    bf16* target(Ten& x, bf16* tmp) { return x.gset ? tmp : x.g; }
    void landed(const std::string& tag, Ten& x, bf16* where) {
        if (where != x.g) accumulate(tag, x, where); else x.gset = true;
    }

    // ---------------------------------------------------------------- ResnetBlock
    TenP resblock(const std::string& p, const TenP& xa, const TenP& xb, int Cout, int film_off) {
        const int H = xa->H, Cin = xa->C + (xb ? xb->C : 0);
        ConvW* c1 = convw(p + ".block1.proj.weight", p + ".block1.proj.bias", Cout, Cin, 3, true);
        ConvW* c2 = convw(p + ".block2.proj.weight", p + ".block2.proj.bias", Cout, Cout, 3, true);
        ConvW* cr = Cin != Cout ? convw(p + ".res_conv.weight", p + ".res_conv.bias", Cout, Cin, 1, false) : nullptr;
        const TParam *g1 = find_p(t, p + ".block1.norm.weight", {Cout}), *b1 = find_p(t, p + ".block1.norm.bias", {Cout});
        const TParam *g2 = find_p(t, p + ".block2.norm.weight", {Cout}), *b2 = find_p(t, p + ".block2.norm.bias", {Cout});
        if (!ok || !g1 || !b1 || !g2 || !b2) { ok = false; return xa; }
        TenP y1 = ten(H, Cout), s1 = ten(H, Cout), y2 = ten(H, Cout), out = ten(H, Cout), rc = cr ? ten(H, Cout) : nullptr;
        float2 *st1 = nullptr, *st2 = nullptr;      // the forward GroupNorms' (mean, rstd), reused by their backward
        if (dalloc(t, &st1, static_cast<size_t>(B) * 8 * sizeof(float2)) || dalloc(t, &st2, static_cast<size_t>(B) * 8 * sizeof(float2))) ok = false;
        if (!ok) return out;
        const int P = H * H;
        if (!conv_fwd(p + ".block1.conv", c1, *xa, xb.get(), y1->p, ConvEpilogue(), true)) return out;
        {
            GroupNormArgs a;
            a.x = y1->p; a.y = s1->p; a.B = B; a.P = P; a.C = Cout; a.part = gn_part; a.gamma = g1->w; a.beta = b1->w; a.eps = EPS;
            a.film_row = iota; a.film_row_stride = 1; a.film_ld = ld;
            if (sr3) { a.postadd = film; a.postadd_off = film_off; } else { a.film = film; a.film_off = film_off; }
            a.stats_out = st1;
            push("groupnorm", p + ".block1.norm", [a](cudaStream_t s) { return groupnorm_film_silu_run(a, s); });
        }
        if (!conv_fwd(p + ".block2.conv", c2, *s1, nullptr, y2->p, ConvEpilogue(), true)) return out;
        if (cr && !conv_fwd(p + ".res_conv", cr, *xa, xb.get(), rc->p, ConvEpilogue(), false)) return out;
        {
            GroupNormArgs a;
            a.x = y2->p; a.y = out->p; a.B = B; a.P = P; a.C = Cout; a.part = gn_part; a.gamma = g2->w; a.beta = b2->w; a.eps = EPS;
            a.res = cr ? rc->p : xa->p;
            a.stats_out = st2;
            push("groupnorm", p + ".block2.norm", [a](cudaStream_t s) { return groupnorm_film_silu_run(a, s); });
        }
        bwd_name.push_back(p);
        bwd.push_back([=]() {
            const bf16* g = out->g;      // d out
            // block2: GroupNorm + SiLU, conv2
            {
                GroupNormBwdArgs a;
                a.y = y2->p; a.ds = g; a.dy = T1; a.B = B; a.P = P; a.C = Cout; a.gamma = g2->w; a.beta = b2->w; a.eps = EPS;
                a.dgamma = g2->g; a.dbeta = b2->g; a.dconv_bias = c2->b->g; a.stats_in = st2;
                float* sc = gn_scratch;
                push("groupnorm_bwd", p + ".block2.norm.bwd", [a, sc](cudaStream_t s) { return groupnorm_silu_bwd_run(a, sc, s); });
            }
            if (!conv_wgrad(p + ".block2.wgrad", c2, T1, H, *s1, nullptr, true)) return;
            if (!conv_dgrad(p + ".block2.dgrad", c2, T1, H, 0, Cout, T2, nullptr)) return;
            // block1: GroupNorm + FiLM + SiLU (in place on T2), conv1
            {
                GroupNormBwdArgs a;
                a.y = y1->p; a.ds = T2; a.dy = T2; a.B = B; a.P = P; a.C = Cout; a.gamma = g1->w; a.beta = b1->w; a.eps = EPS;
                a.ld = ld;
                if (sr3) { a.dpost = dfilm + film_off; }
                else { a.scale = film + film_off; a.shift = film + film_off + Cout; a.dscale = dfilm + film_off; a.dshift = dfilm + film_off + Cout; }
                a.dgamma = g1->g; a.dbeta = b1->g; a.dconv_bias = c1->b->g; a.stats_in = st1;
                float* sc = gn_scratch;
                push("groupnorm_bwd", p + ".block1.norm.bwd", [a, sc](cudaStream_t s) { return groupnorm_silu_bwd_run(a, sc, s); });
            }
            if (!conv_wgrad(p + ".block1.wgrad", c1, T2, H, *xa, xb.get(), true)) return;
            if (cr && !conv_wgrad(p + ".res_conv.wgrad", cr, g, H, *xa, xb.get())) return;
            // input gradients: conv1 path + residual path (identity or 1x1 conv) [+ what the tensor already holds]
            int ci0 = 0;
            Ten* src[2] = {xa.get(), xb.get()};
            for (int i = 0; i < 2; ++i) {
                if (!src[i]) continue;
                Ten& x = *src[i];
                const bf16* res;
                if (cr) {
                    if (!conv_dgrad(p + ".res_conv.dgrad", cr, g, H, ci0, x.C, T3, nullptr)) return;
                    res = T3;
                } else {
                    res = g;     // identity skip: a single source with C == Cout
                }
                if (x.gset) {    // fold the existing gradient into the residual term first (the conv output must not alias it)
                    bf16* t3 = T3;
                    const bf16* xg = x.g;
                    const size_t n = elems(x.H, x.C);
                    push("pointwise", p + ".acc", [=](cudaStream_t s) { return add_bf16_run(res, xg, t3, static_cast<long long>(n), s); });
                    res = T3;
                }
                if (!conv_dgrad(p + ".block1.dgrad", c1, T2, H, ci0, x.C, x.g, res)) return;
                x.gset = true;
                ci0 += x.C;
            }
        });
        return out;
    }

    // ---------------------------------------------------------------- Residual(PreNorm(LinearAttention | Attention))
    TenP attention(const std::string& p, const TenP& x, bool linear) {
        const int H = x->H, C = x->C, n = H * H;
        const long long M = rows(H);
        const TParam* g1 = find_p(t, p + ".fn.norm.g", {1, C, 1, 1});
        ConvW* cq = convw(p + ".fn.fn.to_qkv.weight", "", 384, C, 1, false);
        ConvW* co = convw(p + (linear ? ".fn.fn.to_out.0.weight" : ".fn.fn.to_out.weight"), p + (linear ? ".fn.fn.to_out.0.bias" : ".fn.fn.to_out.bias"),
                          C, 128, 1, false);
        const TParam* g2 = linear ? find_p(t, p + ".fn.fn.to_out.1.g", {1, C, 1, 1}) : nullptr;
        if (!ok || !g1 || (linear && !g2)) { ok = false; return x; }
        TenP xn = ten(H, C), qkv = ten(H, 384), att = ten(H, 128), o = linear ? ten(H, C) : nullptr, out = ten(H, C);
        if (!ok) return out;
        {
            LayerNormArgs a;
            a.x = x->p; a.y = xn->p; a.M = static_cast<int>(M); a.C = C; a.g = g1->w; a.eps = EPS;
            push("layernorm", p + ".norm", [a](cudaStream_t s) { return channel_layernorm_run(a, s); });
        }
        if (!conv_fwd(p + ".to_qkv", cq, *xn, nullptr, qkv->p, ConvEpilogue(), false)) return out;
        if (linear) {
            LinAttnArgs a;
            a.qkv = qkv->p; a.out = att->p; a.ctx = la_ctx; a.B = B; a.n = n;
            push("attention", p + ".linattn", [a](cudaStream_t s) { return linear_attention_run(a, s); });
            if (!conv_fwd(p + ".to_out", co, *att, nullptr, o->p, ConvEpilogue(), false)) return out;
            LayerNormArgs l2;
            l2.x = o->p; l2.y = out->p; l2.M = static_cast<int>(M); l2.C = C; l2.g = g2->w; l2.eps = EPS; l2.res = x->p;
            push("layernorm", p + ".to_out.norm", [l2](cudaStream_t s) { return channel_layernorm_run(l2, s); });
        } else {
            FullAttnArgs a;
            a.qkv = qkv->p; a.out = att->p; a.B = B; a.n = n;
            push("attention", p + ".attn", [a](cudaStream_t s) { return full_attention_run(a, s); });
            ConvEpilogue e;
            e.res = x->p; e.ldr = C;
            if (!conv_fwd(p + ".to_out", co, *att, nullptr, out->p, e, false)) return out;
        }
        bwd_name.push_back(p);
        bwd.push_back([=]() {
            const bf16* g = out->g;
            const bf16* d_o = g;                    // gradient w.r.t. the to_out conv's output
            if (linear) {
                const bf16* op = o->p;
                const float* gain = g2->w;
                float* dg = g2->g;
                bf16* t1 = T1;
                float* part = ln_part;
                push("layernorm_bwd", p + ".to_out.norm.bwd", [=](cudaStream_t s) { return channel_layernorm_bwd_run(op, g, gain, M, C, EPS, t1, dg, part, s); });
                d_o = T1;
            }
            if (!conv_wgrad(p + ".to_out.wgrad", co, d_o, H, *att, nullptr)) return;
            if (!conv_dgrad(p + ".to_out.dgrad", co, d_o, H, 0, 128, att->g, nullptr)) return;
            {
                const bf16 *q = qkv->p, *da = att->g;
                bf16* dq = qkv->g;
                float* sc = la_scratch;
                const int Bn = B;
                if (linear) push("attention_bwd", p + ".linattn.bwd", [=](cudaStream_t s) { return linear_attention_bwd_run(q, da, dq, Bn, n, sc, s); });
                else push("attention_bwd", p + ".attn.bwd", [=](cudaStream_t s) { return full_attention_bwd_run(q, da, dq, Bn, n, s); });
            }
            if (!conv_wgrad(p + ".to_qkv.wgrad", cq, qkv->g, H, *xn, nullptr)) return;
            if (!conv_dgrad(p + ".to_qkv.dgrad", cq, qkv->g, H, 0, C, xn->g, nullptr)) return;
            {
                const bf16 *xp = x->p, *dz = xn->g;
                const float* gain = g1->w;
                float* dg = g1->g;
                bf16* t2 = T2;
                float* part = ln_part;
                push("layernorm_bwd", p + ".norm.bwd", [=](cudaStream_t s) { return channel_layernorm_bwd_run(xp, dz, gain, M, C, EPS, t2, dg, part, s); });
            }
            accumulate(p + ".in", *x, T2);      // PreNorm branch
            accumulate(p + ".skip", *x, g);     // Residual
        });
        return out;
    }

    // ---------------------------------------------------------------- plain conv layers (downs.3.3, ups.3.3) and resampling
    TenP plain_conv(const std::string& p, const std::string& wkey, const std::string& bkey, const TenP& x, int Cout) {
        ConvW* c = convw(wkey, bkey, Cout, x->C, 3, false);
        TenP out = ten(x->H, Cout);
        if (!ok) return out;
        if (!conv_fwd(p, c, *x, nullptr, out->p, ConvEpilogue(), false)) return out;
        bwd_name.push_back(p);
        bwd.push_back([=]() {
            if (!conv_wgrad(p + ".wgrad", c, out->g, x->H, *x, nullptr)) return;
            bf16* where = target(*x, T1);
            if (!conv_dgrad(p + ".dgrad", c, out->g, x->H, 0, x->C, where, nullptr)) return;
            landed(p, *x, where);
        });
        return out;
    }
    TenP downsample(const std::string& p, const TenP& x, int Cout) {      // Rearrange + Conv2d(4C, Cout, 1)
        const int Hl = x->H / 2, C = x->C;
        ConvW* c = convw(p + ".1.weight", p + ".1.bias", Cout, 4 * C, 1, false);
        TenP u = ten(Hl, 4 * C), out = ten(Hl, Cout);
        if (!ok) return out;
        const int Bn = B;
        {
            const bf16* in = x->p;
            bf16* up = u->p;
            push("resample", p + ".unshuffle", [=](cudaStream_t s) { return unshuffle_run(in, up, Bn, Hl, Hl, C, 0, s); });
        }
        if (!conv_fwd(p + ".conv", c, *u, nullptr, out->p, ConvEpilogue(), false)) return out;
        bwd_name.push_back(p);
        bwd.push_back([=]() {
            if (!conv_wgrad(p + ".wgrad", c, out->g, Hl, *u, nullptr)) return;
            if (!conv_dgrad(p + ".dgrad", c, out->g, Hl, 0, 4 * C, u->g, nullptr)) return;
            const bf16* ug = u->g;
            bf16* t1 = T1;
            push("resample", p + ".shuffle", [=](cudaStream_t s) { return unshuffle_run(ug, t1, Bn, Hl, Hl, C, 1, s); });
            accumulate(p, *x, T1);
        });
        return out;
    }
    TenP upsample(const std::string& p, const TenP& x, int Cout) {        // nn.Upsample(2, 'nearest') + Conv2d(C, Cout, 3)
        const int Hl = x->H, C = x->C;
        ConvW* c = convw(p + ".1.weight", p + ".1.bias", Cout, C, 3, false);
        TenP u = ten(2 * Hl, C), out = ten(2 * Hl, Cout);
        if (!ok) return out;
        const int Bn = B;
        {
            const bf16* in = x->p;
            bf16* up = u->p;
            push("resample", p + ".nearest", [=](cudaStream_t s) { return upsample2x_run(in, up, Bn, Hl, Hl, C, s); });
        }
        if (!conv_fwd(p + ".conv", c, *u, nullptr, out->p, ConvEpilogue(), false)) return out;
        bwd_name.push_back(p);
        bwd.push_back([=]() {
            if (!conv_wgrad(p + ".wgrad", c, out->g, 2 * Hl, *u, nullptr)) return;
            if (!conv_dgrad(p + ".dgrad", c, out->g, 2 * Hl, 0, C, u->g, nullptr)) return;
            const bf16* ug = u->g;
            bf16* t1 = T1;
            push("resample", p + ".sumpool", [=](cudaStream_t s) { return sumpool2x_run(ug, t1, Bn, Hl, Hl, C, s); });
            accumulate(p, *x, T1);
        });
        return out;
    }
};

}  // namespace

int build_unet_trainer(hd_trainer* t) {
    const hd_config& c = t->cfg;
    const int B = t->B, dim = c.dim, L = c.num_mults;
    const int cin = c.self_condition ? 2 : 1;
    const int time_dim = dim * 4;
    const long long M0 = static_cast<long long>(B) * S * S;
    if (dim != 64) return tfail("the Unet training step is built for dim = 64 (got %d)", dim);
    std::vector<int> dims{dim};
    for (int i = 0; i < L; ++i) dims.push_back(dim * c.dim_mults[i]);

    const bool sr3 = c.variant == HD_UNET_SR3;
    UB u{t, B};
    u.sr3 = sr3;
    // ---------------------------------------------------------------- FiLM slots (module order, hd_plan_finalize's layout)
    std::vector<std::pair<std::string, int>> blocks;
    for (int i = 0; i < L; ++i) { blocks.push_back({"downs." + std::to_string(i) + ".0", dims[i]}); blocks.push_back({"downs." + std::to_string(i) + ".1", dims[i]}); }
    blocks.push_back({"mid_block1", dims[L]});
    blocks.push_back({"mid_block2", dims[L]});
    for (int k = 0; k < L; ++k) { blocks.push_back({"ups." + std::to_string(k) + ".0", dims[L - k]}); blocks.push_back({"ups." + std::to_string(k) + ".1", dims[L - k]}); }
    blocks.push_back({"final_res_block", dim});
    std::map<std::string, int> film_off;
    int ld = 0;
    for (auto& b : blocks) { film_off[b.first] = ld; ld += (sr3 ? 1 : 2) * b.second; }
    u.ld = ld;

    // ---------------------------------------------------------------- io + shared scratch
    const size_t big = static_cast<size_t>(B) * S * S * 128;      // largest gradient temporary: [B, 64, 64, 128] (ups.3 concat is split per source)
    if (dalloc(t, &t->x, M0 * 4, true) || dalloc(t, &t->cond, M0 * 4, true) || dalloc(t, &t->target, M0 * 4, true) ||
        dalloc(t, &t->eps, M0 * 4) || dalloc(t, &t->d_eps, M0 * 4) || dalloc(t, &t->time, B * 4, true) ||
        dalloc(t, &t->weight, B * 4, true) || dalloc(t, &t->loss, 16, true) ||
        dalloc(t, &u.film, static_cast<size_t>(B) * ld * 4) || dalloc(t, &u.dfilm, static_cast<size_t>(B) * ld * 4, true) ||
        dalloc(t, &u.iota, B * 4) || dalloc(t, &u.gn_part, static_cast<size_t>(B) * (S * S / 32) * 64 * sizeof(float2)) ||
        dalloc(t, &u.gn_scratch, gn_bwd_scratch_floats(B, S * S, 512) * 4) || dalloc(t, &u.la_scratch, linattn_bwd_scratch_floats(B) * 4) ||
        dalloc(t, &u.ln_part, static_cast<size_t>(ln_bwd_blocks(M0)) * 512 * 4) || dalloc(t, &u.cs_part, static_cast<size_t>(colsum_parts(M0)) * 512 * 4) ||
        dalloc(t, &u.la_ctx, static_cast<size_t>(B) * 4 * 32 * 32 * 4) || dalloc(t, &u.zero_bias, 512 * 4, true) ||
        dalloc(t, &u.T1, big * 4 * 2) || dalloc(t, &u.T2, big * 4 * 2) || dalloc(t, &u.T3, big * 4 * 2))
        return 1;
    {
        std::vector<int> h(B);
        for (int i = 0; i < B; ++i) h[i] = i;
        if (cudaMemcpy(u.iota, h.data(), B * 4, cudaMemcpyHostToDevice) != cudaSuccess) return tfail("iota upload failed");
    }

    // ---------------------------------------------------------------- time embedding -> FiLM rows
    const TParam *w1 = find_p(t, "time_mlp.1.weight", {time_dim, dim}), *b1 = find_p(t, "time_mlp.1.bias", {time_dim});
    const TParam *w3 = find_p(t, "time_mlp.3.weight", {time_dim, time_dim}), *b3 = find_p(t, "time_mlp.3.bias", {time_dim});
    if (!w1 || !b1 || !w3 || !b3) return 1;
    float *posenc = nullptr, *z1 = nullptr, *g1 = nullptr, *temb = nullptr, *stemb = nullptr, *d_act = nullptr, *d_g1 = nullptr;
    if (dalloc(t, &posenc, static_cast<size_t>(B) * dim * 4) || dalloc(t, &z1, static_cast<size_t>(B) * time_dim * 4) ||
        dalloc(t, &g1, static_cast<size_t>(B) * time_dim * 4) || dalloc(t, &temb, static_cast<size_t>(B) * time_dim * 4) ||
        dalloc(t, &stemb, static_cast<size_t>(B) * time_dim * 4) || dalloc(t, &d_act, static_cast<size_t>(B) * time_dim * 4) ||
        dalloc(t, &d_g1, static_cast<size_t>(B) * time_dim * 4))
        return 1;
    std::vector<const TParam*> mw, mb;
    for (auto& b : blocks) {
        const int width = (sr3 ? 1 : 2) * b.second;
        mw.push_back(find_p(t, b.first + (sr3 ? ".noise_func.noise_func.0.weight" : ".mlp.1.weight"), {width, time_dim}));
        mb.push_back(find_p(t, b.first + (sr3 ? ".noise_func.noise_func.0.bias" : ".mlp.1.bias"), {width}));
        if (!mw.back() || !mb.back()) return 1;
    }
    // weight preparation ops are pushed by convw() as the layers are declared -> declare the whole net FIRST into a side list,
    // then splice: [prep ...][time ...][forward ...][loss][backward ...]
    std::vector<TOp> time_ops;
    {
        std::vector<TOp> keep;
        keep.swap(t->ops);
        float* tv = t->time;
        const float *w1d = w1->w, *b1d = b1->w, *w3d = w3->w, *b3d = b3->w;
        float* film = u.film;
        u.push("time", "posenc", [=](cudaStream_t s) { return posenc_rows_run(tv, posenc, B, dim, sr3 ? 1 : 0, s); });
        u.push("time", "time_mlp.1", [=](cudaStream_t s) { return linear_rows_run(posenc, dim, w1d, b1d, z1, time_dim, 0, B, dim, time_dim, 0, 0, s); });
        u.push("time", "gelu+time_mlp.3+silu", [=](cudaStream_t s) {
            cudaError_t e = act_apply_run(z1, g1, static_cast<long long>(B) * time_dim, 2, s);
            if (e == cudaSuccess) e = linear_rows_run(g1, time_dim, w3d, b3d, temb, time_dim, 0, B, time_dim, time_dim, 0, 0, s);
            return e != cudaSuccess ? e : act_apply_run(temb, stemb, static_cast<long long>(B) * time_dim, 1, s);
        });
        for (size_t i = 0; i < blocks.size(); ++i) {
            const float *w = mw[i]->w, *bias = mb[i]->w;
            const int off = film_off[blocks[i].first], width = (sr3 ? 1 : 2) * blocks[i].second;
            const float* tin = sr3 ? temb : stemb;
            u.push("time", blocks[i].first + ".mlp", [=](cudaStream_t s) { return linear_rows_run(tin, time_dim, w, bias, film, ld, off, B, time_dim, width, 0, 0, s); });
        }
        time_ops.swap(t->ops);
        t->ops.swap(keep);
    }

    // ---------------------------------------------------------------- forward ([prep.all][time ...][forward ...][loss][backward ...])
    const TParam *iw = find_p(t, "init_conv.weight", {dim, cin, 7, 7}), *ib = find_p(t, "init_conv.bias", {dim});
    const TParam *fw = find_p(t, "final_conv.weight", {1, dim, 1, 1}), *fb = find_p(t, "final_conv.bias", {1});
    if (!iw || !ib || !fw || !fb) return 1;
    u.push_prep();
    t->ops.insert(t->ops.end(), time_ops.begin(), time_ops.end());
    TenP x = u.ten(S, dim);
    if (!u.ok) return 1;
    {
        StemConvArgs a;
        a.x0 = c.self_condition ? t->cond : t->x;
        a.x1 = c.self_condition ? t->x : nullptr;
        a.w = iw->w; a.bias = ib->w; a.y = x->p;
        a.B = B; a.H = S; a.W = S; a.Cout = dim; a.Cin = cin; a.ksize = 7;
        u.push("stem_conv", "init_conv", [a](cudaStream_t s) { return stem_conv_run(a, s); });
    }
    TenP r = x;
    std::vector<TenP> hs;
    for (int i = 0; i < L; ++i) {
        const std::string p = "downs." + std::to_string(i);
        x = u.resblock(p + ".0", x, nullptr, dims[i], film_off[p + ".0"]);
        hs.push_back(x);
        x = u.resblock(p + ".1", x, nullptr, dims[i], film_off[p + ".1"]);
        x = u.attention(p + ".2", x, true);
        hs.push_back(x);
        x = i < L - 1 ? u.downsample(p + ".3", x, dims[i + 1]) : u.plain_conv(p + ".3", p + ".3.weight", p + ".3.bias", x, dims[i + 1]);
        if (!u.ok) return 1;
    }
    x = u.resblock("mid_block1", x, nullptr, dims[L], film_off["mid_block1"]);
    x = u.attention("mid_attn", x, false);
    x = u.resblock("mid_block2", x, nullptr, dims[L], film_off["mid_block2"]);
    if (!u.ok) return 1;
    for (int k = 0; k < L; ++k) {
        const std::string p = "ups." + std::to_string(k);
        const int Cout = dims[L - k], Cprev = dims[L - k - 1];
        TenP s1 = hs.back(); hs.pop_back();
        x = u.resblock(p + ".0", x, s1, Cout, film_off[p + ".0"]);
        TenP s2 = hs.back(); hs.pop_back();
        x = u.resblock(p + ".1", x, s2, Cout, film_off[p + ".1"]);
        x = u.attention(p + ".2", x, true);
        x = k < L - 1 ? u.upsample(p + ".3", x, Cprev) : u.plain_conv(p + ".3", p + ".3.weight", p + ".3.bias", x, Cprev);
        if (!u.ok) return 1;
    }
    x = u.resblock("final_res_block", x, r, dim, film_off["final_res_block"]);
    if (!u.ok) return 1;
    {
        HeadConvArgs a;
        a.x = x->p; a.w = fw->w; a.bias = fb->w; a.eps = t->eps; a.M = static_cast<int>(M0); a.C = dim;
        u.push("head_conv1x1", "final_conv", [a](cudaStream_t s) { return head_conv1x1_run(a, s); });
    }
    // ---------------------------------------------------------------- loss
    float* loss_part = nullptr;
    if (dalloc(t, &loss_part, static_cast<size_t>(loss_parts()) * 4)) return 1;
    {
        hd_trainer* tt = t;
        u.push("loss", "loss", [=](cudaStream_t s) {
            return loss_grad_run(tt->eps, tt->target, tt->weight, tt->loss_kind, B, S * S, tt->d_eps, loss_part, tt->loss, s);
        });
    }
    // ---------------------------------------------------------------- backward: final_conv, then the layers in reverse
    {
        float* hpart = nullptr;
        if (dalloc(t, &hpart, static_cast<size_t>(head_bwd_parts(M0)) * dim * 4)) return 1;
        const bf16* xp = x->p;
        bf16* xg = x->g;
        const float *de = t->d_eps, *w = fw->w;
        float *gw = fw->g, *gb = fb->g;
        u.push("head_bwd", "final_conv.bwd", [=](cudaStream_t s) {
            cudaError_t e = head_bwd_run(xp, de, w, M0, dim, xg, hpart, gw, s);
            return e != cudaSuccess ? e : sum_f32_run(de, M0, loss_part, gb, s);
        });
        x->gset = true;
    }
    {
        size_t next_bucket = 0;
        for (size_t i = u.bwd.size(); i-- > 0;) {
            u.pending_join = true;
            // gradient buckets: everything emitted so far (the modules AFTER this one in forward order) is final once the side
            // stream has joined -- a marker op (join, then an event record) in front of the first module of the next bucket
            while (next_bucket < t->bucket_prefix.size() && u.bwd_name[i].compare(0, t->bucket_prefix[next_bucket].size(), t->bucket_prefix[next_bucket]) == 0) {
                const int b = static_cast<int>(next_bucket++);
                u.push("marker", "grad_bucket." + std::to_string(b), [](cudaStream_t) { return cudaSuccess; });
                t->ops.back().bucket = b;
                t->ops.back().join = 1;
            }
            u.bwd[i]();
            if (!u.ok) return 1;
        }
        if (next_bucket != t->bucket_prefix.size())
            return tfail("gradient bucket boundary '%s' matches no module of this net (in backward order)", t->bucket_prefix[next_bucket].c_str());
    }
    u.pending_join = true;      // init_conv's column sums share cs_part with the side stream
    {   // init_conv: weight / bias gradient (its input needs none)
        float* spart = nullptr;
        if (dalloc(t, &spart, static_cast<size_t>(B) * 8 * 2 * 49 * dim * 4)) return 1;
        const bf16* g = r->g;
        const float* u0 = c.self_condition ? t->cond : t->x;
        const float* u1 = c.self_condition ? t->x : nullptr;
        float *gw = iw->g, *gb = ib->g;
        float* cs = u.cs_part;
        u.push("stem_wgrad", "init_conv.wgrad", [=](cudaStream_t s) {
            cudaError_t e = stem_wgrad_run(g, u0, u1, B, dim, 7, spart, gw, s);
            return e != cudaSuccess ? e : colsum_run(g, M0, dim, cs, 1.0f, 0, gb, s);
        });
    }
    // ---------------------------------------------------------------- backward: time-embedding MLPs
    {
        float* dfilm = u.dfilm;
        {
            const int ns = static_cast<int>(blocks.size());
            std::vector<LinSlot> hs_slots(ns);
            int max_width = 0;
            for (int i = 0; i < ns; ++i) {
                const int width = (sr3 ? 1 : 2) * blocks[i].second;
                hs_slots[i] = LinSlot{mw[i]->w, mw[i]->g, mb[i]->g, film_off[blocks[i].first], width};
                max_width = width > max_width ? width : max_width;
            }
            LinSlot* slots = nullptr;
            float* lpart = nullptr;
            if (dalloc(t, &slots, sizeof(LinSlot) * ns) || dalloc(t, &lpart, static_cast<size_t>(ns) * B * time_dim * 4)) return 1;
            if (cudaMemcpy(slots, hs_slots.data(), sizeof(LinSlot) * ns, cudaMemcpyHostToDevice) != cudaSuccess) return tfail("slot upload failed");
            const float* tin = sr3 ? temb : stemb;
            u.push("time_bwd", "*.mlp.bwd", [=](cudaStream_t s) {
                cudaError_t e = linear_bwd_weight_batched_run(dfilm, ld, tin, time_dim, B, time_dim, slots, ns, max_width, s);
                return e != cudaSuccess ? e : linear_bwd_input_batched_run(dfilm, ld, B, time_dim, slots, ns, lpart, d_act, s);
            });
        }
        float *gw3 = w3->g, *gb3 = b3->g, *gw1 = w1->g, *gb1 = b1->g;
        const float* w3d = w3->w;
        u.push("time_bwd", "time_mlp.bwd", [=](cudaStream_t s) {
            cudaError_t e = sr3 ? cudaSuccess : act_grad_run(d_act, temb, static_cast<long long>(B) * time_dim, 1, s);
            if (e == cudaSuccess) e = linear_bwd_weight_run(d_act, time_dim, 0, g1, time_dim, B, time_dim, time_dim, 0, gw3, gb3, s);
            if (e == cudaSuccess) e = linear_bwd_input_run(d_act, time_dim, 0, w3d, B, time_dim, time_dim, 0, d_g1, time_dim, s);
            if (e == cudaSuccess) e = act_grad_run(d_g1, z1, static_cast<long long>(B) * time_dim, 2, s);
            if (e == cudaSuccess) e = linear_bwd_weight_run(d_g1, time_dim, 0, posenc, dim, B, dim, time_dim, 0, gw1, gb1, s);
            return e;
        });
    }
    if (u.ok && !u.finish_prep()) return 1;
    return u.ok ? 0 : 1;
}

}  // namespace hd
