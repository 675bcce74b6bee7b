// Plan = one eps-predictor + its schedule, resident on one GPU; C ABI of include/hicdiff_b200.h.
//
// The plan turns the reference's module tree (Unet: /root/reference/src/hicdiff_condition.py:255-384,
// hicdiff_sr3.py:310-445; hicedrn_Diff: /root/reference/src/model/hicedrn_Diff.py:210-289) into a flat list of
// kernel launches over a static activation arena, captures one denoising step (eps-net + posterior update +
// step-counter decrement) into a CUDA graph per batch size and replays it T times
// (p_sample_loop, hicdiff_condition.py:600-623).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/hicdiff_b200.h"
#include "kernels.h"

using namespace hd;

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

constexpr int LA_MAX_PARTS = 8;
constexpr float LA_MAX_BOUND = 100.0f;   // beyond this the analytic softmax shift could underflow: use the unfused kernels

struct WeightT {
    float* d = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

struct Act {
    bf16* p = nullptr;
    int H = 0, W = 0, C = 0;
};

// First-fit arena over byte offsets.  Offsets are deterministic for a given op sequence, so a dry run sizes it.
struct Arena {
    struct Blk { size_t off, size; bool free; };
    std::vector<Blk> blks;
    size_t high = 0;
    bool never_free = false;
    size_t alloc(size_t bytes) {
        bytes = (bytes + 1023) & ~size_t(1023);
        for (size_t i = 0; i < blks.size(); ++i) {
            if (blks[i].free && blks[i].size >= bytes) {
                if (blks[i].size > bytes) {
                    Blk rest{blks[i].off + bytes, blks[i].size - bytes, true};
                    blks[i].size = bytes;
                    blks.insert(blks.begin() + i + 1, rest);
                }
                blks[i].free = false;
                return blks[i].off;
            }
        }
        if (!blks.empty() && blks.back().free) {
            blks.back().size = bytes;
            blks.back().free = false;
            high = blks.back().off + bytes;
            return blks.back().off;
        }
        Blk b{high, bytes, false};
        blks.push_back(b);
        high += bytes;
        return b.off;
    }
    void release(size_t off) {
        if (never_free) return;
        for (size_t i = 0; i < blks.size(); ++i) {
            if (blks[i].off == off && !blks[i].free) {
                blks[i].free = true;
                if (i + 1 < blks.size() && blks[i + 1].free) {
                    blks[i].size += blks[i + 1].size;
                    blks.erase(blks.begin() + i + 1);
                }
                if (i > 0 && blks[i - 1].free) {
                    blks[i - 1].size += blks[i].size;
                    blks.erase(blks.begin() + i);
                }
                return;
            }
        }
    }
};

struct Op {
    std::function<cudaError_t(cudaStream_t)> fn;
    std::string tag;
    const char* kernel = "";   // kernel family (for the profile breakdown)
    double flops = 0;          // algorithmic FLOPs of one launch (2*M*N*K for the GEMMs)
    double bytes = 0;          // algorithmic HBM bytes of one launch (each operand read / written once)
};

struct FilmSlot {
    std::string wkey, bkey;  // Linear weight / bias keys
    int off = 0;             // column offset in a FiLM row
    int width = 0;           // 2*Cout (scale|shift) or Cout (SR3 additive)
    bool silu_in = false;    // ResnetBlock.mlp = SiLU -> Linear; SR3 noise_func = Linear only
};

struct DebugEntry {
    const bf16* p;
    int H, W, C;
};

struct Exec {
    int B = 0;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    float* x = nullptr;      // [B, 4096] sample state
    float* cond = nullptr;   // [B, 4096]
    float* eps = nullptr;    // [B, 4096]
    float* time = nullptr;   // [B]
    float* posenc = nullptr; // [B, fourier]
    float* temb0 = nullptr;  // [B, time_dim]
    float* temb = nullptr;   // [B, time_dim]
    float* film_rows = nullptr;  // [B, film_ld]
    int* iota = nullptr;     // [B]
    float* ctx = nullptr;    // linear-attention scratch [B,4,32,32]
    float2* gn_part = nullptr;  // GroupNorm partial statistics [B*P/32][C/8], rewritten by every GN-feeding conv
    int* gn_cnt = nullptr;      // per-image arrival counters of the GroupNorm-fused conv epilogue [B]
    float2* gn_ma = nullptr;    // folded GroupNorm affine (mul, add) per (image, channel) of a block2 conv [B][C <= 512] (resblock tail fusion)
    float* la_ctx = nullptr;    // fused linear attention: partial contexts [B*8][128][32]
    float* la_s = nullptr;      //                          partial softmax denominators [B*8][128]
    bf16* la_mb = nullptr;      //                          per-image to_out * ctx matrices [B][128][128]
    std::vector<Op> ops_rows, ops_table;
    cudaGraphExec_t g_eps = nullptr, g_step = nullptr;
    std::map<std::string, DebugEntry> dbg;
    std::vector<std::string> dbg_order;
    size_t extra_bytes = 0;
};

}  // namespace

struct LinAttnPrep {          // gain-folded to_qkv weights of one LinearAttention block (linattn_fused.cu)
    bf16* wqkv = nullptr;     // [384, C]
    float* rowsum = nullptr;  // [384]
    float* kshift = nullptr;  // [128]
    int* bound_bits = nullptr;
    float bound = 0.f;        // max_d sqrt(C) * ||Wk'_d||_2, read back at finalize
    int C = 0;
};

struct hd_plan {
    hd_config cfg;
    int device = 0;
    int num_sms = 148;
    std::map<std::string, WeightT> w;
    std::map<std::string, bf16*> wq;      // GEMM-layout bf16 weights
    std::map<std::string, float*> padded; // zero-padded fp32 vectors (bias of padded-N convs)
    std::map<std::string, LinAttnPrep> lattn;  // Residual prefix -> fused linear-attention weights
    std::vector<FilmSlot> film;
    std::map<std::string, int> film_index;  // block prefix -> slot
    int film_ld = 0;
    int fourier_dim = 0, time_dim = 0;
    bool sr3 = false, hicedrn = false;
    bool wsplit = false;          // "bf16w2" precision: conv weights kept as hi + lo bf16 pairs (hd_config.reserved[0] bits 4-5 == 1)
    int T = 0;
    float* coef = nullptr;        // [T, 8]
    float* time_values = nullptr; // [T]
    float* film_table = nullptr;  // [T, film_ld]
    SampleCtl* ctl = nullptr;     // device control block (sampling)
    SampleCtl* ctl_one = nullptr; // device control block (hd_ddpm_step)
    cudaStream_t cap_stream = nullptr;  // private stream used only for graph capture
    bool finalized = false;
    size_t weight_bytes = 0;
    std::map<int, std::unique_ptr<Exec>> execs;
};

namespace {

const WeightT* find_w(const hd_plan* P, const std::string& key) {
    auto it = P->w.find(key);
    return it == P->w.end() ? nullptr : &it->second;
}

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// ------------------------------------------------------------------------------------------------------------
// Time embedding -> FiLM rows.  Reference: time_mlp (hicdiff_condition.py:300-305), ResnetBlock.mlp (:176-191),
// SR3 FeatureWiseAffine (hicdiff_sr3.py:167-183); HiCEDRN: hicedrn_Diff.py:232-246,189-200.
// ------------------------------------------------------------------------------------------------------------
int enqueue_film(hd_plan* P, const float* time_values, int rows, float* posenc, float* temb0, float* temb,
                 float* film_out, std::vector<Op>* ops) {
    const WeightT *w1 = find_w(P, "time_mlp.1.weight"), *b1 = find_w(P, "time_mlp.1.bias");
    const WeightT *w3 = find_w(P, "time_mlp.3.weight"), *b3 = find_w(P, "time_mlp.3.bias");
    if (!w1 || !b1 || !w3 || !b3) return fail("time_mlp weights missing");
    const int fd = P->fourier_dim, td = P->time_dim, ld = P->film_ld;
    const int mode = P->sr3 ? 1 : 0;
    ops->push_back({[=](cudaStream_t s) { return posenc_rows_run(time_values, posenc, rows, fd, mode, s); }, "posenc"});
    const float *w1d = w1->d, *b1d = b1->d, *w3d = w3->d, *b3d = b3->d;
    ops->push_back({[=](cudaStream_t s) { return linear_rows_run(posenc, fd, w1d, b1d, temb0, td, 0, rows, fd, td, 0, 1, s); },
                    "time_mlp.1+gelu"});
    ops->push_back({[=](cudaStream_t s) { return linear_rows_run(temb0, td, w3d, b3d, temb, td, 0, rows, td, td, 0, 0, s); },
                    "time_mlp.3"});
    for (const FilmSlot& f : P->film) {
        const WeightT *fw = find_w(P, f.wkey), *fb = find_w(P, f.bkey);
        if (!fw || !fb) return fail("weight %s missing", f.wkey.c_str());
        const float *fwd = fw->d, *fbd = fb->d;
        const int off = f.off, width = f.width, act = f.silu_in ? 1 : 0;
        ops->push_back({[=](cudaStream_t s) { return linear_rows_run(temb, td, fwd, fbd, film_out, ld, off, rows, td, width, act, 0, s); },
                        f.wkey});
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Graph builder
// ------------------------------------------------------------------------------------------------------------
struct FilmRef {
    const float* base;
    const int* row;
    int row_stride;
};

struct Builder {
    hd_plan* P;
    Exec* ex;
    int B;
    bool dry;
    Arena arena;
    std::vector<Op>* ops;
    FilmRef film;
    std::string err;
    bool ok = true;
    int last_part_tpi = 0, last_part_W = 0;   // partial-statistics layout of the last conv that produced GroupNorm partials

    Act alloc_act(int H, int W, int C) {
        Act a;
        a.H = H; a.W = W; a.C = C;
        const size_t off = arena.alloc(static_cast<size_t>(B) * H * W * C * 2);
        a.p = reinterpret_cast<bf16*>(static_cast<char*>(dry ? nullptr : ex->arena) + off);
        return a;
    }
    void free_act(const Act& a) {
        if (a.p == nullptr && !dry) return;
        arena.release(reinterpret_cast<const char*>(a.p) - static_cast<const char*>(dry ? nullptr : ex->arena));
    }
    void note(const std::string& name, const Act& a) {
        if (!dry && P->cfg.debug_keep) {
            ex->dbg[name] = DebugEntry{a.p, a.H, a.W, a.C};
            ex->dbg_order.push_back(name);
        }
    }
    bool bad(const std::string& m) {
        if (ok) { ok = false; err = m; }
        return false;
    }
    const float* wf(const std::string& key) {
        const WeightT* w = find_w(P, key);
        if (!w) { bad("weight '" + key + "' missing from the plan"); return nullptr; }
        return w->d;
    }
    const bf16* wq(const std::string& key) {
        auto it = P->wq.find(key);
        if (it == P->wq.end()) { bad("prepared weight '" + key + "' missing"); return nullptr; }
        return it->second;
    }

    // conv on the tcgen05 path
    Act conv(const std::string& wkey, const std::string& bkey, const Act& x0, const Act* x1, int Cout, int ksize,
             ConvMode mode, ConvEpilogue epi, int n_rows = 0, bool gn_stats = false) {
        const int Ho = mode == CONV_UNSHUFFLE ? x0.H / 2 : (mode == CONV_UPSAMPLE ? x0.H * 2 : x0.H);
        const int Wo = mode == CONV_UNSHUFFLE ? x0.W / 2 : (mode == CONV_UPSAMPLE ? x0.W * 2 : x0.W);
        const int N = n_rows ? n_rows : Cout;
        Act y;
        if (epi.out_f32 == nullptr) y = alloc_act(Ho, Wo, N); else { y.H = Ho; y.W = Wo; y.C = N; }
        if (!bkey.empty()) epi.bias = n_rows ? padded_bias(bkey) : wf(bkey);
        if (gn_stats) epi.gn_part = dry ? nullptr : ex->gn_part;
        ConvGemmDesc d;
        d.src0 = ConvSrc{x0.p, x0.C};
        d.src1 = ConvSrc{x1 ? x1->p : nullptr, x1 ? x1->C : 0};
        d.B = B; d.H = Ho; d.W = Wo; d.ksize = ksize; d.mode = mode;
        d.weight = wq(wkey);
        d.pad_mode = (P->cfg.reserved[0] >> 1) & 3;
        d.cg2_mode = (P->cfg.reserved[0] >> 3) & 1;
        d.dx3_mode = ((P->cfg.reserved[0] >> 6) & 1) ? 0 : 2;   // bit 6: one MMA per tap in the N = 64 3x3 convs
        d.wsplit = P->wsplit ? 1 : 0;
        d.static_weights = 1;       // prepared at hd_plan_finalize, constant for every graph of the plan
        d.N = N;
        d.out = y.p;
        d.epi = epi;
        if (!ok) return y;
        if (!dry) {
            ConvGemmLaunch l;
            char e[256];
            if (conv_gemm_prepare(d, P->num_sms, &l, e, sizeof(e))) { bad(std::string(wkey) + ": " + e); return y; }
            if (gn_stats) { last_part_tpi = l.kind == 3 ? l.tiles_per_img : 0; last_part_W = l.kind == 3 ? Wo : 0; }
            Op op{[l](cudaStream_t s) { return conv_gemm_run(l, s); }, wkey};
            if (epi.gn_gamma != nullptr) {   // the fused GroupNorm epilogue counts warp-block arrivals per image from zero
                int* cnt = epi.gn_counter;
                const size_t cbytes = static_cast<size_t>(B) * 4;
                op.fn = [l, cnt, cbytes](cudaStream_t s) {
                    cudaError_t e = cudaMemsetAsync(cnt, 0, cbytes, s);
                    return e != cudaSuccess ? e : conv_gemm_run(l, s);
                };
            }
            op.kernel = "conv_gemm";
            const double M = static_cast<double>(B) * Ho * Wo;
            const double K = static_cast<double>(l.nkb) * 64;
            // FLOPs the kernel EXECUTES (the upsample fold runs 4 of the reference's 9 taps per output pixel)
            op.flops = 2.0 * M * Cout * K;
            const double in_scale = mode == CONV_UNSHUFFLE ? 4.0 : (mode == CONV_UPSAMPLE ? 0.25 : 1.0);
            op.bytes = 2.0 * (M * (x0.C + (x1 ? x1->C : 0)) * in_scale + static_cast<double>(N) * K * (mode == CONV_UPSAMPLE ? 4 : 1)) +
                       (epi.out_f32 ? 4.0 * M * epi.n_valid : 2.0 * M * N) + (epi.res ? 2.0 * M * N : 0.0);
            ops->push_back(op);
        }
        return y;
    }
    const float* padded_bias(const std::string& key) {
        auto it = P->padded.find(key);
        if (it == P->padded.end()) { bad("padded bias '" + key + "' missing"); return nullptr; }
        return it->second;
    }

    void groupnorm(const Act& x, const std::string& prefix, int film_off, int postadd_off, const Act* res, const Act* y = nullptr) {
        GroupNormArgs g;
        g.x = x.p; g.y = y ? y->p : x.p;   // default in place: every thread rewrites exactly the 16-byte chunks it read
        g.B = B; g.P = x.H * x.W; g.C = x.C;
        g.part = dry ? nullptr : ex->gn_part;
        g.part_tpi = last_part_tpi; g.part_W = last_part_W;
        g.gamma = wf(prefix + ".weight");
        g.beta = wf(prefix + ".bias");
        g.eps = 1e-5f;
        g.film_row = film.row; g.film_row_stride = film.row_stride; g.film_ld = P->film_ld;
        if (film_off >= 0) { g.film = film.base; g.film_off = film_off; }
        if (postadd_off >= 0) { g.postadd = film.base; g.postadd_off = postadd_off; }
        if (res) g.res = res->p;
        g.exact_act = P->wsplit ? 1 : 0;
        if (!ok || dry) return;
        Op op{[g](cudaStream_t s) { return groupnorm_film_silu_run(g, s); }, prefix};
        op.kernel = "groupnorm_film_silu";
        op.bytes = 2.0 * B * x.H * x.W * x.C * (res ? 3 : 2);
        ops->push_back(op);
    }

    Act layernorm(const Act& x, const std::string& gkey, const Act* res, bool up2x, const Act* x_lo = nullptr) {
        Act y = up2x ? alloc_act(x.H * 2, x.W * 2, x.C) : alloc_act(x.H, x.W, x.C);
        LayerNormArgs l;
        l.x = x.p; l.y = y.p; l.M = B * x.H * x.W; l.C = x.C; l.g = wf(gkey); l.eps = 1e-5f;
        l.res = res ? res->p : nullptr; l.upsample2x = up2x ? 1 : 0; l.H = x.H; l.W = x.W;
        l.x_lo = x_lo ? x_lo->p : nullptr;
        if (ok && !dry) {
            Op op{[l](cudaStream_t s) { return channel_layernorm_run(l, s); }, gkey};
            op.kernel = "channel_layernorm";
            op.bytes = 2.0 * B * x.H * x.W * x.C * (1 + (up2x ? 4 : 1) + (res ? 1 : 0));
            ops->push_back(op);
        }
        return y;
    }

    // Block (hicdiff_condition.py:155-171): conv -> GroupNorm -> [FiLM] -> SiLU [-> + SR3 embedding] [-> + residual].  One
    // launch when the conv runs on the slab path (its epilogue then applies the norm itself), else conv + streaming GN.
    Act conv_norm(const std::string& p, const Act& x0, const Act* x1, int Cout, int film_off, int postadd_off, const Act* res) {
        // Measured on B200 (profiles/r01_notes.md): the fused epilogue's per-tile dependency chain (arrival counter, partial
        // merge, second TMEM pass) costs more than the streaming GroupNorm pass it removes (7.3 vs 6.4 ms per step), so the
        // plan keeps conv + groupnorm_apply unless cfg.reserved[0] bit 0 asks for the fused form (kept for the parity tests).
        if ((P->cfg.reserved[0] & 1) && !P->cfg.debug_keep && conv_gemm_can_fuse_gn(B, x0.H, x0.W, Cout, 3, CONV_TAPS)) {
            ConvEpilogue e;
            e.gn_gamma = wf(p + ".norm.weight");
            e.gn_beta = wf(p + ".norm.bias");
            e.gn_eps = 1e-5f;
            e.gn_counter = dry ? nullptr : ex->gn_cnt;
            if (film_off >= 0 || postadd_off >= 0) {
                e.film = film.base; e.film_row = film.row; e.film_row_stride = film.row_stride; e.film_ld = P->film_ld;
                e.film_off = film_off >= 0 ? film_off : postadd_off;
                e.film_has_scale = film_off >= 0 ? 1 : 0;
            }
            if (res) { e.res = res->p; e.ldr = Cout; }
            return conv(p + ".proj.weight", p + ".proj.bias", x0, x1, Cout, 3, CONV_TAPS, e, 0, true);
        }
        Act h = conv(p + ".proj.weight", p + ".proj.bias", x0, x1, Cout, 3, CONV_TAPS, ConvEpilogue(), 0, true);
        note(p + ".proj", h);
        // HD_GN_OUT_OF_PLACE=1 (A/B switch): the 2-stream GroupNorm pass writes a second buffer instead of rewriting its input
        static const int oop_env = [] { const char* v = getenv("HD_GN_OUT_OF_PLACE"); return v ? atoi(v) : 0; }();
        if (oop_env && res == nullptr && !P->cfg.debug_keep) {
            Act y = alloc_act(h.H, h.W, h.C);
            groupnorm(h, p + ".norm", film_off, postadd_off, res, &y);
            free_act(h);
            return y;
        }
        groupnorm(h, p + ".norm", film_off, postadd_off, res);
        return h;
    }

    // ResnetBlock (hicdiff_condition.py:173-197 / hicdiff_sr3.py:235-251)
    Act resblock(const std::string& p, const Act& x0, const Act* x1, int Cout) {
        const int Cin = x0.C + (x1 ? x1->C : 0);
        const int slot = P->film_index.count(p) ? P->film_index[p] : -1;
        if (slot < 0) bad("no time-embedding slot for block " + p);
        const int off = slot >= 0 ? P->film[slot].off : 0;
        Act h = conv_norm(p + ".block1", x0, x1, Cout, P->sr3 ? -1 : off, P->sr3 ? off : -1, nullptr);
        note(p + ".block1", h);
        // Opt-in (HD_RESBLOCK_TAIL=1 / option bit 7) for blocks with a res_conv (Cin != Cout): ONE launch for the tail.  block2's conv
        // leaves its raw output + GroupNorm partials, a one-CTA-per-image kernel folds the statistics into a per-(image, channel)
        // affine, and res_conv's epilogue adds SiLU(GN(raw)) to its own 1x1 result -- instead of res_conv -> r, then a 3-stream
        // groupnorm_apply (raw, r -> out): the r tensor is never written or re-read (2/3 of the two launches' HBM traffic).
        // Measured (profiles/r02_notes.md 11): parity-green, but NOT faster -- at 64x64 the fused launch takes 134.6 us against
        // 67.4 + 67.4 us, because the 8-warp conv epilogue is a latency chain (~5 clk per instruction) and the extra SiLU / affine
        // work lands on it, while groupnorm_apply streams with 16 warps per SM.  So the two-launch form stays the default.
        static const int tail_env = [] { const char* v = getenv("HD_RESBLOCK_TAIL"); return v ? atoi(v) : -1; }();
        const bool tail_opt = tail_env >= 0 ? tail_env != 0 : ((P->cfg.reserved[0] >> 7) & 1) != 0;
        if (Cin != Cout && tail_opt && !P->cfg.debug_keep && !(P->cfg.reserved[0] & 1) && !P->wsplit && ((P->cfg.reserved[0] >> 1) & 3) == 0 &&
            (x0.H * x0.W) % 32 == 0 && Cout % 64 == 0) {
            Act raw = conv(p + ".block2.proj.weight", p + ".block2.proj.bias", h, nullptr, Cout, 3, CONV_TAPS, ConvEpilogue(), 0, true);
            free_act(h);
            if (ok && !dry && last_part_tpi != 0) bad(p + ": the fused ResnetBlock tail needs the dense GroupNorm partial layout (HD_CONV_PAD is set?)");
            if (ok && !dry) {
                const float2* part = ex->gn_part;
                float2* ma = ex->gn_ma;
                const float* gamma = wf(p + ".block2.norm.weight");
                const float* beta = wf(p + ".block2.norm.bias");
                const int Bn = B, Pn = raw.H * raw.W, Cn = raw.C;
                Op op{[=](cudaStream_t s) { return groupnorm_finalize_run(part, gamma, beta, 1e-5f, ma, Bn, Pn, Cn, s); }, p + ".block2.norm"};
                op.kernel = "groupnorm_finalize";
                op.bytes = 8.0 * B * (Pn / 32) * (Cn / 8);
                ops->push_back(op);
            }
            ConvEpilogue e;
            e.res = raw.p; e.ldr = Cout;
            e.gnres = dry ? reinterpret_cast<const float2*>(8) : ex->gn_ma;
            Act out = conv(p + ".res_conv.weight", p + ".res_conv.bias", x0, x1, Cout, 1, CONV_TAPS, e);
            free_act(raw);
            note(p, out);
            return out;
        }
        // res_conv first: block2's conv must be the last writer of the shared partial-statistics buffer before its norm
        Act r;
        bool own_r = false;
        if (Cin != Cout) {
            r = conv(p + ".res_conv.weight", p + ".res_conv.bias", x0, x1, Cout, 1, CONV_TAPS, ConvEpilogue());
            own_r = true;
        } else {
            r = x0;
        }
        Act h2 = conv_norm(p + ".block2", h, nullptr, Cout, -1, -1, &r);
        free_act(h);
        if (own_r) free_act(r);
        note(p, h2);
        return h2;
    }

    // Residual(PreNorm(dim, LinearAttention(dim)))  (hicdiff_condition.py:199-227,319)
    Act linattn(const std::string& p, const Act& x, bool up2x) {
        auto lp = P->lattn.find(p);
        // (the fused block keeps its own bf16 copies of to_qkv / to_out: with split weights the block runs unfused, where both
        // projections are conv_gemm launches over the hi + lo weights)
        if (!P->wsplit && !up2x && lp != P->lattn.end() && lp->second.bound <= LA_MAX_BOUND && lp->second.C == x.C &&
            x.H * x.W >= 128 && (x.H * x.W) % 128 == 0) {
            // whole Residual(PreNorm(LinearAttention)) block in three launches; qkv never reaches HBM
            Act y = alloc_act(x.H, x.W, x.C);
            LinAttnFusedDesc d;
            d.x = x.p; d.y = y.p; d.B = B; d.n = x.H * x.W; d.C = x.C;
            d.wqkv = lp->second.wqkv; d.rowsum = lp->second.rowsum; d.kshift = lp->second.kshift;
            d.wo = wf(p + ".fn.fn.to_out.0.weight"); d.bo = wf(p + ".fn.fn.to_out.0.bias"); d.g2 = wf(p + ".fn.fn.to_out.1.g");
            d.ctx_part = dry ? nullptr : ex->la_ctx; d.s_part = dry ? nullptr : ex->la_s; d.mb = dry ? nullptr : ex->la_mb;
            d.max_parts = LA_MAX_PARTS; d.eps = 1e-5f;
            if (ok && !dry) {
                LinAttnFusedLaunch l;
                char e[256];
                if (linattn_fused_prepare(d, P->num_sms, &l, e, sizeof(e))) { bad(p + ": " + e); return y; }
                Op op{[l](cudaStream_t s) { return linattn_fused_run(l, s); }, p + ".linattn_fused"};
                op.kernel = "linattn_fused";
                const double M = static_cast<double>(B) * d.n;
                op.flops = 2.0 * M * x.C * 384 + 2.0 * M * 128 * 128 + 2.0 * M * 128 * x.C;   // projections, context, mixed output
                op.bytes = 2.0 * M * x.C * 3;                                                 // x twice, y once
                ops->push_back(op);
            }
            note(p, y);
            return y;
        }
        Act xn = layernorm(x, p + ".fn.norm.g", nullptr, false);
        Act qkv = conv(p + ".fn.fn.to_qkv.weight", "", xn, nullptr, 384, 1, CONV_TAPS, ConvEpilogue());
        free_act(xn);
        Act att = alloc_act(x.H, x.W, 128);
        LinAttnArgs la;
        la.qkv = qkv.p; la.out = att.p; la.ctx = dry ? nullptr : ex->ctx; la.B = B; la.n = x.H * x.W;
        if (ok && !dry) {
            Op op{[la](cudaStream_t s) { return linear_attention_run(la, s); }, p + ".linattn"};
            op.kernel = "linear_attention";
            const double n = static_cast<double>(x.H) * x.W;
            op.flops = 2.0 * 2.0 * B * 4 * 32 * 32 * n;
            op.bytes = 2.0 * B * n * (384 + 256 + 128);   // k,v twice (max pass + ctx pass), q once, out once
            ops->push_back(op);
        }
        free_act(qkv);
        note(p + ".attn_core", att);
        // "bf16w2" precision: the to_out conv keeps its output as hi + lo bf16 (the LayerNorm behind it cancels a large
        // per-pixel common component, which amplifies that one tensor's rounding; kernels.h, ConvEpilogue::out_lo)
        ConvEpilogue eo;
        Act o_lo;
        if (P->wsplit) { o_lo = alloc_act(x.H, x.W, x.C); eo.out_lo = o_lo.p; }
        Act o = conv(p + ".fn.fn.to_out.0.weight", p + ".fn.fn.to_out.0.bias", att, nullptr, x.C, 1, CONV_TAPS, eo);
        free_act(att);
        Act y = layernorm(o, p + ".fn.fn.to_out.1.g", &x, up2x, P->wsplit ? &o_lo : nullptr);
        free_act(o);
        if (P->wsplit) free_act(o_lo);
        note(p, y);
        return y;
    }

    // Residual(PreNorm(dim, Attention(dim)))  (hicdiff_condition.py:229-251,326)
    Act fullattn(const std::string& p, const Act& x) {
        Act xn = layernorm(x, p + ".fn.norm.g", nullptr, false);
        Act qkv = conv(p + ".fn.fn.to_qkv.weight", "", xn, nullptr, 384, 1, CONV_TAPS, ConvEpilogue());
        free_act(xn);
        Act att = alloc_act(x.H, x.W, 128);
        FullAttnArgs fa;
        fa.qkv = qkv.p; fa.out = att.p; fa.B = B; fa.n = x.H * x.W;
        if (ok && !dry) {
            Op op{[fa](cudaStream_t s) { return full_attention_run(fa, s); }, p + ".attn"};
            op.kernel = "full_attention";
            const double n = static_cast<double>(x.H) * x.W;
            op.flops = 2.0 * 2.0 * B * 4 * 32 * n * n;
            op.bytes = 2.0 * B * n * (384 + 128);
            ops->push_back(op);
        }
        free_act(qkv);
        ConvEpilogue e;
        e.res = x.p; e.ldr = x.C;
        Act y = conv(p + ".fn.fn.to_out.weight", p + ".fn.fn.to_out.bias", att, nullptr, x.C, 1, CONV_TAPS, e);
        free_act(att);
        note(p, y);
        return y;
    }

    Act stem(const std::string& wkey, const std::string& bkey, int Cout, int ksize) {
        const int S = P->cfg.image_size;
        Act y = alloc_act(S, S, Cout);
        StemConvArgs a;
        a.x0 = P->cfg.self_condition ? (dry ? nullptr : ex->cond) : (dry ? nullptr : ex->x);
        a.x1 = P->cfg.self_condition ? (dry ? nullptr : ex->x) : nullptr;
        a.w = wf(wkey); a.bias = wf(bkey); a.y = y.p;
        a.B = B; a.H = S; a.W = S; a.Cout = Cout; a.Cin = P->cfg.self_condition ? 2 : 1; a.ksize = ksize;
        a.precise = P->wsplit ? 1 : 0;      // fp32 weights (and inputs) on the FMA pipe instead of bf16 fragments
        if (ok && !dry) {
            Op op{[a](cudaStream_t s) { return stem_conv_run(a, s); }, wkey};
            op.kernel = "stem_conv";
            op.flops = 2.0 * B * S * S * Cout * a.Cin * ksize * ksize;
            op.bytes = 4.0 * B * S * S * a.Cin + 2.0 * B * S * S * Cout;
            ops->push_back(op);
        }
        return y;
    }

    // Unet.forward (hicdiff_condition.py:345-384)
    void build_unet() {
        const hd_config& c = P->cfg;
        const int dim = c.dim;
        std::vector<int> dims;
        dims.push_back(dim);
        for (int i = 0; i < c.num_mults; ++i) dims.push_back(dim * c.dim_mults[i]);
        const int L = c.num_mults;
        Act x = stem("init_conv.weight", "init_conv.bias", dim, 7);
        note("init_conv", x);
        Act r = x;
        std::vector<Act> hs;
        bool x_is_r = true;
        for (int i = 0; i < L; ++i) {
            const int din = dims[i], dout = dims[i + 1];
            const bool last = i == L - 1;
            const std::string p = "downs." + std::to_string(i);
            Act a = resblock(p + ".0", x, nullptr, din);
            if (!x_is_r) free_act(x);
            x_is_r = false;
            hs.push_back(a);
            Act b = resblock(p + ".1", a, nullptr, din);
            Act cst = linattn(p + ".2", b, false);
            free_act(b);
            hs.push_back(cst);
            if (!last) {
                x = conv(p + ".3.1.weight", p + ".3.1.bias", cst, nullptr, dout, 1, CONV_UNSHUFFLE, ConvEpilogue());
            } else {
                x = conv(p + ".3.weight", p + ".3.bias", cst, nullptr, dout, 3, CONV_TAPS, ConvEpilogue());
            }
            note(p + ".3", x);
        }
        {
            Act m1 = resblock("mid_block1", x, nullptr, dims[L]);
            free_act(x);
            Act m2 = fullattn("mid_attn", m1);
            free_act(m1);
            x = resblock("mid_block2", m2, nullptr, dims[L]);
            free_act(m2);
        }
        for (int k = 0; k < L; ++k) {
            const int i = L - 1 - k;            // reversed(in_out)
            const int din = dims[i], dout = dims[i + 1];
            const bool last = k == L - 1;
            const std::string p = "ups." + std::to_string(k);
            Act s1 = hs.back(); hs.pop_back();
            Act a = resblock(p + ".0", x, &s1, dout);
            free_act(x); free_act(s1);
            Act s2 = hs.back(); hs.pop_back();
            Act b = resblock(p + ".1", a, &s2, dout);
            free_act(a); free_act(s2);
            Act cst = linattn(p + ".2", b, false);
            free_act(b);
            // Upsample: the nearest x2 is folded into the conv (four 2x2 phase convs over the low-res tensor)
            if (!last) x = conv(p + ".3.1.weight", p + ".3.1.bias", cst, nullptr, din, 3, CONV_UPSAMPLE, ConvEpilogue());
            else x = conv(p + ".3.weight", p + ".3.bias", cst, nullptr, din, 3, CONV_TAPS, ConvEpilogue());
            free_act(cst);
            note(p + ".3", x);
        }
        Act f = resblock("final_res_block", x, &r, dim);
        free_act(x); free_act(r);
        HeadConvArgs hca;
        hca.x = f.p; hca.w = wf("final_conv.weight"); hca.bias = wf("final_conv.bias");
        hca.eps = dry ? nullptr : ex->eps; hca.M = B * c.image_size * c.image_size; hca.C = dim;
        if (ok && !dry) {
            Op op{[hca](cudaStream_t s) { return head_conv1x1_run(hca, s); }, "final_conv"};
            op.kernel = "head_conv1x1";
            op.flops = 2.0 * hca.M * dim;
            op.bytes = 2.0 * hca.M * dim + 4.0 * hca.M;
            ops->push_back(op);
        }
        free_act(f);
    }

    // hicedrn_Diff.forward (hicedrn_Diff.py:267-289; ResnetBlock :182-208)
    void build_hicedrn() {
        const hd_config& c = P->cfg;
        const int F = 256;
        Act x = stem("head.weight", "head.bias", F, 3);
        note("head", x);
        Act r = x;
        bool x_is_r = true;
        for (int i = 0; i < c.num_blocks; ++i) {
            const std::string p = "body." + std::to_string(i);
            const int slot = P->film_index.count(p) ? P->film_index[p] : -1;
            if (slot < 0) { bad("no time-embedding slot for block " + p); return; }
            ConvEpilogue e1;
            e1.film = film.base; e1.film_row = film.row; e1.film_row_stride = film.row_stride; e1.film_ld = P->film_ld;
            e1.film_off = P->film[slot].off; e1.film_has_scale = P->sr3 ? 0 : 1; e1.silu = P->wsplit ? 2 : 1;
            Act h = conv(p + ".conv.proj.weight", p + ".conv.proj.bias", x, nullptr, F, 3, CONV_TAPS, e1);
            ConvEpilogue e2;
            e2.out_scale = 0.1f; e2.res = x.p; e2.ldr = F;
            Act y = conv(p + ".conv.proj.weight", p + ".conv.proj.bias", h, nullptr, F, 3, CONV_TAPS, e2);
            free_act(h);
            if (!x_is_r) free_act(x);
            x_is_r = false;
            x = y;
            note(p, x);
        }
        ConvEpilogue et;
        et.res = r.p; et.ldr = F;
        Act bt = conv("body_tail.weight", "body_tail.bias", x, nullptr, F, 3, CONV_TAPS, et);
        if (!x_is_r) free_act(x);
        free_act(r);
        note("body_tail", bt);
        ConvEpilogue eo;
        eo.out_f32 = dry ? reinterpret_cast<float*>(1) : ex->eps;
        eo.n_valid = 1;
        conv("tail.weight", "tail.bias", bt, nullptr, 1, 3, CONV_TAPS, eo, 16);
        free_act(bt);
    }

    void build() {
        if (P->hicedrn) build_hicedrn(); else build_unet();
    }
};

int ensure_device(hd_plan* P) {
    CUDA_TRY(cudaSetDevice(P->device));
    return 0;
}

void free_exec(Exec* ex) {
    if (ex->g_eps) cudaGraphExecDestroy(ex->g_eps);
    if (ex->g_step) cudaGraphExecDestroy(ex->g_step);
    cudaFree(ex->arena); cudaFree(ex->x); cudaFree(ex->cond); cudaFree(ex->eps); cudaFree(ex->time);
    cudaFree(ex->posenc); cudaFree(ex->temb0); cudaFree(ex->temb); cudaFree(ex->film_rows); cudaFree(ex->iota);
    cudaFree(ex->ctx); cudaFree(ex->gn_part); cudaFree(ex->la_ctx); cudaFree(ex->la_s); cudaFree(ex->la_mb); cudaFree(ex->gn_cnt); cudaFree(ex->gn_ma);
}

int run_ops(const std::vector<Op>& ops, cudaStream_t s) {
    for (const Op& op : ops) {
        cudaError_t e = op.fn(s);
        if (e != cudaSuccess) return fail("launch of '%s' failed: %s", op.tag.c_str(), cudaGetErrorString(e));
    }
    return 0;
}

// Capture happens on a plan-private stream: the caller's stream may be the legacy default stream, which cannot be
// captured; the instantiated graph is then launched on whatever stream the caller passes.
int capture(const std::vector<Op>& ops, cudaStream_t s, cudaGraphExec_t* out) {
    NvtxRange range("hd: graph capture");
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = run_ops(ops, s);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return 0;
}

int get_exec(hd_plan* P, int B, cudaStream_t s, Exec** out) {
    if (!P->finalized) return fail("hd_plan_finalize has not been called");
    if (B <= 0) return fail("batch size must be positive (got %d)", B);
    auto it = P->execs.find(B);
    if (it != P->execs.end()) { *out = it->second.get(); return 0; }
    NvtxRange range("hd: build executor (arena, launch list, graphs)");
    std::unique_ptr<Exec> ex(new Exec());
    ex->B = B;
    const int S = P->cfg.image_size;
    const size_t tile = static_cast<size_t>(S) * S;

    // dry run sizes the arena
    {
        Builder d{P, ex.get(), B, true};
        d.arena.never_free = P->cfg.debug_keep != 0;
        std::vector<Op> none;
        d.ops = &none;
        d.film = FilmRef{nullptr, nullptr, 0};
        d.build();
        if (!d.ok) return fail("%s", d.err.c_str());
        ex->arena_bytes = d.arena.high;
    }
    auto cleanup_fail = [&](int rc) { free_exec(ex.get()); return rc; };
#define EX_TRY(expr)                                                                                      \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) return cleanup_fail(fail("%s failed: %s", #expr, cudaGetErrorString(e__))); \
    } while (0)
    EX_TRY(cudaMalloc(&ex->arena, ex->arena_bytes));
    EX_TRY(cudaMalloc(&ex->x, B * tile * 4));
    EX_TRY(cudaMalloc(&ex->cond, B * tile * 4));
    EX_TRY(cudaMalloc(&ex->eps, B * tile * 4));
    EX_TRY(cudaMalloc(&ex->time, B * 4));
    EX_TRY(cudaMalloc(&ex->posenc, static_cast<size_t>(B) * P->fourier_dim * 4));
    EX_TRY(cudaMalloc(&ex->temb0, static_cast<size_t>(B) * P->time_dim * 4));
    EX_TRY(cudaMalloc(&ex->temb, static_cast<size_t>(B) * P->time_dim * 4));
    EX_TRY(cudaMalloc(&ex->film_rows, static_cast<size_t>(B) * P->film_ld * 4));
    EX_TRY(cudaMalloc(&ex->iota, B * 4));
    EX_TRY(cudaMalloc(&ex->ctx, static_cast<size_t>(B) * 4 * 32 * 32 * 4));
    EX_TRY(cudaMalloc(&ex->gn_part, static_cast<size_t>(B) * (tile / 32) * 64 * sizeof(float2)));   // [M/32][C/8], C <= 512
    EX_TRY(cudaMalloc(&ex->gn_cnt, static_cast<size_t>(B) * 4));
    EX_TRY(cudaMalloc(&ex->gn_ma, static_cast<size_t>(B) * 512 * sizeof(float2)));
    if (!P->lattn.empty()) {
        EX_TRY(cudaMalloc(&ex->la_ctx, static_cast<size_t>(B) * LA_MAX_PARTS * 128 * 32 * 4));
        EX_TRY(cudaMalloc(&ex->la_s, static_cast<size_t>(B) * LA_MAX_PARTS * 128 * 4));
        EX_TRY(cudaMalloc(&ex->la_mb, static_cast<size_t>(B) * 128 * 128 * 2));
    }
    EX_TRY(cudaMemsetAsync(ex->x, 0, B * tile * 4, s));
    EX_TRY(cudaMemsetAsync(ex->cond, 0, B * tile * 4, s));
    EX_TRY(cudaMemsetAsync(ex->time, 0, B * 4, s));
    ex->extra_bytes = 3 * B * tile * 4 + static_cast<size_t>(B) * (P->fourier_dim + 2 * P->time_dim + P->film_ld + 2 + 4096) * 4;
    {
        std::vector<int> h(B);
        for (int i = 0; i < B; ++i) h[i] = i;
        EX_TRY(cudaMemcpyAsync(ex->iota, h.data(), B * 4, cudaMemcpyHostToDevice, s));
        EX_TRY(cudaStreamSynchronize(s));
    }

    // rows mode: time embedding evaluated per call for B rows (Unet.forward with arbitrary `time`)
    if (enqueue_film(P, ex->time, B, ex->posenc, ex->temb0, ex->temb, ex->film_rows, &ex->ops_rows)) return cleanup_fail(1);
    {
        Builder b{P, ex.get(), B, false};
        b.arena.never_free = P->cfg.debug_keep != 0;
        b.ops = &ex->ops_rows;
        b.film = FilmRef{ex->film_rows, ex->iota, 1};
        b.build();
        if (!b.ok) return cleanup_fail(fail("%s", b.err.c_str()));
    }
    // table mode: FiLM row picked by the device step counter; + posterior update + counter decrement
    {
        Builder b{P, ex.get(), B, false};
        b.arena.never_free = P->cfg.debug_keep != 0;
        b.ops = &ex->ops_table;
        b.film = FilmRef{P->film_table, &P->ctl->step, 0};
        const bool keep = P->cfg.debug_keep != 0;
        P->cfg.debug_keep = 0;   // debug names refer to the rows-mode graph
        b.build();
        P->cfg.debug_keep = keep ? 1 : 0;
        if (!b.ok) return cleanup_fail(fail("%s", b.err.c_str()));
        PosteriorArgs pa;
        pa.x = ex->x; pa.eps = ex->eps; pa.coef = P->coef; pa.ctl = P->ctl; pa.T = P->T;
        pa.n = static_cast<long long>(B) * tile; pa.tile_elems = static_cast<int>(tile); pa.x0_out = nullptr;
        Op pop{[pa](cudaStream_t st) { return posterior_step_run(pa, st); }, "posterior"};
        pop.kernel = "posterior_step";
        pop.bytes = 12.0 * pa.n;   // x read + eps read + x write (Philox noise is generated in-kernel)
        ex->ops_table.push_back(pop);
        SampleCtl* ctl = P->ctl;
        Op aop{[ctl](cudaStream_t st) { return step_advance_run(ctl, -1, st); }, "step--"};
        aop.kernel = "step_advance";
        ex->ops_table.push_back(aop);
    }
    // eager validation pass (surfaces launch-configuration errors with the op name), then capture
    {
        SampleCtl h;
        memset(&h, 0, sizeof(h));
        h.step = 0;
        EX_TRY(cudaMemcpyAsync(P->ctl, &h, sizeof(h), cudaMemcpyHostToDevice, s));
        if (run_ops(ex->ops_rows, s)) return cleanup_fail(1);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return cleanup_fail(fail("eps-net validation run failed: %s", cudaGetErrorString(e)));
        if (run_ops(ex->ops_table, s)) return cleanup_fail(1);
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return cleanup_fail(fail("sampling-step validation run failed: %s", cudaGetErrorString(e)));
        if (capture(ex->ops_rows, P->cap_stream, &ex->g_eps)) return cleanup_fail(1);
        if (capture(ex->ops_table, P->cap_stream, &ex->g_step)) return cleanup_fail(1);
    }
#undef EX_TRY
    *out = ex.get();
    P->execs[B] = std::move(ex);
    return 0;
}

int set_ctl(SampleCtl* dev, int step, int single, const float* noise, uint64_t seed, uint64_t tile_offset,
            cudaStream_t s) {
    SampleCtl h;
    memset(&h, 0, sizeof(h));
    h.step = step; h.noise_single = single; h.noise = noise; h.seed = seed; h.tile_offset = tile_offset;
    CUDA_TRY(cudaMemcpyAsync(dev, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    return 0;
}

}  // namespace

namespace hd {
int set_error(const char* msg) { g_err = msg ? msg : ""; return 1; }
// Programmatic dependent launch (kernels.h::launch_pdl) of the forward-path kernels (conv_gemm, GroupNorm / LayerNorm apply, the
// attention kernels, stem / head convs, the posterior step).  Measured on B200, same box, three repeats (profiles/r02_notes.md 8):
//   training step (Unet, ~820 launches):  plain 11.96 - 11.98 ms | these kernels as programmatic dependents 11.78 - 11.80 ms | EVERY kernel
//     of the step (a mechanical pass over the 63 backward launch sites, not kept) 12.1 - 13.6 ms;  hicedrn_Diff 58.29 / 58.29 / 59.16 ms
//   sampling step (134 launches of mostly persistent kernels):  plain 5.99 ms | programmatic 6.15 ms
// -> default: on only while a trainer enqueues its step (PdlScope); HD_PDL=1 forces it on everywhere, HD_PDL=0 off.
static thread_local int g_pdl_scope = 0;
PdlScope::PdlScope() { ++g_pdl_scope; }
PdlScope::~PdlScope() { --g_pdl_scope; }
bool pdl_enabled() {
    static const int mode = [] { const char* v = getenv("HD_PDL"); return !v ? -1 : (v[0] == '0' ? 0 : 1); }();
    return mode < 0 ? g_pdl_scope > 0 : mode == 1;
}   // trainer.cu reports through the same hd_last_error()
}

// ================================================================================================= C ABI
extern "C" {

const char* hd_last_error(void) { return g_err.c_str(); }
int hd_abi_version(void) { return HD_ABI_VERSION; }

int hd_plan_create(const hd_config* cfg, hd_plan** out) {
    if (!cfg || !out) return fail("hd_plan_create: null argument");
    if (cfg->abi_version != HD_ABI_VERSION) return fail("ABI version mismatch: header %d, library %d", cfg->abi_version, HD_ABI_VERSION);
    if (cfg->variant < HD_UNET || cfg->variant > HD_HICEDRN_SR3) return fail("unknown variant %d", cfg->variant);
    if (cfg->image_size != 64) return fail("image_size must be 64 (the reference's piece_size); got %d", cfg->image_size);
    if (cfg->timesteps <= 0) return fail("timesteps must be positive");
    const bool hic = cfg->variant == HD_HICEDRN || cfg->variant == HD_HICEDRN_SR3;
    if (!hic) {
        if (cfg->dim <= 0 || cfg->dim % 64 != 0) return fail("Unet dim must be a positive multiple of 64 (got %d)", cfg->dim);
        if (cfg->num_mults < 1 || cfg->num_mults > 4) return fail("len(dim_mults) must be 1..4 for 64x64 tiles (got %d)", cfg->num_mults);
        for (int i = 0; i < cfg->num_mults; ++i)
            if (cfg->dim_mults[i] < 1 || cfg->dim * cfg->dim_mults[i] > 512)
                return fail("dim * dim_mults[%d] = %d exceeds the supported 512 channels", i, cfg->dim * cfg->dim_mults[i]);
    } else if (cfg->num_blocks < 1) {
        return fail("HiCEDRN needs num_blocks >= 1");
    }
    int dev = 0, cc_major = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major != 10) return fail("hicdiff_b200 needs an sm_100a GPU (compute capability 10.x); device %d is %d.x", dev, cc_major);
    hd_plan* P = new hd_plan();
    P->cfg = *cfg;
    P->device = dev;
    P->num_sms = sms;
    P->sr3 = cfg->variant == HD_UNET_SR3 || cfg->variant == HD_HICEDRN_SR3;
    P->hicedrn = hic;
    {
        const int precision = (cfg->reserved[0] >> 4) & 3;
        if (precision > 1) { delete P; return fail("unknown precision mode %d (0 = bf16, 1 = bf16 activations + split hi/lo weights)", precision); }
        P->wsplit = precision == 1;
    }
    P->T = cfg->timesteps;
    P->fourier_dim = hic ? 256 : cfg->dim;
    P->time_dim = hic ? 1024 : cfg->dim * 4;
    cudaError_t se = cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete P; return fail("cudaStreamCreateWithFlags failed: %s", cudaGetErrorString(se)); }
    *out = P;
    return 0;
}

int hd_plan_set_weight(hd_plan* P, const char* key, const float* dev_ptr, const int64_t* shape, int32_t ndim, void* stream) {
    if (!P || !key || !dev_ptr || (ndim > 0 && !shape)) return fail("hd_plan_set_weight: null argument");
    if (ensure_device(P)) return 1;
    size_t numel = 1;
    std::vector<int64_t> sh;
    for (int i = 0; i < ndim; ++i) { numel *= static_cast<size_t>(shape[i]); sh.push_back(shape[i]); }
    WeightT& w = P->w[key];
    if (w.d != nullptr && w.numel != numel) { cudaFree(w.d); w.d = nullptr; P->weight_bytes -= w.numel * 4; }
    if (w.d == nullptr) { CUDA_TRY(cudaMalloc(&w.d, numel * 4)); P->weight_bytes += numel * 4; }
    w.shape = sh;
    w.numel = numel;
    CUDA_TRY(cudaMemcpyAsync(w.d, dev_ptr, numel * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    P->finalized = false;
    return 0;
}

int hd_plan_set_schedule(hd_plan* P, const float* sqrt_recip, const float* sqrt_recipm1, const float* coef1,
                         const float* coef2, const float* sigma, const float* time_values, int32_t T, void* stream) {
    if (!P || !sqrt_recip || !sqrt_recipm1 || !coef1 || !coef2 || !sigma || !time_values) return fail("hd_plan_set_schedule: null argument");
    if (T != P->T) return fail("schedule length %d does not match the plan's timesteps %d", T, P->T);
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!P->coef) { CUDA_TRY(cudaMalloc(&P->coef, static_cast<size_t>(T) * 8 * 4)); }
    if (!P->time_values) { CUDA_TRY(cudaMalloc(&P->time_values, static_cast<size_t>(T) * 4)); }
    CUDA_TRY(cudaMemsetAsync(P->coef, 0, static_cast<size_t>(T) * 8 * 4, s));
    const float* cols[5] = {sqrt_recip, sqrt_recipm1, coef1, coef2, sigma};
    for (int c = 0; c < 5; ++c)
        CUDA_TRY(cudaMemcpy2DAsync(P->coef + c, 8 * 4, cols[c], 4, 4, T, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(P->time_values, time_values, static_cast<size_t>(T) * 4, cudaMemcpyDeviceToDevice, s));
    P->finalized = false;
    return 0;
}

int hd_plan_finalize(hd_plan* P, void* stream) {
    NvtxRange range("hd_plan_finalize (weight standardisation, GEMM layouts, FiLM table)");
    if (!P) return fail("hd_plan_finalize: null plan");
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!P->coef) return fail("hd_plan_set_schedule must be called before hd_plan_finalize");
    // drop executors built against older weights
    for (auto& kv : P->execs) free_exec(kv.second.get());
    P->execs.clear();

    // ---- FiLM slots (one per ResnetBlock, in module order)
    P->film.clear();
    P->film_index.clear();
    std::vector<std::pair<std::string, int>> blocks;   // prefix, Cout
    if (P->hicedrn) {
        for (int i = 0; i < P->cfg.num_blocks; ++i) blocks.push_back({"body." + std::to_string(i), 256});
    } else {
        const hd_config& c = P->cfg;
        std::vector<int> dims;
        dims.push_back(c.dim);
        for (int i = 0; i < c.num_mults; ++i) dims.push_back(c.dim * c.dim_mults[i]);
        const int L = c.num_mults;
        for (int i = 0; i < L; ++i) {
            blocks.push_back({"downs." + std::to_string(i) + ".0", dims[i]});
            blocks.push_back({"downs." + std::to_string(i) + ".1", dims[i]});
        }
        blocks.push_back({"mid_block1", dims[L]});
        blocks.push_back({"mid_block2", dims[L]});
        for (int k = 0; k < L; ++k) {
            blocks.push_back({"ups." + std::to_string(k) + ".0", dims[L - k]});
            blocks.push_back({"ups." + std::to_string(k) + ".1", dims[L - k]});
        }
        blocks.push_back({"final_res_block", c.dim});
    }
    int off = 0;
    for (auto& b : blocks) {
        FilmSlot f;
        if (P->sr3) {
            f.wkey = b.first + ".noise_func.noise_func.0.weight";
            f.bkey = b.first + ".noise_func.noise_func.0.bias";
            f.width = b.second;
            f.silu_in = false;
        } else {
            f.wkey = b.first + ".mlp.1.weight";
            f.bkey = b.first + ".mlp.1.bias";
            f.width = 2 * b.second;
            f.silu_in = true;
        }
        f.off = off;
        off += f.width;
        const WeightT* fw = find_w(P, f.wkey);
        if (!fw) return fail("weight '%s' missing (did you call hd_plan_set_weight for every state_dict entry?)", f.wkey.c_str());
        if (fw->shape.size() != 2 || fw->shape[0] != f.width || fw->shape[1] != P->time_dim)
            return fail("weight '%s' has an unexpected shape", f.wkey.c_str());
        P->film_index[b.first] = static_cast<int>(P->film.size());
        P->film.push_back(f);
    }
    P->film_ld = off;

    // ---- GEMM-layout weights
    for (auto& kv : P->wq) cudaFree(kv.second);
    P->wq.clear();
    for (auto& kv : P->padded) cudaFree(kv.second);
    P->padded.clear();
    size_t wq_bytes = 0;
    for (auto& kv : P->w) {
        const std::string& key = kv.first;
        const WeightT& w = kv.second;
        if (w.shape.size() != 4 || !ends_with(key, ".weight")) continue;
        const int Cout = static_cast<int>(w.shape[0]), Cin = static_cast<int>(w.shape[1]), k = static_cast<int>(w.shape[2]);
        if (Cin < 64 || key == "final_conv.weight") continue;   // stem / 1x1 head run on dedicated kernels
        if (Cin % 64 != 0) return fail("conv weight '%s': Cin = %d is not a multiple of 64", key.c_str(), Cin);
        const int Npad = Cout < 16 ? 16 : Cout;
        bf16* q = nullptr;
        const int split = P->wsplit ? 1 : 0;
        const size_t bytes = static_cast<size_t>(Npad) * Cin * k * k * 2 * (split + 1);
        CUDA_TRY(cudaMalloc(&q, bytes));
        wq_bytes += bytes;
        P->wq[key] = q;
        const bool is_down = !P->hicedrn && key.compare(0, 6, "downs.") == 0 && ends_with(key, ".3.1.weight");
        const bool is_up = !P->hicedrn && key.compare(0, 4, "ups.") == 0 && ends_with(key, ".3.1.weight");
        if (is_up) {
            if (k != 3) return fail("Upsample weight '%s' has an unexpected shape", key.c_str());
            cudaFree(q);
            const size_t ub = static_cast<size_t>(4) * Cout * 4 * Cin * 2 * (split + 1);   // four phase matrices [Cout, 4*Cin]
            CUDA_TRY(cudaMalloc(&q, ub));
            P->wq[key] = q;
            CUDA_TRY(prep_upsample_weight_run(w.d, q, Cout, Cin, s, split));
        } else if (is_down) {
            if (k != 1 || Cin % 4 != 0) return fail("Downsample weight '%s' has an unexpected shape", key.c_str());
            CUDA_TRY(prep_unshuffle_weight_run(w.d, q, Cout, Cin / 4, s, split));
        } else {
            const int standardize = (!P->hicedrn && ends_with(key, ".proj.weight")) ? 1 : 0;
            CUDA_TRY(prep_conv_weight_run(w.d, q, Cout, Cin, k, standardize, 1e-5f, Npad, s, split));
        }
        if (Npad != Cout) {
            const std::string bkey = key.substr(0, key.size() - 6) + "bias";
            const WeightT* b = find_w(P, bkey);
            if (b) {
                float* pb = nullptr;
                CUDA_TRY(cudaMalloc(&pb, Npad * 4));
                CUDA_TRY(cudaMemsetAsync(pb, 0, Npad * 4, s));
                CUDA_TRY(cudaMemcpyAsync(pb, b->d, Cout * 4, cudaMemcpyDeviceToDevice, s));
                P->padded[bkey] = pb;
            }
        }
    }

    // ---- fused linear-attention blocks: fold the PreNorm gain into to_qkv, row sums, softmax shifts
    for (auto& kv : P->lattn) { cudaFree(kv.second.wqkv); cudaFree(kv.second.rowsum); cudaFree(kv.second.kshift); cudaFree(kv.second.bound_bits); }
    P->lattn.clear();
    if (!P->hicedrn) {
        const char* suf = ".fn.fn.to_qkv.weight";
        for (auto& kv : P->w) {
            const std::string& key = kv.first;
            if (!ends_with(key, suf) || key.compare(0, 8, "mid_attn") == 0) continue;
            const std::string pre = key.substr(0, key.size() - strlen(suf));
            const int C = static_cast<int>(kv.second.shape[1]);
            const WeightT* g1 = find_w(P, pre + ".fn.norm.g");
            if ((C != 64 && C != 128) || kv.second.shape[0] != 384 || !g1) continue;
            LinAttnPrep lp;
            lp.C = C;
            CUDA_TRY(cudaMalloc(&lp.wqkv, static_cast<size_t>(384) * C * 2));
            CUDA_TRY(cudaMalloc(&lp.rowsum, 384 * 4));
            CUDA_TRY(cudaMalloc(&lp.kshift, 128 * 4));
            CUDA_TRY(cudaMalloc(&lp.bound_bits, 4));
            CUDA_TRY(linattn_prep_run(kv.second.d, g1->d, lp.wqkv, lp.rowsum, lp.kshift, lp.bound_bits, C, s));
            P->lattn[pre] = lp;
        }
        CUDA_TRY(cudaStreamSynchronize(s));
        for (auto& kv : P->lattn) {
            int bits = 0;
            CUDA_TRY(cudaMemcpy(&bits, kv.second.bound_bits, 4, cudaMemcpyDeviceToHost));
            memcpy(&kv.second.bound, &bits, 4);
        }
    }

    // ---- control blocks + [T, film_ld] table
    if (!P->ctl) { CUDA_TRY(cudaMalloc(&P->ctl, sizeof(SampleCtl))); CUDA_TRY(cudaMemsetAsync(P->ctl, 0, sizeof(SampleCtl), s)); }
    if (!P->ctl_one) { CUDA_TRY(cudaMalloc(&P->ctl_one, sizeof(SampleCtl))); CUDA_TRY(cudaMemsetAsync(P->ctl_one, 0, sizeof(SampleCtl), s)); }
    if (P->film_table) { cudaFree(P->film_table); P->film_table = nullptr; }
    CUDA_TRY(cudaMalloc(&P->film_table, static_cast<size_t>(P->T) * P->film_ld * 4));
    {
        float *pe = nullptr, *t0 = nullptr, *t1 = nullptr;
        CUDA_TRY(cudaMalloc(&pe, static_cast<size_t>(P->T) * P->fourier_dim * 4));
        CUDA_TRY(cudaMalloc(&t0, static_cast<size_t>(P->T) * P->time_dim * 4));
        CUDA_TRY(cudaMalloc(&t1, static_cast<size_t>(P->T) * P->time_dim * 4));
        std::vector<Op> ops;
        int rc = enqueue_film(P, P->time_values, P->T, pe, t0, t1, P->film_table, &ops);
        if (!rc) rc = run_ops(ops, s);
        cudaError_t e = cudaStreamSynchronize(s);
        cudaFree(pe); cudaFree(t0); cudaFree(t1);
        if (rc) return rc;
        if (e != cudaSuccess) return fail("time-embedding table build failed: %s", cudaGetErrorString(e));
    }
    P->weight_bytes += 0;
    (void)wq_bytes;
    P->finalized = true;
    return 0;
}

void hd_plan_destroy(hd_plan* P) {
    if (!P) return;
    cudaSetDevice(P->device);
    for (auto& kv : P->execs) free_exec(kv.second.get());
    for (auto& kv : P->w) cudaFree(kv.second.d);
    for (auto& kv : P->wq) cudaFree(kv.second);
    for (auto& kv : P->padded) cudaFree(kv.second);
    for (auto& kv : P->lattn) { cudaFree(kv.second.wqkv); cudaFree(kv.second.rowsum); cudaFree(kv.second.kshift); cudaFree(kv.second.bound_bits); }
    cudaFree(P->coef); cudaFree(P->time_values); cudaFree(P->film_table); cudaFree(P->ctl); cudaFree(P->ctl_one);
    if (P->cap_stream) cudaStreamDestroy(P->cap_stream);
    delete P;
}

int hd_eps_forward(hd_plan* P, const float* x, const float* cond, const float* time, float* eps, int32_t B, void* stream) {
    NvtxRange range("hd_eps_forward");
    if (!P || !x || !time || !eps) return fail("hd_eps_forward: null argument");
    if (P->cfg.self_condition && !cond) return fail("hd_eps_forward: the plan is self-conditioned, cond must not be NULL");
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Exec* ex = nullptr;
    if (get_exec(P, B, s, &ex)) return 1;
    const size_t bytes = static_cast<size_t>(B) * 4096 * 4;
    CUDA_TRY(cudaMemcpyAsync(ex->x, x, bytes, cudaMemcpyDeviceToDevice, s));
    if (cond) CUDA_TRY(cudaMemcpyAsync(ex->cond, cond, bytes, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(ex->time, time, B * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaGraphLaunch(ex->g_eps, s));
    CUDA_TRY(cudaMemcpyAsync(eps, ex->eps, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int hd_ddpm_step(hd_plan* P, float* x, const float* eps, const float* noise, float* x0_out, int32_t t, int32_t B,
                 uint64_t seed, uint64_t tile_offset, void* stream) {
    NvtxRange range("hd_ddpm_step");
    if (!P || !x || !eps) return fail("hd_ddpm_step: null argument");
    if (!P->coef || !P->ctl_one) return fail("hd_ddpm_step: plan not finalized");
    if (t < 0 || t >= P->T) return fail("hd_ddpm_step: t = %d out of range [0, %d)", t, P->T);
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (set_ctl(P->ctl_one, t, 1, noise, seed, tile_offset, s)) return 1;
    PosteriorArgs pa;
    pa.x = x; pa.eps = eps; pa.coef = P->coef; pa.ctl = P->ctl_one; pa.T = P->T;
    pa.n = static_cast<long long>(B) * 4096; pa.tile_elems = 4096; pa.x0_out = x0_out;
    CUDA_TRY(posterior_step_run(pa, s));
    return 0;
}

int hd_sample(hd_plan* P, const float* cond, const float* noise, float* out, float* trace, int32_t B, uint64_t seed,
              uint64_t tile_offset, int32_t t_start, int32_t t_end, void* stream) {
    NvtxRange range("hd_sample (reverse chain)");
    if (!P || !out) return fail("hd_sample: null argument");
    if (P->cfg.self_condition && !cond) return fail("hd_sample: the plan is self-conditioned, cond must not be NULL");
    if (t_start >= P->T || t_end < 0 || t_end > t_start) return fail("hd_sample: bad step range [%d, %d] for T = %d", t_start, t_end, P->T);
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Exec* ex = nullptr;
    if (get_exec(P, B, s, &ex)) return 1;
    const size_t n = static_cast<size_t>(B) * 4096;
    if (cond) CUDA_TRY(cudaMemcpyAsync(ex->cond, cond, n * 4, cudaMemcpyDeviceToDevice, s));
    if (t_start == P->T - 1) {
        if (noise) CUDA_TRY(cudaMemcpyAsync(ex->x, noise, n * 4, cudaMemcpyDeviceToDevice, s));
        else CUDA_TRY(philox_normal_run(ex->x, static_cast<long long>(n), seed, tile_offset, 4096, 0, s));
    } else {
        CUDA_TRY(cudaMemcpyAsync(ex->x, out, n * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (set_ctl(P->ctl, t_start, 0, noise, seed, tile_offset, s)) return 1;
    for (int t = t_start; t >= t_end; --t) {
        CUDA_TRY(cudaGraphLaunch(ex->g_step, s));
        if (trace) CUDA_TRY(cudaMemcpyAsync(trace + static_cast<size_t>(t_start - t) * n, ex->x, n * 4, cudaMemcpyDeviceToDevice, s));
    }
    CUDA_TRY(cudaMemcpyAsync(out, ex->x, n * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int hd_ddrm_step(float* x, const float* eps, const float* y, const float* noise, float* x0_out, int32_t mode, float sqrt_at,
                 float sqrt_1m_at, float sqrt_at_next, float c0, float c1, float c2, float sigma_0, int64_t n, uint64_t seed,
                 uint64_t tile_offset, uint32_t step_id, void* stream) {
    if (!x || !eps || !y) return fail("hd_ddrm_step: null argument");
    if (mode < 0 || mode > 2) return fail("hd_ddrm_step: mode must be 0 (before), 1 (after) or 2 (equal), got %d", mode);
    if (n <= 0 || n % 4096 != 0) return fail("hd_ddrm_step: n = %lld is not a whole number of 64x64 tiles", static_cast<long long>(n));
    DdrmArgs a;
    a.x = x; a.eps = eps; a.y = y; a.noise = noise; a.x0_out = x0_out; a.mode = mode;
    a.sqrt_at = sqrt_at; a.sqrt_1m_at = sqrt_1m_at; a.sqrt_at_next = sqrt_at_next; a.c0 = c0; a.c1 = c1; a.c2 = c2; a.sigma_0 = sigma_0;
    a.n = n; a.tile_elems = 4096; a.seed = seed; a.tile_offset = tile_offset; a.step_id = step_id;
    CUDA_TRY(ddrm_step_run(a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_ddim_step(float* x, const float* eps, const float* noise, float* x0_out, float sqrt_recip, float sqrt_recipm1, float sqrt_a_next,
                 float c, float sigma, int32_t last, int64_t n, uint64_t seed, uint64_t tile_offset, uint32_t step_id, void* stream) {
    if (!x || !eps || n <= 0) return fail("hd_ddim_step: bad argument");
    DdimArgs a;
    a.x = x; a.eps = eps; a.noise = noise; a.x0_out = x0_out;
    a.sr = sqrt_recip; a.srm1 = sqrt_recipm1; a.sqrt_a_next = sqrt_a_next; a.c = c; a.sigma = sigma; a.last = last ? 1 : 0;
    a.n = n; a.tile_elems = 4096; a.seed = seed; a.tile_offset = tile_offset; a.step_id = step_id;
    CUDA_TRY(ddim_step_run(a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_ssim_mse_tiles(const float* a, const float* b, const float* window121, float* ssim_out, float* mse_out, int32_t B,
                      int32_t rescale, void* stream) {
    if (B < 0) return fail("hd_ssim_mse_tiles: negative tile count");
    if (B == 0) return 0;
    if (!a || !b || !window121 || !ssim_out || !mse_out) return fail("hd_ssim_mse_tiles: null argument");
    CUDA_TRY(ssim_mse_tiles_run(a, b, window121, ssim_out, mse_out, B, rescale, static_cast<cudaStream_t>(stream)));
    return 0;
}

// ---- data preparation (dataprep.cu)
int hd_coo_to_dense(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz, int64_t smallbin, int64_t n,
                    float* mat, void* stream) {
    if (n < 0 || nnz < 0) return fail("hd_coo_to_dense: negative size");
    if (n == 0) return 0;
    if (!mat || (nnz > 0 && (!rows || !cols || !vals))) return fail("hd_coo_to_dense: null argument");
    if (nnz > 2147483647LL) return fail("hd_coo_to_dense: more than 2^31 - 1 triples");
    void* scratch = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, coo_scratch_bytes(n)));
    int bad = 0;
    cudaError_t e = coo_to_dense_run(reinterpret_cast<const long long*>(rows), reinterpret_cast<const long long*>(cols), vals, nnz,
                                     smallbin, n, mat, scratch, &bad, static_cast<cudaStream_t>(stream));
    cudaFree(scratch);
    if (e != cudaSuccess) return fail("hd_coo_to_dense failed: %s", cudaGetErrorString(e));
    if (bad) return fail("hd_coo_to_dense: a (row, col) pair lies outside [smallbin, smallbin + n)");
    return 0;
}

int hd_remove_empty_bins(const float* mat, int64_t n, float* out, int64_t* kept_idx, int64_t* n_kept_host, void* stream) {
    if (n < 0 || !n_kept_host) return fail("hd_remove_empty_bins: bad argument");
    *n_kept_host = 0;
    if (n == 0) return 0;
    if (!mat || !out || !kept_idx) return fail("hd_remove_empty_bins: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long* cnt = nullptr;
    CUDA_TRY(cudaMalloc(&cnt, 8));
    cudaError_t e = keep_map_run(mat, n, reinterpret_cast<long long*>(kept_idx), cnt, s);
    long long m = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&m, cnt, 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = compact_run(mat, n, reinterpret_cast<const long long*>(kept_idx), m, out, s);
    cudaFree(cnt);
    if (e != cudaSuccess) return fail("hd_remove_empty_bins failed: %s", cudaGetErrorString(e));
    *n_kept_host = m;
    return 0;
}

int hd_select_ranks(const float* x, int64_t n, const int64_t* ranks_host, int32_t nranks, float* out_host, void* stream) {
    if (!x || !ranks_host || !out_host || n <= 0 || nranks <= 0) return fail("hd_select_ranks: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    void* scratch = nullptr;
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(&scratch, select_scratch_bytes()));
    CUDA_TRY(cudaMalloc(&out, static_cast<size_t>(nranks) * 4));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < nranks && e == cudaSuccess; ++i) {
        if (ranks_host[i] < 0 || ranks_host[i] >= n) { cudaFree(scratch); cudaFree(out); return fail("hd_select_ranks: rank %lld outside [0, %lld)", (long long)ranks_host[i], (long long)n); }
        e = select_rank_run(x, n, static_cast<unsigned long long>(ranks_host[i]), out + i, scratch, s);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, out, static_cast<size_t>(nranks) * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(scratch); cudaFree(out);
    if (e != cudaSuccess) return fail("hd_select_ranks failed: %s", cudaGetErrorString(e));
    return 0;
}

int hd_normalize_contacts(float* x, int64_t n, float per, void* stream) {
    if (n < 0 || (n > 0 && !x)) return fail("hd_normalize_contacts: bad argument");
    CUDA_TRY(normalize_contacts_run(x, n, per, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_add_noise(const float* x, const float* noise, float sigma, int64_t n, float* out, void* stream) {
    if (n < 0 || (n > 0 && (!x || !noise || !out))) return fail("hd_add_noise: bad argument");
    CUDA_TRY(axpy_noise_run(x, noise, sigma, n, out, static_cast<cudaStream_t>(stream)));
    return 0;
}

int64_t hd_tile_count(int64_t n, int32_t piece, int32_t band_blocks) {
    if (n < 0 || piece <= 0 || band_blocks < 0) return -1;
    return tile_count(static_cast<int>(n), piece, band_blocks);
}

int hd_tile_extract(const float* mat, int64_t n, float* tiles, int32_t piece, int32_t band_blocks, void* stream) {
    if (n < 0 || piece <= 0 || band_blocks < 0) return fail("hd_tile_extract: bad arguments");
    if (n == 0) return 0;
    if (!mat || !tiles) return fail("hd_tile_extract: null argument");
    CUDA_TRY(tile_extract_run(mat, static_cast<int>(n), tiles, piece, band_blocks, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_tile_scatter(const float* tiles, float* mat, int64_t n, int32_t piece, int32_t band_blocks, void* stream) {
    if (n < 0 || piece <= 0 || band_blocks < 0) return fail("hd_tile_scatter: bad arguments");
    if (n == 0) return 0;
    if (!mat || !tiles) return fail("hd_tile_scatter: null argument");
    CUDA_TRY(tile_scatter_run(tiles, mat, static_cast<int>(n), piece, band_blocks, static_cast<cudaStream_t>(stream)));
    return 0;
}

// ------------------------------------------------------------------------------------------------- single ops
int hd_op_conv2d(const uint16_t* x0, int32_t C0, const uint16_t* x1, int32_t C1, const float* w, const float* bias,
                 const uint16_t* res, uint16_t* out, int32_t B, int32_t H, int32_t W, int32_t Cout, int32_t ksize,
                 int32_t mode, int32_t standardize, void* stream) {
    if (!x0 || !w || !out) return fail("hd_op_conv2d: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int Cin = C0 + (x1 ? C1 : 0);
    const int taps = mode == CONV_UNSHUFFLE ? 4 : (mode == CONV_UPSAMPLE ? 16 : ksize * ksize);
    if (mode == CONV_UPSAMPLE && (ksize != 3 || x1 != nullptr)) return fail("hd_op_conv2d: upsample mode is a 3x3 conv of one source");
    bf16* q = nullptr;
    CUDA_TRY(cudaMalloc(&q, static_cast<size_t>(Cout) * Cin * taps * 2));
    cudaError_t e = mode == CONV_UNSHUFFLE ? prep_unshuffle_weight_run(w, q, Cout, Cin, s)
                    : mode == CONV_UPSAMPLE ? prep_upsample_weight_run(w, q, Cout, Cin, s)
                                            : prep_conv_weight_run(w, q, Cout, Cin, ksize, standardize & 1, 1e-5f, Cout, s);
    if (e != cudaSuccess) { cudaFree(q); return fail("weight prep failed: %s", cudaGetErrorString(e)); }
    ConvGemmDesc d;
    d.src0 = ConvSrc{reinterpret_cast<const bf16*>(x0), C0};
    d.src1 = ConvSrc{reinterpret_cast<const bf16*>(x1), x1 ? C1 : 0};
    d.B = B; d.H = H; d.W = W; d.ksize = ksize; d.mode = static_cast<ConvMode>(mode);
    d.weight = q; d.N = Cout; d.out = reinterpret_cast<bf16*>(out);
    d.pad_mode = (standardize >> 1) & 3;   // bits 1-2 of `standardize`: opt into the padded-slab conv form (parity tests)
    d.cg2_mode = (standardize >> 3) & 1;   // bit 3: run on CTA pairs (tcgen05 cta_group::2)
    {   // bits 4-5: form of the 3x3, Cout = 64 conv: 0 = default (dx-stacked, two epilogue groups), 1 = one MMA per tap, 2 = dx-stacked, one group
        const int f = (standardize >> 4) & 3;
        d.dx3_mode = f == 1 ? 0 : (f == 2 ? 1 : 2);
    }
    d.epi.bias = bias;
    if (res) { d.epi.res = reinterpret_cast<const bf16*>(res); d.epi.ldr = Cout; }
    ConvGemmLaunch l;
    char msg[256];
    if (conv_gemm_prepare(d, sms, &l, msg, sizeof(msg))) { cudaFree(q); return fail("%s", msg); }
    e = conv_gemm_run(l, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(q);
    if (e != cudaSuccess) return fail("conv launch failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail("conv kernel failed: %s", cudaGetErrorString(e2));
    return 0;
}

int hd_op_conv_gn(const uint16_t* x0, int32_t C0, const uint16_t* x1, int32_t C1, const float* w, const float* bias,
                  const float* gamma, const float* beta, const float* scale, const float* shift, const uint16_t* res,
                  uint16_t* out, int32_t B, int32_t H, int32_t W, int32_t Cout, int32_t standardize, void* stream) {
    if (!x0 || !w || !bias || !gamma || !beta || !out) return fail("hd_op_conv_gn: null argument");
    if ((scale == nullptr) != (shift == nullptr)) return fail("hd_op_conv_gn: scale and shift must come together");
    if (!conv_gemm_can_fuse_gn(B, H, W, Cout, 3, CONV_TAPS)) return fail("hd_op_conv_gn: shape %dx%d -> %d channels is not on the fused path", H, W, Cout);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int Cin = C0 + (x1 ? C1 : 0);
    bf16* q = nullptr;
    float* row = nullptr;
    int* zero = nullptr;
    int* cnt = nullptr;
    float2* part = nullptr;
    auto cleanup = [&]() { cudaFree(q); cudaFree(row); cudaFree(zero); cudaFree(cnt); cudaFree(part); };
#define OP_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) { cleanup(); return fail("%s failed: %s", #expr, cudaGetErrorString(e__)); } \
    } while (0)
    OP_TRY(cudaMalloc(&q, static_cast<size_t>(Cout) * Cin * 9 * 2));
    OP_TRY(cudaMalloc(&cnt, static_cast<size_t>(B) * 4));
    OP_TRY(cudaMalloc(&zero, 4));
    OP_TRY(cudaMalloc(&part, static_cast<size_t>(B) * (H * W / 32) * (Cout / 8) * sizeof(float2)));
    OP_TRY(cudaMemsetAsync(cnt, 0, static_cast<size_t>(B) * 4, s));
    OP_TRY(cudaMemsetAsync(zero, 0, 4, s));
    OP_TRY(prep_conv_weight_run(w, q, Cout, Cin, 3, standardize, 1e-5f, Cout, s));
    ConvGemmDesc d;
    d.src0 = ConvSrc{reinterpret_cast<const bf16*>(x0), C0};
    d.src1 = ConvSrc{reinterpret_cast<const bf16*>(x1), x1 ? C1 : 0};
    d.B = B; d.H = H; d.W = W; d.ksize = 3; d.mode = CONV_TAPS;
    d.weight = q; d.N = Cout; d.out = reinterpret_cast<bf16*>(out);
    d.epi.bias = bias; d.epi.gn_gamma = gamma; d.epi.gn_beta = beta; d.epi.gn_eps = 1e-5f; d.epi.gn_counter = cnt; d.epi.gn_part = part;
    if (scale) {
        OP_TRY(cudaMalloc(&row, 2 * Cout * 4));
        OP_TRY(cudaMemcpyAsync(row, scale, Cout * 4, cudaMemcpyDeviceToDevice, s));
        OP_TRY(cudaMemcpyAsync(row + Cout, shift, Cout * 4, cudaMemcpyDeviceToDevice, s));
        d.epi.film = row; d.epi.film_row = zero; d.epi.film_row_stride = 0; d.epi.film_ld = 2 * Cout; d.epi.film_off = 0;
        d.epi.film_has_scale = 1;
    }
    if (res) { d.epi.res = reinterpret_cast<const bf16*>(res); d.epi.ldr = Cout; }
    ConvGemmLaunch l;
    char msg[256];
    if (conv_gemm_prepare(d, sms, &l, msg, sizeof(msg))) { cleanup(); return fail("%s", msg); }
    OP_TRY(conv_gemm_run(l, s));
    OP_TRY(cudaStreamSynchronize(s));
#undef OP_TRY
    cleanup();
    return 0;
}

int hd_op_groupnorm_silu(const uint16_t* x, uint16_t* y, const float* gamma, const float* beta, const float* scale,
                         const float* shift, const uint16_t* res, int32_t B, int32_t P, int32_t C, void* stream) {
    if (!x || !y || !gamma || !beta) return fail("hd_op_groupnorm_silu: null argument");
    if ((scale == nullptr) != (shift == nullptr)) return fail("hd_op_groupnorm_silu: scale and shift must come together");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GroupNormArgs g;
    g.x = reinterpret_cast<const bf16*>(x); g.y = reinterpret_cast<bf16*>(y);
    g.B = B; g.P = P; g.C = C; g.gamma = gamma; g.beta = beta; g.eps = 1e-5f;
    float* row = nullptr;
    int* zero = nullptr;
    if (scale) {   // pack [scale | shift] into one FiLM row
        CUDA_TRY(cudaMalloc(&row, 2 * C * 4));
        CUDA_TRY(cudaMalloc(&zero, 4));
        CUDA_TRY(cudaMemsetAsync(zero, 0, 4, s));
        CUDA_TRY(cudaMemcpyAsync(row, scale, C * 4, cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(row + C, shift, C * 4, cudaMemcpyDeviceToDevice, s));
        g.film = row; g.film_row = zero; g.film_row_stride = 0; g.film_ld = 2 * C; g.film_off = 0;
    }
    if (res) g.res = reinterpret_cast<const bf16*>(res);
    float2* part = nullptr;
    CUDA_TRY(cudaMalloc(&part, static_cast<size_t>(B) * (P / 32 + 1) * (C / 8) * sizeof(float2)));
    g.part = part;
    cudaError_t e = groupnorm_stats_run(g.x, part, B, P, C, s);
    if (e == cudaSuccess) e = groupnorm_film_silu_run(g, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(row); cudaFree(zero); cudaFree(part);
    if (e != cudaSuccess) return fail("groupnorm launch failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail("groupnorm kernel failed: %s", cudaGetErrorString(e2));
    return 0;
}

int hd_op_channel_layernorm(const uint16_t* x, uint16_t* y, const float* g, const uint16_t* res, int32_t B, int32_t H,
                            int32_t W, int32_t C, int32_t upsample2x, void* stream) {
    if (!x || !y || !g) return fail("hd_op_channel_layernorm: null argument");
    LayerNormArgs l;
    l.x = reinterpret_cast<const bf16*>(x); l.y = reinterpret_cast<bf16*>(y); l.M = B * H * W; l.C = C; l.g = g;
    l.eps = 1e-5f; l.res = reinterpret_cast<const bf16*>(res); l.upsample2x = upsample2x; l.H = H; l.W = W;
    CUDA_TRY(channel_layernorm_run(l, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_op_linear_attention(const uint16_t* qkv, uint16_t* out, int32_t B, int32_t n, void* stream) {
    if (!qkv || !out) return fail("hd_op_linear_attention: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* ctx = nullptr;
    CUDA_TRY(cudaMalloc(&ctx, static_cast<size_t>(B) * 4 * 32 * 32 * 4));
    LinAttnArgs a;
    a.qkv = reinterpret_cast<const bf16*>(qkv); a.out = reinterpret_cast<bf16*>(out); a.ctx = ctx; a.B = B; a.n = n;
    cudaError_t e = linear_attention_run(a, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(ctx);
    if (e != cudaSuccess) return fail("linear attention launch failed: %s", cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail("linear attention kernel failed: %s", cudaGetErrorString(e2));
    return 0;
}

int hd_op_linattn_block(const uint16_t* x, const float* g1, const float* wqkv, const float* wo, const float* bo,
                        const float* g2, uint16_t* y, int32_t B, int32_t n, int32_t C, float* bound_out, void* stream) {
    if (!x || !g1 || !wqkv || !wo || !bo || !g2 || !y) return fail("hd_op_linattn_block: null argument");
    if (C != 64 && C != 128) return fail("hd_op_linattn_block: C must be 64 or 128 (got %d)", C);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    bf16 *wq = nullptr, *mb = nullptr;
    float *rowsum = nullptr, *kshift = nullptr, *ctx = nullptr, *sp = nullptr;
    int* bits = nullptr;
    auto cleanup = [&]() { cudaFree(wq); cudaFree(mb); cudaFree(rowsum); cudaFree(kshift); cudaFree(ctx); cudaFree(sp); cudaFree(bits); };
#define OP_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) { cleanup(); return fail("%s failed: %s", #expr, cudaGetErrorString(e__)); } \
    } while (0)
    OP_TRY(cudaMalloc(&wq, static_cast<size_t>(384) * C * 2));
    OP_TRY(cudaMalloc(&rowsum, 384 * 4));
    OP_TRY(cudaMalloc(&kshift, 128 * 4));
    OP_TRY(cudaMalloc(&bits, 4));
    OP_TRY(cudaMalloc(&ctx, static_cast<size_t>(B) * LA_MAX_PARTS * 128 * 32 * 4));
    OP_TRY(cudaMalloc(&sp, static_cast<size_t>(B) * LA_MAX_PARTS * 128 * 4));
    OP_TRY(cudaMalloc(&mb, static_cast<size_t>(B) * C * 128 * 2));
    OP_TRY(linattn_prep_run(wqkv, g1, wq, rowsum, kshift, bits, C, s));
    OP_TRY(cudaStreamSynchronize(s));
    int hb = 0;
    OP_TRY(cudaMemcpy(&hb, bits, 4, cudaMemcpyDeviceToHost));
    float bound;
    memcpy(&bound, &hb, 4);
    if (bound_out) *bound_out = bound;
    if (bound > LA_MAX_BOUND) { cleanup(); return fail("hd_op_linattn_block: softmax bound %.1f exceeds %.1f (the plan uses the unfused kernels)", bound, LA_MAX_BOUND); }
    LinAttnFusedDesc d;
    d.x = reinterpret_cast<const bf16*>(x); d.y = reinterpret_cast<bf16*>(y); d.B = B; d.n = n; d.C = C;
    d.wqkv = wq; d.rowsum = rowsum; d.kshift = kshift; d.wo = wo; d.bo = bo; d.g2 = g2;
    d.ctx_part = ctx; d.s_part = sp; d.mb = mb; d.max_parts = LA_MAX_PARTS; d.eps = 1e-5f;
    LinAttnFusedLaunch l;
    char msg[256];
    if (linattn_fused_prepare(d, sms, &l, msg, sizeof(msg))) { cleanup(); return fail("%s", msg); }
    OP_TRY(linattn_fused_run(l, s));
    OP_TRY(cudaStreamSynchronize(s));
#undef OP_TRY
    cleanup();
    return 0;
}

int hd_op_full_attention(const uint16_t* qkv, uint16_t* out, int32_t B, int32_t n, void* stream) {
    if (!qkv || !out) return fail("hd_op_full_attention: null argument");
    FullAttnArgs a;
    a.qkv = reinterpret_cast<const bf16*>(qkv); a.out = reinterpret_cast<bf16*>(out); a.B = B; a.n = n;
    CUDA_TRY(full_attention_run(a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_op_stem_conv(const float* x0, const float* x1, const float* w, const float* bias, uint16_t* y, int32_t B,
                    int32_t Cout, int32_t Cin, int32_t ksize, void* stream) {
    if (!x0 || !w || !bias || !y) return fail("hd_op_stem_conv: null argument");
    StemConvArgs a;
    a.x0 = x0; a.x1 = x1; a.w = w; a.bias = bias; a.y = reinterpret_cast<bf16*>(y);
    a.B = B; a.H = 64; a.W = 64; a.Cout = Cout; a.Cin = Cin; a.ksize = ksize;
    CUDA_TRY(stem_conv_run(a, static_cast<cudaStream_t>(stream)));
    return 0;
}

int hd_op_philox_normal(float* out, int64_t n, uint64_t seed, uint64_t tile_offset, void* stream) {
    if (!out) return fail("hd_op_philox_normal: null argument");
    CUDA_TRY(philox_normal_run(out, n, seed, tile_offset, 4096, 0, static_cast<cudaStream_t>(stream)));
    return 0;
}

// ------------------------------------------------------------------------------------------------- debug / stats
namespace {
__global__ void nhwc_bf16_to_nchw_f32(const bf16* x, float* y, int B, int H, int W, int C) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long total = static_cast<long long>(B) * H * W * C;
    if (i >= total) return;
    const int c = static_cast<int>(i % C);
    const long long p = i / C;
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const int b = static_cast<int>(p / (static_cast<long long>(W) * H));
    y[((static_cast<long long>(b) * C + c) * H + h) * W + w] = __bfloat162float(x[i]);
}
}  // namespace

int hd_debug_read(hd_plan* P, int32_t B, const char* name, float* out, int64_t* numel, int32_t* shape4, void* stream) {
    if (!P || !name) return fail("hd_debug_read: null argument");
    auto it = P->execs.find(B);
    if (it == P->execs.end()) return fail("hd_debug_read: no executor for batch %d (run hd_eps_forward first)", B);
    auto d = it->second->dbg.find(name);
    if (d == it->second->dbg.end()) return fail("hd_debug_read: unknown activation '%s' (is cfg.debug_keep set?)", name);
    const DebugEntry& e = d->second;
    const long long total = static_cast<long long>(B) * e.H * e.W * e.C;
    if (numel) *numel = total;
    if (shape4) { shape4[0] = B; shape4[1] = e.C; shape4[2] = e.H; shape4[3] = e.W; }
    if (out) {
        const int grid = static_cast<int>((total + 255) / 256);
        nhwc_bf16_to_nchw_f32<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(e.p, out, B, e.H, e.W, e.C);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int hd_debug_names(hd_plan* P, int32_t B, char* buf, int64_t buflen) {
    if (!P || !buf || buflen <= 0) return fail("hd_debug_names: null argument");
    auto it = P->execs.find(B);
    if (it == P->execs.end()) return fail("hd_debug_names: no executor for batch %d", B);
    std::string all;
    for (const std::string& n : it->second->dbg_order) { all += n; all += '\n'; }
    if (static_cast<int64_t>(all.size()) + 1 > buflen) return fail("hd_debug_names: buffer too small (%zu needed)", all.size() + 1);
    memcpy(buf, all.c_str(), all.size() + 1);
    return 0;
}

int hd_plan_launches_per_step(hd_plan* P, int32_t B, int32_t* eps_launches, int32_t* step_launches) {
    if (!P) return fail("null plan");
    auto it = P->execs.find(B);
    if (it == P->execs.end()) return fail("no executor for batch %d yet", B);
    // linear attention is two kernels behind one op
    auto count = [](const std::vector<Op>& ops) {
        int n = 0;
        for (const Op& o : ops) n += ends_with(o.tag, ".linattn") ? 2 : (ends_with(o.tag, ".linattn_fused") ? 3 : 1);
        return n;
    };
    if (eps_launches) *eps_launches = count(it->second->ops_rows);
    if (step_launches) *step_launches = count(it->second->ops_table);
    return 0;
}

int hd_plan_profile_step(hd_plan* P, int32_t B, int32_t reps, char* buf, int64_t buflen, void* stream) {
    if (!P || !buf || buflen <= 0 || reps <= 0) return fail("hd_plan_profile_step: bad argument");
    if (ensure_device(P)) return 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Exec* ex = nullptr;
    if (get_exec(P, B, s, &ex)) return 1;
    if (set_ctl(P->ctl, P->T - 1, 0, nullptr, 1, 0, s)) return 1;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    std::string out = "[";
    bool first = true;
    for (const Op& op : ex->ops_table) {
        if (strcmp(op.kernel, "step_advance") == 0) continue;   // would walk the step counter off the table
        cudaError_t e = op.fn(s);   // warm-up
        if (e == cudaSuccess) e = cudaEventRecord(e0, s);
        for (int r = 0; r < reps && e == cudaSuccess; ++r) e = op.fn(s);
        if (e == cudaSuccess) e = cudaEventRecord(e1, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return fail("profiling '%s' failed: %s", op.tag.c_str(), cudaGetErrorString(e));
        }
        char line[512];
        snprintf(line, sizeof(line), "%s{\"tag\":\"%s\",\"kernel\":\"%s\",\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
                 first ? "" : ",", op.tag.c_str(), op.kernel, ms / reps, op.flops, op.bytes);
        out += line;
        first = false;
    }
    out += "]";
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (static_cast<int64_t>(out.size()) + 1 > buflen) return fail("hd_plan_profile_step: buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

int64_t hd_plan_device_bytes(hd_plan* P) {
    if (!P) return 0;
    size_t total = P->weight_bytes;
    for (auto& kv : P->w) {
        auto q = P->wq.find(kv.first);
        if (q != P->wq.end()) total += kv.second.numel * 2;
    }
    total += static_cast<size_t>(P->T) * (P->film_ld + 9) * 4;
    for (auto& kv : P->execs) total += kv.second->arena_bytes + kv.second->extra_bytes;
    return static_cast<int64_t>(total);
}

}  // extern "C"
