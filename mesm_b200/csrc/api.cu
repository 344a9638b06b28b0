// C ABI of mesm_b200 (include/mesm_b200.h): context, weight ingestion/packing and the forward orchestration.
//
// The forward is a stream-ordered sequence of fused-linear, attention and row-wise kernels over caller-owned
// workspace.  Pairs are processed in chunks of whole video groups so that a chunk's activations stay L2-resident
// between the producing and the consuming kernel; the text side (tiny) is computed once for the whole batch because
// the reference's negative branch and attn_mask quirk reach across pairs.
#include "ctx.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace mesm;

namespace {


// ---- packing kernels ---------------------------------------------------------------------------------------------
__global__ void transpose_pack_kernel(const float* __restrict__ W, int row0, int nrows, int K, const float* __restrict__ gamma,
                                      float* __restrict__ Wt, int ldw, int Kp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Kp * ldw) return;
    const int k = (int)(idx / ldw), n = (int)(idx % ldw);
    float v = 0.f;
    if (k < K && n < nrows) v = W[(long long)(row0 + n) * K + k] * (gamma ? gamma[k] : 1.f);
    Wt[idx] = v;
}
__global__ void colsum_kernel(const float* __restrict__ Wt, int Kp, int ldw, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int k = 0; k < Kp; ++k) s += Wt[(long long)k * ldw + n];
    out[n] = s;
}
__global__ void fold_bias_kernel(const float* __restrict__ W, int row0, int N, int K, const float* __restrict__ beta,
                                 const float* __restrict__ b, float* __restrict__ out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(W[(long long)(row0 + n) * K + k], beta[k], s);
    s = warp_sum(s);
    if (lane == 0) out[n] = s + (b ? b[row0 + n] : 0.f);
}
__global__ void vec_add_kernel(const float* a, const float* b, float* o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}

}  // namespace
namespace mesm {
cudaError_t launch_colsum(const float* Wt, int Kp, int ldw, int N, float* out, cudaStream_t s) {
    colsum_kernel<<<(N + 127) / 128, 128, 0, s>>>(Wt, Kp, ldw, N, out);
    g_stats.launches++;
    return cudaGetLastError();
}
cudaError_t launch_transpose_pack(const float* W, int row0, int nrows, int K, float* Wt, int ldw, int Kp, cudaStream_t s) {
    const long long tot = (long long)Kp * ldw;
    transpose_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W, row0, nrows, K, nullptr, Wt, ldw, Kp);
    g_stats.launches++;
    return cudaGetLastError();
}
}  // namespace mesm
namespace {

// SS-MESM reconstructor (model/model.py:467-488) evaluated on the unprojected clips needs, per head h, qk_h = Wk_h^T q_h (32 -> 256) and
// out_h = Wv_h pooled_h (256 -> 32).  As dense block-structured matrices both are ordinary linears over all heads at once:
//   Wqk[(h, c), k] = (k / 32 == h) ? Wk[k, c] : 0      [8 * 256, 256]      qk[b, h, :] = Wqk q_b
//   Wvo[n, (h, c)] = (n / 32 == h) ? Wv[n, c] : 0      [256, 8 * 256]      out_b      = Wvo pooled_b + bv
// (7/8 of the MACs multiply zeros - 4 GFLOP per launch, nothing - but the launches move from the fp32 SIMT kernel to tcgen05).
__global__ void recon_block_weights_kernel(const float* __restrict__ in_w, float* __restrict__ Wqk, float* __restrict__ Wvo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;            // 0 .. 8 * 256 * 256
    if (i >= NH * D * D) return;
    {   // Wqk row r = h * 256 + c, column k
        const int r = i / D, k = i - r * D, h = r / D, c = r - h * D;
        Wqk[i] = (k / HD == h) ? in_w[(long long)(D + k) * D + c] : 0.f;
    }
    {   // Wvo row n, column q = h * 256 + c
        const int n = i / (NH * D), q = i - n * (NH * D), h = q / D, c = q - h * D;
        Wvo[i] = (n / HD == h) ? in_w[(long long)(2 * D + n) * D + c] : 0.f;
    }
}

struct Packer {
    mesm_ctx* ctx;
    cudaStream_t s;
    std::string missing;
    bool ok = true;
    cudaError_t cerr = cudaSuccess;

    float* alloc(size_t nfloats) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(nfloats, 4) * sizeof(float));
        if (e != cudaSuccess) { ok = false; cerr = e; return nullptr; }
        ctx->owned.push_back(p);
        return (float*)p;
    }
    const Tensor* get(const std::string& key, std::initializer_list<int64_t> shape) {
        auto it = ctx->w.find(key);
        if (it == ctx->w.end()) { missing += key + " "; return nullptr; }
        if (std::vector<int64_t>(shape) != it->second.shape) { missing += key + "(shape) "; return nullptr; }
        return &it->second;
    }
    const float* vec(const std::string& key, int64_t n) { const Tensor* t = get(key, {n}); return t ? t->p : nullptr; }
    Norm norm(const std::string& p, int64_t n = D) { Norm r; r.g = vec(p + ".weight", n); r.b = vec(p + ".bias", n); return r; }

    // rows [row0,row0+nrows) of W[Ntot,K] (+ bias rows) -> packed linear
    PL pack(const std::string& wkey, const std::string& bkey, int64_t Ntot, int64_t K, int row0, int nrows,
            const float* gamma = nullptr, const float* beta = nullptr) {
        PL r;
        const Tensor* W = get(wkey, {Ntot, K});
        const Tensor* b = bkey.empty() ? nullptr : get(bkey, {Ntot});
        if (!W || (!bkey.empty() && !b)) return r;
        const int Kp = (int)((K + 15) / 16 * 16), ldw = (nrows + 3) / 4 * 4;
        float* Wt = alloc((size_t)Kp * ldw);
        if (!Wt) return r;
        const long long tot = (long long)Kp * ldw;
        transpose_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W->p, row0, nrows, (int)K, gamma, Wt, ldw, Kp);
        r.Wt = Wt; r.ldw = ldw; r.K = (int)K; r.N = nrows;
        if (nrows >= 64) {
            void* wp = nullptr;
            cudaError_t e = cudaMalloc(&wp, tc_packed_bytes(nrows, (int)K));
            if (e != cudaSuccess) { ok = false; cerr = e; return r; }
            ctx->owned.push_back(wp);
            launch_pack_tc(W->p, row0, nrows, (int)K, gamma, wp, s);
            r.Wp = wp;
        }
        if (nrows % 256 == 0 && K % 8 == 0 && !gamma) {          // TMA-fed kernel: weights as plain bf16 hi/lo planes + tensor maps
            void* wt = tma_pack_weights(W->p, row0, nrows, (int)K, nullptr, false, s);
            if (wt) { ctx->owned_tma.push_back(wt); r.Wtm = wt; }
        }
        if (gamma) {
            float* cs = alloc(nrows);
            float* cb = alloc(nrows);
            if (!cs || !cb) return r;
            colsum_kernel<<<(nrows + 127) / 128, 128, 0, s>>>(Wt, Kp, ldw, nrows, cs);
            fold_bias_kernel<<<(nrows + 7) / 8, 256, 0, s>>>(W->p, row0, nrows, (int)K, beta, b ? b->p : nullptr, cb);
            r.colsum = cs; r.bias = cb;
        } else {
            r.bias = b ? b->p + row0 : nullptr;
        }
        return r;
    }
    PL lin(const std::string& p, int64_t N, int64_t K) { return pack(p + ".weight", p + ".bias", N, K, 0, (int)N); }
    // a weight matrix built on the device (W [N, K] fp32 row-major, bias [N] or null): fp32 transposed copy + tcgen05 image
    PL pack_dev(const float* W, const float* bias, int N, int K) {
        PL r;
        const int Kp = (K + 15) / 16 * 16, ldw = (N + 3) / 4 * 4;
        float* Wt = alloc((size_t)Kp * ldw);
        if (!Wt) return r;
        const long long tot = (long long)Kp * ldw;
        transpose_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W, 0, N, K, nullptr, Wt, ldw, Kp);
        r.Wt = Wt; r.ldw = ldw; r.K = K; r.N = N; r.bias = bias;
        void* wp = nullptr;
        cudaError_t e = cudaMalloc(&wp, tc_packed_bytes(N, K));
        if (e != cudaSuccess) { ok = false; cerr = e; return r; }
        ctx->owned.push_back(wp);
        launch_pack_tc(W, 0, N, K, nullptr, wp, s);
        r.Wp = wp;
        return r;
    }
    const float* bias_sum(const float* a, const float* b, int n) {
        if (!a || !b) return nullptr;
        float* o = alloc(n);
        if (o) vec_add_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, o, n);
        return o;
    }
    AttnFfn attn_ffn(const std::string& p, bool recon) {
        AttnFfn L;
        const std::string iw = p + "self_attn.in_proj_weight", ib = p + "self_attn.in_proj_bias";
        L.q = pack(iw, ib, 3 * D, D, 0, D);
        L.kv = pack(iw, ib, 3 * D, D, D, 2 * D);
        L.qk = pack(iw, ib, 3 * D, D, 0, 2 * D);
        L.v = pack(iw, ib, 3 * D, D, 2 * D, D);
        L.out = lin(p + "self_attn.out_proj", D, D);
        L.l1 = lin(p + "linear1", FF, D);
        L.l2 = lin(p + "linear2", D, FF);
        L.n1 = norm(p + "norm1"); L.n2 = norm(p + "norm2");
        L.prelu = vec(p + "activation.weight", 1);
        const Tensor* W = get(iw, {3 * D, D});
        const Tensor* b = get(ib, {3 * D});
        L.in_w = W ? W->p : nullptr; L.in_b = b ? b->p : nullptr;
        if (recon) L.vT = L.v.Wt;
        if (recon && W && b && D == 256) {
            float* Wqk = alloc((size_t)NH * D * D);
            float* Wvo = alloc((size_t)NH * D * D);
            if (Wqk && Wvo) {
                recon_block_weights_kernel<<<(NH * D * D + 255) / 256, 256, 0, s>>>(W->p, Wqk, Wvo);
                L.blk_qk = pack_dev(Wqk, nullptr, NH * D, D);
                L.blk_v = pack_dev(Wvo, b->p + 2 * D, D, NH * D);
            }
        }
        if (!recon && D == 256 && FF == 1024) {             // fused FFN kernel: weight slots in consumption order
            const Tensor* W1 = get(p + "linear1.weight", {FF, D});
            const Tensor* W2 = get(p + "linear2.weight", {D, FF});
            void *w1 = nullptr, *w2 = nullptr;
            if (W1 && W2 && cudaMalloc(&w1, ffn_packed_bytes()) == cudaSuccess && cudaMalloc(&w2, ffn_packed_bytes()) == cudaSuccess) {
                ctx->owned.push_back(w1); ctx->owned.push_back(w2);
                launch_pack_ffn(W1->p, W2->p, w1, w2, s);
                L.ffn_w1 = w1; L.ffn_w2 = w2;
                L.ffn_maps = ffn_make_maps(w1, w2);
                if (L.ffn_maps) ctx->owned_host.push_back(const_cast<void*>(L.ffn_maps));
            }
        }
        return L;
    }
};


}  // namespace

namespace mesm {

thread_local std::string g_create_error;
int fail(mesm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code ? code : 1;
}

// FFN block of a layer: out = LN2(res + W2 PReLU(W1 x + b1) + b2).  One fused tcgen05 kernel when the shape allows
// (MESM_FFN_FUSED=0 keeps the two-GEMM path, which also serves small row counts).
cudaError_t ffn_block(const AttnFfn& L, int R, const float* x, const float* res, float* H, float* out, int ldo, RowMap omap,
                      cudaStream_t s, Planes xp = Planes(), Planes outp = Planes(), const float* ln1_stats = nullptr, bool res_ln1 = false) {
    static int fused = -1;
    if (fused < 0) { const char* e = getenv("MESM_FFN_FUSED"); fused = (e && e[0] == '0') ? 0 : 1; }
    FfnArgs a;
    a.X = x; a.ldx = D; a.R = res; a.ldr = D; a.out = out; a.ldo = ldo; a.omap = omap; a.M = R;
    a.W1f = L.ffn_w1; a.W2f = L.ffn_w2; a.maps = L.ffn_maps; a.b1 = L.l1.bias; a.b2 = L.l2.bias; a.ln_g = L.n2.g; a.ln_b = L.n2.b; a.prelu = L.prelu;
    a.x_hi = xp.hi; a.x_lo = xp.lo; a.out_hi = outp.hi; a.out_lo = outp.lo;
    if (ln1_stats) { a.ln1_stats = ln1_stats; a.ln1_g = L.n1.g; a.ln1_b = L.n1.b; a.res_ln1 = res_ln1 ? 1 : 0; }   // x (and res) are PRE-LayerNorm-1 rows
    if (fused && ffn_fused_eligible(a)) {
        ProfScope _ps("ffn_fused", s, 4.0 * R * (double)D * FF, R);
        return launch_ffn_fused(a, s);
    }
    if (xp || outp || ln1_stats) return cudaErrorInvalidValue;  // the plane flow is only set up where the fused kernel runs
    MESM_CHECK(Lin(R, L.l1, x, D, H, FF).act(ACT_PRELU, L.prelu).run(s));
    MESM_CHECK(Lin(R, L.l2, H, FF, out, ldo).omap(omap).res(res, D).ln(L.n2).run(s));
    return cudaSuccess;
}

// A/B switches of the plane flow (developer).  MESM_FFN_X=planes feeds the fused FFN's X operand as pre-split planes through the TMA
// engine; the default keeps it fp32 and converts it in the kernel, which measured 1 ms per step FASTER (26.1 -> 25.0 ms over the 12
// launches): a tile's start-up is bound by the HBM burst of a whole wave asking for its X rows at once, not by the conversion.
// MESM_FFN_OUTP=0 stops the FFN from storing its result as planes (the next layer then falls back to fp32 operands).
static bool ffn_x_f32() { static int v = -1; if (v < 0) { const char* e = getenv("MESM_FFN_X"); v = (e && e[0] == 'p') ? 0 : 1; } return v == 1; }
// MESM_FFN_LN1=0: the output projection normalises (LN1) and stores the result for the FFN, as in round 1; default: it stores only the
// pre-LN rows and their (mean, rstd), and the fused FFN applies LayerNorm-1 while it converts its X operand
static bool ffn_ln1_fused() { static int v = -1; if (v < 0) { const char* e = getenv("MESM_FFN_LN1"); v = (e && e[0] == '0') ? 0 : 1; } return v == 1; }
static bool ffn_outp_off() { static int v = -1; if (v < 0) { const char* e = getenv("MESM_FFN_OUTP"); v = (e && e[0] == '0') ? 1 : 0; } return v == 1; }

// the plane flow needs the fused FFN (M > 128 rows), its weight images and the TMA weight planes of the layer
static bool planes_ok(const AttnFfn& L, int R, const PlaneIO* pio) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MESM_PLANES"); on = (e && e[0] == '0') ? 0 : 1; }
    return on && pio && pio->ao && pio->y1 && R > 128 && L.ffn_maps && L.out.Wtm && L.q.Wtm && L.qk.Wtm && L.v.Wtm;
}

// T2V layer (model/transformer.py:508-540).  txt rows: [Bc*Lk] through tmap; vid rows [Bc*Lq] contiguous.
cudaError_t t2v_layer(const AttnFfn& L, const float* txt, RowMap tmap, const float* pos_txt, int Lk, const float* vid,
                      const float* pos_vid, int Lq, int Bc, int b0, int Btot, const uint8_t* q_pad, const uint8_t* k_pad,
                      const T2VBuffers& t, float* out, int ldo, RowMap omap, cudaStream_t s, bool reuse_q, const int* cu,
                      int Rv_packed, int q_pad_ld, const float* posW, const int* t_pos, PlaneIO* pio) {
    const int Rt = Bc * Lk, Rv = cu ? Rv_packed : Bc * Lq;
    const bool pl = planes_ok(L, Rv, pio) && ldo == D;
    if (pio) { pio->wrote_out = pl && pio->out; pio->wrote_fp32 = true; }
    if (!vid && !(pl && pio->in && posW && ffn_x_f32() && ffn_ln1_fused())) return cudaErrorInvalidValue;     // planes-only input needs the full plane path
    if (pos_txt) {
        PL kw = L.kv; kw.N = D; kw.Wtm = nullptr;
        MESM_CHECK(Lin(Rt, kw, txt, D, t.KV, 2 * D).amap(tmap).apos(pos_txt).run(s));
        PL vw = L.v;
        MESM_CHECK(Lin(Rt, vw, txt, D, t.KV + D, 2 * D).amap(tmap).run(s));
    } else {
        MESM_CHECK(Lin(Rt, L.kv, txt, D, t.KV, 2 * D).amap(tmap).run(s));
    }
    if (!reuse_q) {                                                                     // Q depends on the clips only
        Lin q(Rv, L.q, vid, D, t.Q, D);
        if (posW) q.res(posW, D, table_map(t_pos)); else q.apos(pos_vid);             // (x + pos) Wq = x Wq + (pos Wq)[table]
        if (pl && pio->in && posW) q.aplanes(pio->in.hi, pio->in.lo, D);              // input already pre-split: TMA-fed kernel
        MESM_CHECK(q.run(s));
    }
    MhaRowsArgs a;
    a.q = t.Q; a.ldq = D; a.k = t.KV; a.ldk = 2 * D; a.v = t.KV + D; a.ldv = 2 * D;
    a.k_pad = k_pad; a.q_pad = q_pad; a.out = t.AO; a.ldo = D; a.B = Bc; a.Lq = Lq; a.Lk = Lk; a.b0 = b0; a.Btot = Btot;
    a.q_scale = kScale32;
    a.q_cu = cu; a.q_enc = 0; a.q_pad_ld = q_pad_ld;         // packed clips: pair b's queries are rows cu[b]-cu[0] ...
    if (pl) { a.out = nullptr; a.out_hi = pio->ao.hi; a.out_lo = pio->ao.lo; }      // attention output straight into operand planes
    MESM_CHECK(launch_mha_rows(a, s));
    if (pl) {
        // out-proj (+ residual, pre-LN copy, LN1) on the TMA-fed kernel; LN1's result only exists as the FFN's operand planes
        const bool xf = ffn_x_f32();
        if (ffn_outp_off()) pio->wrote_out = false;
        const Planes outp = pio->wrote_out ? pio->out : Planes();
        if (xf && ffn_ln1_fused()) {
            // out-proj (+ residual) stores src2 = vid + attn (pre-LN) and its row statistics; LN1 happens inside the FFN's converter:
            // LN1's output never exists in HBM (1 KB/row written + read back per layer in round 1)
            Lin op_(Rv, L.out, nullptr, D, nullptr, D);
            op_.aplanes(pio->ao.hi, pio->ao.lo, D).pre_ln(t.X1).ln_stats(t.H);
            if (vid) op_.res(vid, D); else op_.res_planes(pio->in.hi, pio->in.lo);
            MESM_CHECK(op_.run(s));
            const bool po = pio->planes_only && outp;                       // intermediate layer: the next consumer reads planes only
            pio->wrote_fp32 = !po;
            MESM_CHECK(ffn_block(L, Rv, t.X1, t.X1, nullptr, po ? nullptr : out, ldo, omap, s, Planes(), outp, t.H, false));
            return cudaSuccess;
        }
        MESM_CHECK(Lin(Rv, L.out, nullptr, D, xf ? t.Y1 : nullptr, D).aplanes(pio->ao.hi, pio->ao.lo, D).res(vid, D).pre_ln(t.X1).ln(L.n1)
                       .oplanes(xf ? nullptr : pio->y1.hi, xf ? nullptr : pio->y1.lo, D).run(s));
        MESM_CHECK(ffn_block(L, Rv, xf ? t.Y1 : nullptr, t.X1, t.H, out, ldo, omap, s, xf ? Planes() : pio->y1, outp));
        return cudaSuccess;
    }
    MESM_CHECK(Lin(Rv, L.out, t.AO, D, t.Y1, D).res(vid, D).pre_ln(t.X1).ln(L.n1).run(s));
    MESM_CHECK(ffn_block(L, Rv, t.Y1, t.X1, t.H, out, ldo, omap, s));
    return cudaSuccess;
}

// Encoder layer (model/transformer.py:637-650) on the [Bc, L1, 256] buffer (L1 = Lv + 1, global token first).
cudaError_t enc_layer(const AttnFfn& L, const float* src, const float* pos, const uint8_t* pad, int L1, int Bc,
                      const EncBuffers& t, float* out, cudaStream_t s, const int* cu, int R_packed, const float* posW,
                      const int* t_pos, PlaneIO* pio) {
    const int R = cu ? R_packed : Bc * L1;
    const bool pl = planes_ok(L, R, pio) && pio->in && posW;
    if (pio) { pio->wrote_out = pl && pio->out; pio->wrote_fp32 = true; }
    if (!src && !(pl && ffn_x_f32() && ffn_ln1_fused())) return cudaErrorInvalidValue;
    {
        Lin qk(R, L.qk, src, D, t.QKV, 3 * D);
        if (posW) qk.res(posW, 2 * D, table_map(t_pos)); else qk.apos(pos);
        if (pl) qk.aplanes(pio->in.hi, pio->in.lo, D);
        MESM_CHECK(qk.run(s));
    }
    {
        Lin v(R, L.v, src, D, t.QKV + 2 * D, 3 * D);
        if (pl) v.aplanes(pio->in.hi, pio->in.lo, D);
        MESM_CHECK(v.run(s));
    }
    MhaRowsArgs a;
    a.q = t.QKV; a.ldq = 3 * D; a.k = t.QKV + D; a.ldk = 3 * D; a.v = t.QKV + 2 * D; a.ldv = 3 * D;
    a.k_pad = pad; a.q_pad = nullptr; a.out = t.AO; a.ldo = D; a.B = Bc; a.Lq = L1; a.Lk = L1; a.b0 = 0; a.Btot = Bc;
    a.q_scale = kScale32;
    a.q_cu = cu; a.q_enc = 1; a.k_cu = cu; a.k_enc = 1;      // packed encoder rows (global token + this pair's clips)
    a.split_ws = t.split; a.split_rows = R;                  // more keys than one 224-key tile (max_video_l = 600): key-split passes
    if (pl) { a.out = nullptr; a.out_hi = pio->ao.hi; a.out_lo = pio->ao.lo; }
    MESM_CHECK(launch_mha_rows(a, s));
    if (pl) {
        // LN1's result is both the FFN's operand (planes) and its residual (fp32)
        const bool xf = ffn_x_f32();
        if (ffn_outp_off()) pio->wrote_out = false;
        const Planes outp = pio->wrote_out ? pio->out : Planes();
        if (xf && ffn_ln1_fused()) {
            // pre-LN rows + statistics only; the FFN normalises them for its X operand AND for its residual (LN1's result is both)
            Lin op_(R, L.out, nullptr, D, nullptr, D);
            op_.aplanes(pio->ao.hi, pio->ao.lo, D).pre_ln(t.Y1).ln_stats(t.H);
            if (src) op_.res(src, D); else op_.res_planes(pio->in.hi, pio->in.lo);
            MESM_CHECK(op_.run(s));
            const bool po = pio->planes_only && outp;
            pio->wrote_fp32 = !po;
            MESM_CHECK(ffn_block(L, R, t.Y1, t.Y1, nullptr, po ? nullptr : out, D, identity_map(), s, Planes(), outp, t.H, true));
            return cudaSuccess;
        }
        MESM_CHECK(Lin(R, L.out, nullptr, D, t.Y1, D).aplanes(pio->ao.hi, pio->ao.lo, D).res(src, D).ln(L.n1)
                       .oplanes(xf ? nullptr : pio->y1.hi, xf ? nullptr : pio->y1.lo, D).run(s));
        MESM_CHECK(ffn_block(L, R, xf ? t.Y1 : nullptr, t.Y1, t.H, out, D, identity_map(), s, xf ? Planes() : pio->y1, outp));
        return cudaSuccess;
    }
    MESM_CHECK(Lin(R, L.out, t.AO, D, t.Y1, D).res(src, D).ln(L.n1).run(s));
    MESM_CHECK(ffn_block(L, R, t.Y1, t.Y1, t.H, out, D, identity_map(), s));
    return cudaSuccess;
}

// DAB-DETR decoder + heads (model/transformer.py:333-420, model/model.py:246-252) on a chunk.
size_t dec_alloc(Arena& ar, DecBuffers& d, int Bc, int nq, int L1, int nl) {
    const size_t R = (size_t)Bc * nq, Re = (size_t)Bc * L1;
    d.tgtA = ar.get<float>(R * D); d.tgtB = ar.get<float>(R * D); d.ref = ar.get<float>(R * 2);
    d.refs = ar.get<float>((size_t)nl * R * 2); d.sine = ar.get<float>(R * D); d.sine_s = ar.get<float>(R * D);
    d.h1 = ar.get<float>(R * D); d.h2 = ar.get<float>(R * D); d.qpos = ar.get<float>(R * D); d.ptrans = ar.get<float>(R * D);
    d.anc = ar.get<float>(R); d.qsa = ar.get<float>(R * D); d.ksa = ar.get<float>(R * D); d.vsa = ar.get<float>(R * D);
    d.ao = ar.get<float>(R * D); d.t1 = ar.get<float>(R * D); d.qca = ar.get<float>(R * D); d.sinep = ar.get<float>(R * D);
    d.t2 = ar.get<float>(R * D); d.hff = ar.get<float>(R * FF); d.d2 = ar.get<float>(R * 2);
    d.hs = ar.get<float>((size_t)nl * R * D); d.Kc = ar.get<float>(Re * D); d.Kp = ar.get<float>(Re * D);
    d.Vd = ar.get<float>(Re * D); d.tmpref = ar.get<float>(R * 2);
    return ar.off;
}

cudaError_t run_decoder(const mesm_ctx* c, const float* qembed, const float* E, const float* posE, const uint8_t* padV, int Lv, int Bc,
                        const DecBuffers& d, float* logits_out, float* spans_out, float* aux_logits, float* aux_spans,
                        long long aux_layer_stride, float* hs_out, long long hs_layer_stride, float* refs_out,
                        long long refs_layer_stride, cudaStream_t s, const int* cu, int Re_packed, const float* const* posWkp,
                        const int* t_posE, Planes Ep) {
    const int nq = c->cfg.num_queries, nl = c->cfg.dec_layers, L1 = Lv + 1;
    const int R = Bc * nq, Re = cu ? Re_packed : Bc * L1;
    MESM_CHECK(launch_fill(d.tgtA, (long long)R * D, 0.f, s));                         // tgt = 0 (transformer.py:201)
    MESM_CHECK(launch_dec_init_ref(qembed, Bc, nq, d.refs, s));                       // refs[0] = sigmoid(query_embed)
    float* tgt = d.tgtA;
    float* tgt_next = d.tgtB;
    const float* ref = d.refs;
    for (int lid = 0; lid < nl; ++lid) {
        const DecLayer& L = c->dec[lid];
        MESM_CHECK(Lin(R, c->ra0, tgt, D, d.h1, D).act(ACT_RELU).run(s));
        MESM_CHECK(Lin(R, c->ra1, d.h1, D, d.anc, 1).run(s));
        const float* ptrans = nullptr;
        if (lid > 0) {
            MESM_CHECK(Lin(R, c->qs0, tgt, D, d.h1, D).act(ACT_RELU).run(s));
            MESM_CHECK(Lin(R, c->qs1, d.h1, D, d.ptrans, D).run(s));
            ptrans = d.ptrans;
        }
        MESM_CHECK(launch_dec_sine(ref, R, ptrans, d.anc, d.sine, d.sine_s, s));
        MESM_CHECK(Lin(R, c->rph0, d.sine, D, d.h1, D).act(ACT_RELU).run(s));
        MESM_CHECK(Lin(R, c->rph1, d.h1, D, d.qpos, D).run(s));
        // self-attention over the queries
        MESM_CHECK(Lin(R, L.sa_qc, tgt, D, d.qsa, D).second(d.qpos, D, L.sa_qp).bias(L.sa_q_bias).run(s));
        MESM_CHECK(Lin(R, L.sa_kc, tgt, D, d.ksa, D).second(d.qpos, D, L.sa_kp).bias(L.sa_k_bias).run(s));
        MESM_CHECK(Lin(R, L.sa_v, tgt, D, d.vsa, D).run(s));
        MhaSmallArgs a;
        a.q = d.qsa; a.ldq = D; a.q2 = nullptr; a.ldq2 = 0; a.k = d.ksa; a.ldk = D; a.k2 = nullptr; a.ldk2 = 0;
        a.v = d.vsa; a.ldv = D; a.k_pad = nullptr; a.out = d.ao; a.ldo = D; a.attn_w = nullptr;
        a.B = Bc; a.L = nq; a.S = nq; a.nheads = NH; a.hq = HD; a.hv = HD; a.scale = kScale32;
        a.q_bs = nq; a.q_is = 1; a.k_bs = nq; a.k_is = 1; a.k_off = 0;
        MESM_CHECK(launch_mha_small(a, s));
        MESM_CHECK(Lin(R, L.sa_out, d.ao, D, d.t1, D).res(tgt, D).ln(L.n1).run(s));
        // cross-attention into the encoder memory
        if (lid == 0) {
            MESM_CHECK(Lin(R, L.ca_qc, d.t1, D, d.qca, D).second(d.qpos, D, L.ca_qp).bias(L.ca_q_bias0).run(s));
            if (posWkp) {                                                                  // k_content + k_pos
                Lin kc(Re, L.ca_kc, E, D, d.Kc, D);
                kc.res(posWkp[0], D, table_map(t_posE));
                if (Ep && L.ca_kc.Wtm) kc.aplanes(Ep.hi, Ep.lo, D);
                MESM_CHECK(kc.run(s));
            }
            else MESM_CHECK(Lin(Re, L.ca_kc, E, D, d.Kc, D).second(posE, D, L.ca_kp).bias(L.ca_k_bias0).run(s));
        } else {
            MESM_CHECK(Lin(R, L.ca_qc, d.t1, D, d.qca, D).run(s));
            Lin kc(Re, L.ca_kc, E, D, d.Kc, D);
            if (Ep && L.ca_kc.Wtm) kc.aplanes(Ep.hi, Ep.lo, D);
            MESM_CHECK(kc.run(s));
        }
        if (!posWkp) MESM_CHECK(Lin(Re, L.ca_kp, posE, D, d.Kp, D).run(s));          // else: rows of the position table product
        {
            Lin cv(Re, L.ca_v, E, D, d.Vd, D);
            if (Ep && L.ca_v.Wtm) cv.aplanes(Ep.hi, Ep.lo, D);
            MESM_CHECK(cv.run(s));
        }
        MESM_CHECK(Lin(R, L.ca_sine, d.sine_s, D, d.sinep, D).run(s));
        a.q = d.qca; a.q2 = d.sinep; a.ldq2 = D; a.k = d.Kc; a.k2 = posWkp ? posWkp[lid] : d.Kp; a.ldk2 = D; a.v = d.Vd; a.k_pad = padV;
        a.k2_table = posWkp ? t_posE : nullptr;
        a.S = Lv; a.scale = kScale64; a.k_bs = L1; a.k_is = 1; a.k_off = 1;
        a.k_cu = cu; a.k_enc = 1;                            // packed memory: pair b's clips follow its global token
        MESM_CHECK(launch_mha_small(a, s));
        a.k_cu = nullptr; a.k_enc = 0; a.k2_table = nullptr;
        MESM_CHECK(Lin(R, L.ca_out, d.ao, D, d.t2, D).res(d.t1, D).ln(L.n2).run(s));
        MESM_CHECK(Lin(R, L.l1, d.t2, D, d.hff, FF).act(ACT_PRELU, L.prelu).run(s));
        MESM_CHECK(Lin(R, L.l2, d.hff, FF, tgt_next, D).res(d.t2, D).ln(L.n3).run(s));
        std::swap(tgt, tgt_next);
        // iterative reference refinement
        MESM_CHECK(Lin(R, c->bb0, tgt, D, d.h1, D).act(ACT_RELU).run(s));
        MESM_CHECK(Lin(R, c->bb1, d.h1, D, d.h2, D).act(ACT_RELU).run(s));
        MESM_CHECK(Lin(R, c->bb2, d.h2, D, d.d2, 2).run(s));
        float* nref = (lid != nl - 1) ? d.refs + (size_t)(lid + 1) * R * 2 : d.tmpref;
        MESM_CHECK(launch_ref_update(d.d2, 2, ref, R, nref, s));
        ref = nref;
        MESM_CHECK(launch_layernorm_rows(tgt, R, c->dec_norm.g, c->dec_norm.b, d.hs + (size_t)lid * R * D, s));
    }
    // heads on every intermediate (aux_outputs = all but the last)
    for (int lid = 0; lid < nl; ++lid) {
        const float* hs = d.hs + (size_t)lid * R * D;
        const float* rf = d.refs + (size_t)lid * R * 2;
        float* lo = (lid == nl - 1) ? logits_out : (aux_logits ? aux_logits + lid * aux_layer_stride : nullptr);
        float* so = (lid == nl - 1) ? spans_out : (aux_spans ? aux_spans + lid * aux_layer_stride : nullptr);
        if (lo) MESM_CHECK(Lin(R, c->cls, hs, D, lo, 2).run(s));
        if (so) {
            MESM_CHECK(Lin(R, c->span0, hs, D, d.h1, D).act(ACT_RELU).run(s));
            MESM_CHECK(Lin(R, c->span1, d.h1, D, d.h2, D).act(ACT_RELU).run(s));
            MESM_CHECK(Lin(R, c->span2, d.h2, D, d.d2, 2).run(s));
            MESM_CHECK(launch_ref_update(d.d2, 2, rf, R, so, s));
        }
        if (hs_out) MESM_CHECK(cudaMemcpyAsync(hs_out + lid * hs_layer_stride, hs, (size_t)R * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (refs_out) MESM_CHECK(cudaMemcpyAsync(refs_out + lid * refs_layer_stride, rf, (size_t)R * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return cudaSuccess;
}

}  // namespace mesm

namespace {

int check_cfg(const mesm_cfg* c, std::string& why) {
    if (!c) { why = "cfg is NULL"; return 1; }
    if (c->hidden_dim != D || c->nheads != NH || c->dim_feedforward != FF) { why = "only hidden_dim=256, nheads=8, dim_feedforward=1024 (every shipped MESM config) are supported"; return 1; }
    if (c->num_queries < 1 || c->num_queries > 32) { why = "num_queries must be in [1,32]"; return 1; }
    if (c->n_input_proj != 2) { why = "n_input_proj must be 2"; return 1; }
    if (c->v_feat_dim < 1 || c->t_feat_dim < 1) { why = "feature dims must be positive"; return 1; }
    if (c->dec_layers < 1 || c->enc_layers < 0 || c->t2v_layers < 0 || c->num_recfw_layers < 0 || c->num_recss_layers < 0) { why = "bad layer counts"; return 1; }
    return 0;
}

}  // namespace

namespace mesm {
struct ProfRec { cudaEvent_t a, b; double flops; int M; const char* name; };
static thread_local std::vector<ProfRec> g_prof;
static thread_local std::string g_prof_report;
ProfScope::ProfScope(const char* name_, cudaStream_t s_, double flops_, int M_) : s(s_), name(name_), flops(flops_), M(M_), a(nullptr), on(g_stats.profile) {
    if (on) { cudaEventCreate(&a); cudaEventRecord(a, s); }
}
ProfScope::~ProfScope() {
    if (!on) return;
    ProfRec r; r.a = a; cudaEventCreate(&r.b); cudaEventRecord(r.b, s); r.flops = flops; r.M = M; r.name = name;
    g_prof.push_back(r);
}
static int recon_blk() {               // MESM_RECON_BLOCK=0: per-head batched back-projections on the fp32 SIMT kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("MESM_RECON_BLOCK"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static int force_simt() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MESM_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}
cudaError_t launch_linear(const LinearOp& op, cudaStream_t s) {
    if (op.a_hi) {                                     // A as pre-split planes: the TMA-fed kernel is the only consumer
        if (!linear_tma_eligible(op)) return cudaErrorInvalidValue;
        const double fl = 2.0 * op.M * (double)op.N * (double)op.K;
        ProfScope ps(op.a_lo ? (op.ln_g ? "linear_tma +LN" : (op.ln_stats ? "linear_tma +LN stats" : "linear_tma")) : "input_proj linear_tma fp16 features", s, fl, op.M);
        return launch_linear_tma(op, s);
    }
    const bool tc = !force_simt() && linear_tc_eligible(op);
    const double flops = 2.0 * op.M * (double)op.N * ((double)op.K + op.K2) * (op.nbatch > 1 ? op.nbatch : 1);
    const char* name = "linear_simt";
    if (tc) name = op.K > 1024 ? "input_proj linear_tc fp32 features" : (op.K == 1024 ? "linear_tc K=1024" : (op.N > 256 ? "linear_tc K<=512 N>256" : (op.ln_g ? "linear_tc K<=512 N=256 +LN" : "linear_tc K<=512 N=256")));
    static int shapes = -1;            // MESM_PROFILE_SHAPES=1: one profile row per (kernel, M, N, K) instead of per kernel class
    if (shapes < 0) { const char* e = getenv("MESM_PROFILE_SHAPES"); shapes = (e && e[0] == '1') ? 1 : 0; }
    if (shapes && g_stats.profile) {
        static thread_local std::unordered_map<std::string, std::string> names;      // ProfScope keeps the pointer: the strings must outlive it
        const std::string key = std::string(tc ? "linear_tc " : "linear_simt ") + std::to_string(op.M) + "x" + std::to_string(op.N) + "x" + std::to_string(op.K + op.K2) +
                                (op.nbatch > 1 ? " nb" + std::to_string(op.nbatch) : "") + (op.act ? " act" + std::to_string(op.act) : "") + (op.ln_g ? " ln" : "");
        name = names.emplace(key, key).first->second.c_str();
    }
    ProfScope ps(name, s, flops, op.M);
    return tc ? launch_linear_tc(op, s) : launch_linear_simt(op, s);
}
void profile_collect() {
    std::unordered_map<std::string, std::pair<double, long long>> by;
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        by[r.name].first += ms; by[r.name].second++;
        if (r.flops > 0) {
            g_stats.lin_ms += ms; g_stats.lin_flops += r.flops; g_stats.lin_launches++;
            if (r.M >= 16384) { g_stats.big_ms += ms; g_stats.big_flops += r.flops; g_stats.big_launches++; }
        }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    g_prof_report.clear();
    for (auto& kv : by) g_prof_report += kv.first + "\t" + std::to_string(kv.second.second) + "\t" + std::to_string(kv.second.first) + "\n";
}
const char* profile_report() { return g_prof_report.c_str(); }
}  // namespace mesm

// =====================================================================================================================
extern "C" {

int mesm_abi_version(void) { return MESM_ABI_VERSION; }

const char* mesm_last_error(const mesm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

mesm_ctx* mesm_create(const mesm_cfg* cfg, int device) {
    std::string why;
    if (check_cfg(cfg, why)) { g_create_error = why; return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (mesm_b200 has no CPU fallback)"; return nullptr; }
    if (device < 0 || device >= ndev) { g_create_error = "bad device index"; return nullptr; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) { g_create_error = std::string("mesm_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor); return nullptr; }
    cudaSetDevice(device);
    mesm_ctx* c = new mesm_ctx();
    c->cfg = *cfg;
    c->device = device;
    for (int i = 0; i < mesm_ctx::kTabSlots; ++i) cudaEventCreateWithFlags(&c->tab_event[i], cudaEventDisableTiming);
    return c;
}

void mesm_destroy(mesm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (void* p : ctx->owned) cudaFree(p);
    for (void* p : ctx->owned_host) free(p);
    for (void* p : ctx->owned_tma) tma_free_weights(p);
    for (auto& kv : ctx->w) cudaFree(kv.second.p);
    for (int* p : ctx->graph_tabs) cudaFreeHost(p);
    for (int i = 0; i < mesm_ctx::kTabSlots; ++i) {
        if (ctx->h_tab[i]) cudaFreeHost(ctx->h_tab[i]);
        if (ctx->tab_event[i]) cudaEventDestroy(ctx->tab_event[i]);
    }
    delete ctx;
}

int mesm_set_chunk_pairs(mesm_ctx* ctx, int32_t pairs) {
    if (!ctx || pairs < 1) return 1;
    ctx->chunk_pairs = pairs;
    return 0;
}

int64_t mesm_last_launch_count(const mesm_ctx* ctx) { return ctx ? ctx->last_launches : 0; }
int64_t mesm_last_feature_bytes(const mesm_ctx* ctx) { return ctx ? ctx->last_feature_bytes : 0; }

void mesm_profile_begin(void) {
    g_stats.profile = true;
    g_stats.lin_ms = g_stats.lin_flops = g_stats.lin_bytes = 0; g_stats.lin_launches = 0;
    g_stats.big_ms = g_stats.big_flops = 0; g_stats.big_launches = 0;
}

const char* mesm_profile_report(void) { return mesm::profile_report(); }

void mesm_profile_end(double* out7) {
    mesm::profile_collect();
    g_stats.profile = false;
    if (out7) {
        out7[0] = g_stats.lin_ms; out7[1] = g_stats.lin_flops; out7[2] = g_stats.lin_bytes; out7[3] = (double)g_stats.lin_launches;
        out7[4] = g_stats.big_ms; out7[5] = g_stats.big_flops; out7[6] = (double)g_stats.big_launches;
    }
}

int mesm_load_weight(mesm_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim, int is_device,
                     void* stream) {
    if (!ctx || !key || !data || ndim < 0 || ndim > 4) return fail(ctx, 1, "mesm_load_weight: bad argument");
    CK(cudaSetDevice(ctx->device));
    size_t n = 1;
    std::vector<int64_t> shp;
    for (int i = 0; i < ndim; ++i) { n *= (size_t)shape[i]; shp.push_back(shape[i]); }
    Tensor& t = ctx->w[key];
    if (t.p && t.n != n) { cudaFree(t.p); t.p = nullptr; }
    if (!t.p) CK(cudaMalloc((void**)&t.p, std::max<size_t>(n, 4) * sizeof(float)));
    t.n = n; t.shape = shp;
    CK(cudaMemcpyAsync(t.p, data, n * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    ctx->finalized = false;
    return 0;
}

int mesm_finalize_weights(mesm_ctx* ctx, void* stream) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaStreamSynchronize(s));                  // previous packed buffers may still be in use
    for (void* p : ctx->owned) cudaFree(p);
    ctx->owned.clear();
    for (void* p : ctx->owned_host) free(p);
    ctx->owned_host.clear();
    for (void* p : ctx->owned_tma) tma_free_weights(p);
    ctx->owned_tma.clear();
    ctx->vid0_f16 = nullptr;
    const mesm_cfg& cf = ctx->cfg;
    Packer P{ctx, s};
    // input projections: LinearLayer.0 = LN(in)->Linear->ReLU (LN folded into the GEMM), LinearLayer.1 = LN(256)->Linear
    {
        const Tensor* g = P.get("input_vid_proj.0.LayerNorm.weight", {cf.v_feat_dim});
        const Tensor* b = P.get("input_vid_proj.0.LayerNorm.bias", {cf.v_feat_dim});
        ctx->vid0 = P.pack("input_vid_proj.0.net.1.weight", "input_vid_proj.0.net.1.bias", D, cf.v_feat_dim, 0, D,
                           g ? g->p : nullptr, b ? b->p : nullptr);
        {   // the same folded weights as fp16 hi/lo planes: first projection of 16-bit stored features (exact fp16 A plane)
            const Tensor* W0 = P.get("input_vid_proj.0.net.1.weight", {D, cf.v_feat_dim});
            if (W0 && g) {
                void* wt = tma_pack_weights(W0->p, 0, D, cf.v_feat_dim, g->p, true, s);
                if (wt) { ctx->owned_tma.push_back(wt); ctx->vid0_f16 = wt; }
            }
        }
        ctx->vid1_ln = P.norm("input_vid_proj.1.LayerNorm");
        ctx->vid1 = P.lin("input_vid_proj.1.net.1", D, D);
        g = P.get("input_txt_proj.0.LayerNorm.weight", {cf.t_feat_dim});
        b = P.get("input_txt_proj.0.LayerNorm.bias", {cf.t_feat_dim});
        ctx->txt0 = P.pack("input_txt_proj.0.net.1.weight", "input_txt_proj.0.net.1.bias", D, cf.t_feat_dim, 0, D,
                           g ? g->p : nullptr, b ? b->p : nullptr);
        ctx->txt1_ln = P.norm("input_txt_proj.1.LayerNorm");
        ctx->txt1 = P.lin("input_txt_proj.1.net.1", D, D);
    }
    ctx->enh.clear(); ctx->aln.clear(); ctx->rec.clear(); ctx->enc.clear(); ctx->dec.clear();
    if (cf.rec_fw)
        for (int i = 0; i < cf.num_recfw_layers; ++i) { const size_t m0 = P.missing.size(); AttnFfn L = P.attn_ffn("enhance_encoder.t2v_encoder.layers." + std::to_string(i) + ".", false); if (P.missing.size() == m0) ctx->enh.push_back(L); }
    for (int i = 0; i < cf.t2v_layers; ++i) { const size_t m0 = P.missing.size(); AttnFfn L = P.attn_ffn("t2v_encoder.t2v_encoder.layers." + std::to_string(i) + ".", false); if (P.missing.size() == m0) ctx->aln.push_back(L); }
    if (cf.rec_ss) {
        for (int i = 0; i < cf.num_recss_layers; ++i) { const size_t m0 = P.missing.size(); AttnFfn L = P.attn_ffn("ss_reconstructor.recon_trans.layers." + std::to_string(i) + ".", true); if (P.missing.size() == m0) ctx->rec.push_back(L); }
        ctx->msent = P.vec("ss_reconstructor.masked_sent_token", D);
        ctx->osp0_ln = P.norm("ss_reconstructor.output_sent_proj.0.LayerNorm");
        ctx->osp0 = P.lin("ss_reconstructor.output_sent_proj.0.net.1", D, D);
        ctx->osp1_ln = P.norm("ss_reconstructor.output_sent_proj.1.LayerNorm");
        ctx->osp1 = P.lin("ss_reconstructor.output_sent_proj.1.net.1", D, D);
    }
    for (int i = 0; i < cf.enc_layers; ++i) { const size_t m0 = P.missing.size(); AttnFfn L = P.attn_ffn("transformer.encoder.layers." + std::to_string(i) + ".", false); if (P.missing.size() == m0) ctx->enc.push_back(L); }
    for (int i = 0; i < cf.dec_layers; ++i) {
        const std::string p = "transformer.decoder.layers." + std::to_string(i) + ".";
        const size_t dec_m0 = P.missing.size();
        DecLayer L;
        L.sa_qc = P.lin(p + "sa_qcontent_proj", D, D); L.sa_qp = P.lin(p + "sa_qpos_proj", D, D);
        L.sa_kc = P.lin(p + "sa_kcontent_proj", D, D); L.sa_kp = P.lin(p + "sa_kpos_proj", D, D);
        L.sa_v = P.lin(p + "sa_v_proj", D, D); L.sa_out = P.lin(p + "self_attn.out_proj", D, D);
        L.ca_qc = P.lin(p + "ca_qcontent_proj", D, D); L.ca_kc = P.lin(p + "ca_kcontent_proj", D, D);
        L.ca_kp = P.lin(p + "ca_kpos_proj", D, D); L.ca_v = P.lin(p + "ca_v_proj", D, D);
        L.ca_sine = P.lin(p + "ca_qpos_sine_proj", D, D); L.ca_out = P.lin(p + "cross_attn.out_proj", D, D);
        L.l1 = P.lin(p + "linear1", FF, D); L.l2 = P.lin(p + "linear2", D, FF);
        L.n1 = P.norm(p + "norm1"); L.n2 = P.norm(p + "norm2"); L.n3 = P.norm(p + "norm3");
        L.prelu = P.vec(p + "activation.weight", 1);
        L.sa_q_bias = P.bias_sum(L.sa_qc.bias, L.sa_qp.bias, D);
        L.sa_k_bias = P.bias_sum(L.sa_kc.bias, L.sa_kp.bias, D);
        if (i == 0) {
            L.ca_qp = P.lin(p + "ca_qpos_proj", D, D);
            L.ca_q_bias0 = P.bias_sum(L.ca_qc.bias, L.ca_qp.bias, D);
            L.ca_k_bias0 = P.bias_sum(L.ca_kc.bias, L.ca_kp.bias, D);
        }
        if (P.missing.size() == dec_m0) ctx->dec.push_back(L);
    }
    ctx->dec_norm = P.norm("transformer.decoder.norm");
    ctx->qs0 = P.lin("transformer.decoder.query_scale.layers.0", D, D);
    ctx->qs1 = P.lin("transformer.decoder.query_scale.layers.1", D, D);
    ctx->rph0 = P.lin("transformer.decoder.ref_point_head.layers.0", D, D);
    ctx->rph1 = P.lin("transformer.decoder.ref_point_head.layers.1", D, D);
    ctx->bb0 = P.lin("transformer.decoder.bbox_embed.layers.0", D, D);
    ctx->bb1 = P.lin("transformer.decoder.bbox_embed.layers.1", D, D);
    ctx->bb2 = P.lin("transformer.decoder.bbox_embed.layers.2", 2, D);
    ctx->ra0 = P.lin("transformer.decoder.ref_anchor_head.layers.0", D, D);
    ctx->ra1 = P.lin("transformer.decoder.ref_anchor_head.layers.1", 1, D);
    ctx->span0 = P.lin("span_embed.layers.0", D, D);
    ctx->span1 = P.lin("span_embed.layers.1", D, D);
    ctx->span2 = P.lin("span_embed.layers.2", 2, D);
    ctx->cls = P.lin("class_embed", 2, D);
    ctx->sal1 = P.lin("saliency_proj1", D, D);
    ctx->sal2 = P.lin("saliency_proj2", D, D);
    ctx->gtok = P.vec("global_rep_token", D);
    ctx->gpos = P.vec("global_rep_pos", D);
    {
        const Tensor* q = P.get("query_embed.weight", {cf.num_queries, 2});
        ctx->qembed = q ? q->p : nullptr;
    }
    if (P.cerr != cudaSuccess) return fail(ctx, (int)P.cerr, std::string("weight packing: ") + cudaGetErrorString(P.cerr));
    ctx->missing = P.missing;       // sub-module contexts (T2VEncoder / Transformer) legitimately hold a partial state_dict
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    ctx->finalized = true;
    return 0;
}

}  // extern "C"

// =====================================================================================================================
// forward
// =====================================================================================================================
namespace {

constexpr int kF16ChunkRows = 65536;      // rows of 16-bit features repacked + projected per launch pair (bounds the staging buffer)

struct FwdPlan {
    // whole-batch buffers
    float *wn, *wstat, *t1, *expw, *negw, *projV, *recon;
    uint8_t *wmask, *emask, *epad, *wpad, *neg_epad, *neg_wpad, *padV_all;
    int* d_tab;
    int *t_pad, *t_in, *t_c2e, *t_g, *t_posV, *t_posE;   // packed-layout gather tables (kernels.h: launch_pack_table / launch_chunk_tables)
    int *t_vin, *t_p2v;                                  // per-video tables (launch_video_tables)
    uint16_t* xf16;                                      // staging of the repacked fp16 feature rows (kF16ChunkRows at a time)
    // chunk buffers
    float *vstat, *v1, *posV, *posE, *xa, *xb, *enh, *E, *E2, *P1, *P2, *Qenh0;
    uint8_t *padV, *padE;
    T2VBuffers t2v;
    EncBuffers encb;
    DecBuffers dec;
    float *rS, *rS2, *rq, *rqk, *rpool, *rao, *rX1, *rY1, *rH, *rtmp;
    float *PT, *PWq, *PWqk, *PWkp;
    Planes xaP, xbP, enhP, EP, E2P, aoP, y1P;            // pre-split activations of the chunk (PlaneIO, ctx.h)
    size_t total = 0;
};

void plan_forward(const mesm_ctx* c, Arena& ar, FwdPlan& p, int B, int Lv, int Lt, int G, int Bc, bool need_projV) {
    const mesm_cfg& cf = c->cfg;
    const size_t Rt = (size_t)B * Lt, Rte = (size_t)B * (Lt + 1);
    p.wn = ar.get<float>(Rt * cf.t_feat_dim); p.wstat = ar.get<float>(Rt * 2); p.t1 = ar.get<float>(Rt * D);
    p.expw = ar.get<float>(Rte * D); p.negw = ar.get<float>(Rte * D);
    p.projV = need_projV ? ar.get<float>((size_t)B * Lv * D) : nullptr;
    p.recon = ar.get<float>((size_t)B * D);
    p.wmask = ar.get<uint8_t>(Rt); p.emask = ar.get<uint8_t>(Rte); p.epad = ar.get<uint8_t>(Rte); p.wpad = ar.get<uint8_t>(Rt);
    p.neg_epad = ar.get<uint8_t>(Rte); p.neg_wpad = ar.get<uint8_t>(Rt); p.padV_all = ar.get<uint8_t>((size_t)B * Lv);
    p.d_tab = ar.get<int>((size_t)3 * B + 3 * (G + 1) + 2 + 3 * (size_t)Lv + 2);
    p.t_vin = ar.get<int>((size_t)B * Lv); p.t_p2v = ar.get<int>((size_t)B * Lv);
    p.xf16 = ar.get<uint16_t>((size_t)kF16ChunkRows * ((c->cfg.v_feat_dim + 7) / 8 * 8));
    p.t_posV = ar.get<int>((size_t)Bc * Lv); p.t_posE = ar.get<int>((size_t)Bc * (Lv + 1));
    {   // position table and its products with every projection that consumes positions (packed layout only)
        const size_t nPT = 1 + std::min<size_t>((size_t)Lv * (Lv + 1) / 2, (size_t)B * Lv);
        p.PT = ar.get<float>(nPT * D);
        p.PWq = ar.get<float>(nPT * D * (c->enh.size() + c->aln.size() + 1));
        p.PWqk = ar.get<float>(nPT * 2 * D * (c->enc.size() + 1));
        p.PWkp = ar.get<float>(nPT * D * (c->dec.size() + 1));
    }
    p.t_pad = ar.get<int>((size_t)B * Lv); p.t_in = ar.get<int>((size_t)B * Lv); p.t_c2e = ar.get<int>((size_t)Bc * Lv); p.t_g = ar.get<int>((size_t)Bc);
    p.vstat = ar.get<float>((size_t)B * Lv * 2); p.v1 = ar.get<float>((size_t)B * Lv * D);
    const int L1 = Lv + 1;
    const size_t Rv = (size_t)Bc * Lv, Re = (size_t)Bc * L1, Rk = (size_t)Bc * (Lt + 1);
    p.posV = ar.get<float>(Rv * D); p.posE = ar.get<float>(Re * D);
    p.xa = ar.get<float>(Rv * D); p.xb = ar.get<float>(Rv * D); p.enh = ar.get<float>(Rv * D); p.Qenh0 = ar.get<float>(Rv * D);
    p.E = ar.get<float>(Re * D); p.E2 = ar.get<float>(Re * D); p.P1 = ar.get<float>(Re * D); p.P2 = ar.get<float>((size_t)Bc * D);
    p.padV = ar.get<uint8_t>(Rv); p.padE = ar.get<uint8_t>(Re);
    for (Planes* pp : {&p.xaP, &p.xbP, &p.enhP, &p.EP, &p.E2P, &p.aoP, &p.y1P}) { pp->hi = ar.get<uint16_t>(Re * D); pp->lo = ar.get<uint16_t>(Re * D); }
    p.t2v.KV = ar.get<float>(Rk * 2 * D); p.t2v.Q = ar.get<float>(Re * D); p.t2v.AO = ar.get<float>(Re * D);
    p.t2v.X1 = ar.get<float>(Re * D); p.t2v.Y1 = ar.get<float>(Re * D); p.t2v.H = ar.get<float>(Re * FF);
    p.encb.QKV = ar.get<float>(Re * 3 * D); p.encb.AO = p.t2v.AO; p.encb.Y1 = p.t2v.Y1; p.encb.H = p.t2v.H;
    p.encb.split = attn_split_floats((long long)Re, L1) ? ar.get<float>(attn_split_floats((long long)Re, L1)) : nullptr;
    dec_alloc(ar, p.dec, Bc, cf.num_queries, L1, cf.dec_layers);
    p.rS = ar.get<float>((size_t)B * D); p.rS2 = ar.get<float>((size_t)B * D); p.rq = ar.get<float>((size_t)B * D);
    p.rqk = ar.get<float>((size_t)B * NH * D); p.rpool = ar.get<float>((size_t)B * NH * D); p.rao = ar.get<float>((size_t)B * D);
    p.rX1 = ar.get<float>((size_t)B * D); p.rY1 = ar.get<float>((size_t)B * D); p.rH = ar.get<float>((size_t)B * FF);
    p.rtmp = ar.get<float>((size_t)B * D);
    p.total = ar.off + 256;
}

}  // namespace

extern "C" size_t mesm_workspace_bytes(const mesm_ctx* ctx, int32_t B, int32_t Lv, int32_t Lt, int32_t G) {
    if (!ctx || B < 1 || Lv < 1 || Lt < 1) return 0;
    Arena ar(nullptr, 0);
    FwdPlan p;
    plan_forward(ctx, ar, p, B, Lv, Lt, std::max(G, 1), std::min<int>(B, 2 * ctx->chunk_pairs), true);
    return p.total;
}

extern "C" int mesm_forward(mesm_ctx* ctx, const mesm_inputs* in, const mesm_outputs* out, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!ctx) return 1;
    if (!in || !out || !workspace) return fail(ctx, 1, "mesm_forward: NULL argument");
    if (!ctx->finalized) return fail(ctx, 1, "mesm_forward: weights not finalized (call mesm_finalize_weights)");
    if (!ctx->missing.empty()) return fail(ctx, 2, "mesm_forward: state_dict incomplete or mis-shaped; missing: " + ctx->missing.substr(0, 600));
    const mesm_cfg& cf = ctx->cfg;
    const int B = in->B, Lv = in->Lv, Lt = in->Lt, G = in->G, L1 = Lv + 1, Lk = Lt + 1, nq = cf.num_queries;
    if (B < 1 || Lv < 1 || Lt < 1 || G < 1 || !in->video_feat || !in->video_mask || !in->words_feat || !in->num_clips)
        return fail(ctx, 1, "mesm_forward: bad input shapes / NULL input");
    if (Lv + 1 > 1024) return fail(ctx, 1, "mesm_forward: Lv must be <= 1023");
    if (!cf.rec_ss) return fail(ctx, 1, "mesm_forward: rec_ss = false is not supported (every shipped config enables it)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long launches0 = g_stats.launches;

    // ---- host tables from num_clips ------------------------------------------------------------------------------
    long long tot = 0; int max_nc = 0;
    for (int g = 0; g < G; ++g) { if (in->num_clips[g] < 1) return fail(ctx, 1, "num_clips entries must be >= 1"); tot += in->num_clips[g]; max_nc = std::max<int>(max_nc, (int)in->num_clips[g]); }
    if (tot != B) return fail(ctx, 1, "sum(num_clips) != B");
    if (in->neg_index && G < 2) return fail(ctx, 1, "the negative branch needs >= 2 video groups (sample_outclass_neg raises in the reference)");
    const bool packed = in->video_len != nullptr;       // variable-length clip rows: no work on the zero padding
    const size_t tab_ints = (size_t)3 * B + (G + 1) + 1 + 3 * (size_t)Lv + 1 + (G + 1);
    // Stream capture (CUDA graphs, Engine.capture): the kernel that pulls the host tables is replayed with the graph, so the
    // tables of a captured forward get a pinned buffer of their own that lives as long as the context (relaxed capture mode
    // permits the allocation); the ring and its events serve eager forwards only.
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(s, &cap_status));
    const bool capturing = cap_status == cudaStreamCaptureStatusActive;
    const int ts = ctx->tab_turn;                    // this forward's slot of the pinned-table ring
    int* h_table = nullptr;
    if (capturing) {
        CK(cudaMallocHost((void**)&h_table, tab_ints * sizeof(int)));
        ctx->graph_tabs.push_back(h_table);
    } else {
        ctx->tab_turn = (ts + 1) % mesm_ctx::kTabSlots;
        if (ctx->tab_event_pending[ts]) { CK(cudaEventSynchronize(ctx->tab_event[ts])); ctx->tab_event_pending[ts] = false; }   // its last reader has run
        if (ctx->h_tab_cap[ts] < tab_ints) {
            if (ctx->h_tab[ts]) cudaFreeHost(ctx->h_tab[ts]);
            ctx->h_tab[ts] = nullptr; ctx->h_tab_cap[ts] = 0;
            CK(cudaMallocHost((void**)&ctx->h_tab[ts], tab_ints * 2 * sizeof(int)));
            ctx->h_tab_cap[ts] = tab_ints * 2;
        }
        h_table = ctx->h_tab[ts];
    }
    int* h_group = h_table; int* h_slot = h_group + B; int* h_gstart = h_slot + B; int* h_cu = h_gstart + (G + 1);
    h_cu[0] = 0;
    for (int b = 0; b < B; ++b) {
        const int n = packed ? in->video_len[b] : Lv;
        if (n < 1 || n > Lv) return fail(ctx, 1, "mesm_forward: video_len entries must be in [1, Lv]");
        h_cu[b + 1] = h_cu[b] + n;
    }
    // distinct clip counts -> rows of the position table (row 0 = the global token's position)
    int* h_lenoff = h_cu + (B + 1); int* h_dllen = h_lenoff + (Lv + 1); int* h_dloff = h_dllen + Lv;
    int nd = 0, nPT = 1;
    for (int n = 0; n <= Lv; ++n) h_lenoff[n] = -1;
    if (packed)
        for (int b = 0; b < B; ++b) {
            const int n = in->video_len[b];
            if (h_lenoff[n] < 0) { h_lenoff[n] = nPT; h_dllen[nd] = n; h_dloff[nd] = nPT; nPT += n; ++nd; }
        }
    for (int j = nd; j < Lv; ++j) { h_dllen[j] = 0; h_dloff[j] = 0; }
    std::vector<std::pair<int, int>> chunks;
    {
        // chunks of whole video groups, greedily filled to chunk_pairs; a small tail chunk is re-balanced with its
        // predecessor so that no launch sequence runs on a handful of pairs
        int b = 0, c0 = 0;
        for (int g = 0; g < G; ++g) {
            h_gstart[g] = b;
            const int n = (int)in->num_clips[g];
            if (b > c0 && b + n - c0 > ctx->chunk_pairs) { chunks.push_back({c0, b}); c0 = b; }
            for (int i = 0; i < n; ++i) { h_group[b] = g; h_slot[b] = i; ++b; }
        }
        h_gstart[G] = B;
        chunks.push_back({c0, B});
        const size_t nc = chunks.size();
        if (nc >= 2 && chunks[nc - 1].second - chunks[nc - 1].first < ctx->chunk_pairs / 2) {
            const int lo = chunks[nc - 2].first, hi = B, mid = lo + (hi - lo) / 2;
            int g = h_group[mid], cut = h_gstart[g];                 // split on a group boundary near the middle
            if (cut <= lo) cut = h_gstart[g + 1];
            if (cut > lo && cut < hi) { chunks[nc - 2] = {lo, cut}; chunks[nc - 1] = {cut, hi}; }
        }
    }
    int Bc_max = 0;
    for (auto& ch : chunks) Bc_max = std::max(Bc_max, ch.second - ch.first);
    // videos of the batch: one per group when the group's pairs share it (prefix sums of their clip counts), else one per pair
    const bool shared_video = in->shared_group_video != 0;
    const bool f16 = in->video_feat_f16 != 0;
    if ((shared_video || f16) && !packed) return fail(ctx, 1, "mesm_forward: shared_group_video / video_feat_f16 need video_len");
    if (shared_video && cf.qvh_grouping) return fail(ctx, 1, "mesm_forward: shared_group_video needs the charades / tacos grouping");
    if (f16 && !ctx->vid0_f16) return fail(ctx, 1, "mesm_forward: fp16 features need the TMA weight planes (cuTensorMapEncodeTiled unavailable?)");
    int* h_vcu = h_dloff + Lv;
    h_vcu[0] = 0;
    for (int g = 0; g < G; ++g) h_vcu[g + 1] = h_vcu[g] + (shared_video ? in->video_len[h_gstart[g]] : 0);
    const long long n_vrows = shared_video ? h_vcu[G] : (packed ? h_cu[B] : (long long)B * Lv);

    Arena ar(workspace, workspace_bytes);
    FwdPlan p;
    float* projV_all = out->projed_video_feat;
    plan_forward(ctx, ar, p, B, Lv, Lt, G, Bc_max, projV_all == nullptr || packed);
    if (p.total > workspace_bytes)
        return fail(ctx, 1, "mesm_forward: workspace too small: need " + std::to_string(p.total) + " bytes, got " + std::to_string(workspace_bytes));
    // projV_all: the projected clips in the layout the rest of the forward reads - [B, Lv, 256] zero-padded rows, or
    // packed rows (pair b at row cu[b]) when the caller passed the clip counts
    float* projV_padded_out = packed ? projV_all : nullptr;
    if (!projV_all || packed) projV_all = p.projV;
    int* d_group = p.d_tab; int* d_slot = d_group + B; int* d_gstart = d_slot + B; int* d_cu_all = d_gstart + (G + 1);
    int* d_lenoff = d_cu_all + (B + 1); int* d_dllen = d_lenoff + (Lv + 1); int* d_dloff = d_dllen + Lv;
    int* d_vcu = d_dloff + Lv;                          // per-video prefix sums (shared group videos)
    int* d_glen = d_vcu + (G + 1);
    const int* d_cu = packed ? d_cu_all : nullptr;
    {
        int* h_dev = nullptr;                       // device alias of the pinned table (identical under UVA)
        CK(cudaHostGetDevicePointer((void**)&h_dev, h_group, 0));
        CK(launch_pull_ints(h_dev, d_group, (long long)tab_ints, s));
    }
    if (!capturing) {
        CK(cudaEventRecord(ctx->tab_event[ts], s));
        ctx->tab_event_pending[ts] = true;
    }
    if (cf.qvh_grouping) CK(launch_group_len(in->video_mask, Lv, d_gstart, G, d_glen, s));
    // t_pad: packed clip row -> this pair's row in the zero-padded [B, Lv] layout (outputs); t_in: the row its features are
    // read from (the same, or the group's first pair when the collate-replicated video was uploaded once per group)
    if (packed) CK(launch_pack_table(d_cu, B, Lv, p.t_pad, s));
    if (shared_video) CK(launch_video_tables(d_cu, d_vcu, d_group, d_gstart, B, G, Lv, p.t_vin, p.t_p2v, s));
    const int* t_vin = shared_video ? p.t_vin : p.t_pad;          // video row -> row of the padded input
    // Position terms (packed layout).  PositionEmbeddingSine of a clip depends only on (clip count, clip index): build the
    // table once per distinct clip count of the batch and push it through every projection that adds positions to its
    // input - (x + pos) W = x W + (pos W)[row] - so those GEMMs lose their second K-sweep and the decoder's k_pos GEMMs
    // shrink from B*(Lv+1) rows to the table's rows.
    static int pos_tables_on = -1;
    if (pos_tables_on < 0) { const char* e = getenv("MESM_POS_TABLES"); pos_tables_on = (e && e[0] == '0') ? 0 : 1; }
    const bool ptab = packed && pos_tables_on;
    std::vector<const float*> PWq_enh(ctx->enh.size(), nullptr), PWq_aln(ctx->aln.size(), nullptr), PWqk(ctx->enc.size(), nullptr),
        PWkp(ctx->dec.size(), nullptr);
    if (ptab) {
        CK(launch_pos_table(d_dllen, d_dloff, nd, ctx->gpos, p.PT, s));
        size_t slot = 0;
        auto project = [&](const PL& w0, bool keep_bias, float* dst, int N) -> cudaError_t {
            PL w = w0;
            if (!keep_bias) w.bias = nullptr;
            return Lin(nPT, w, p.PT, D, dst, N).run(s);
        };
        for (size_t l = 0; l < ctx->enh.size(); ++l, ++slot) { float* d = p.PWq + slot * (size_t)nPT * D; CK(project(ctx->enh[l].q, false, d, D)); PWq_enh[l] = d; }
        for (size_t l = 0; l < ctx->aln.size(); ++l, ++slot) { float* d = p.PWq + slot * (size_t)nPT * D; CK(project(ctx->aln[l].q, false, d, D)); PWq_aln[l] = d; }
        for (size_t l = 0; l < ctx->enc.size(); ++l) { float* d = p.PWqk + l * (size_t)nPT * 2 * D; CK(project(ctx->enc[l].qk, false, d, 2 * D)); PWqk[l] = d; }
        for (size_t l = 0; l < ctx->dec.size(); ++l) { float* d = p.PWkp + l * (size_t)nPT * D; CK(project(ctx->dec[l].ca_kp, true, d, D)); PWkp[l] = d; }
    }

    // ---- text side, whole batch (model/model.py:145-152, 167) --------------------------------------------------------
    const int Rt = B * Lt;
    float* expw = out->expanded_words_feat ? out->expanded_words_feat : p.expw;
    uint8_t* emask = out->expanded_words_mask ? out->expanded_words_mask : p.emask;
    CK(launch_text_prep(in->words_feat, Rt, cf.t_feat_dim, p.wn, p.wmask, p.wstat, s));
    CK(Lin(Rt, ctx->txt0, p.wn, cf.t_feat_dim, p.t1, D).fold(p.wstat, ctx->txt0.colsum).act(ACT_RELU).ln(ctx->txt1_ln).run(s));
    CK(Lin(Rt, ctx->txt1, p.t1, D, expw, D).omap(RowMap{Lt, Lk, 1}).run(s));
    CK(launch_expand_mask(p.wmask, B, Lt, emask, p.epad, p.wpad, s));

    CK(launch_invert_mask(in->video_mask, p.padV_all, (long long)B * Lv, s));

    const RowMap wordsMap{Lt, Lk, 1};          // projed_words rows inside the expanded [B, Lt+1, 256] buffer
    float* recon_all = out->recon_feat ? out->recon_feat : p.recon;
    const int recon_max_keys = cf.qvh_grouping ? max_nc * Lv : Lv;

    // ---- whole batch: input projection of the clips (model/model.py:166) — row-wise, no reason to chunk ---------------
    //      LinearLayer 0 (K = v_feat_dim: the one HBM-bound stage) runs once per VIDEO: the queries of a charades / tacos group
    //      share their video (dataset/base.py:307-309), so its projection is computed for the group's first pair only and
    //      LinearLayer 1 gathers it per pair (t_p2v).
    {
        const long long Rall = packed ? h_cu[B] : (long long)B * Lv;
        const RowMap outmap = packed ? table_map(p.t_pad) : identity_map();    // packed row -> row of a padded output
        // LayerNorm(Dv) is folded into the GEMM: out = rstd * (x . W gamma - mean * colsum) + (W beta + b)
        if (f16) {
            // 16-bit stored features: rows gathered into a 16-byte-aligned staging buffer (with their LayerNorm statistics), then
            // ONE exact fp16 A plane x fp16 hi/lo weight planes through the TMA engine (linear_tma.cu), kF16ChunkRows at a time
            const int ldx = (cf.v_feat_dim + 7) / 8 * 8;
            for (long long r0 = 0; r0 < n_vrows; r0 += kF16ChunkRows) {
                const int R = (int)std::min<long long>(kF16ChunkRows, n_vrows - r0);
                CK(launch_repack_f16_rows((const uint16_t*)in->video_feat, t_vin, r0, R, cf.v_feat_dim, p.xf16, ldx, p.vstat + 2 * r0, s));
                Lin k1(R, ctx->vid0, nullptr, 0, p.v1 + r0 * D, D);
                k1.aplanes(p.xf16, nullptr, ldx).wtm(ctx->vid0_f16).fold(p.vstat + 2 * r0, ctx->vid0.colsum).act(ACT_RELU).ln(ctx->vid1_ln);
                CK(k1.run(s));
            }
            ctx->last_feature_bytes = n_vrows * (long long)cf.v_feat_dim * 2;
        } else {
            const float* vf = (const float*)in->video_feat;
            const RowMap inmap = packed ? table_map(t_vin) : identity_map();      // video row -> row of the padded input
            // the row statistics are accumulated by the kernel's operand converters while the features stream through
            Lin fused((int)n_vrows, ctx->vid0, vf, cf.v_feat_dim, p.v1, D);
            fused.amap(inmap).fold_fused(ctx->vid0.colsum).act(ACT_RELU).ln(ctx->vid1_ln);
            if (linear_tc_eligible(fused.op) && !getenv("MESM_FORCE_SIMT")) {
                CK(fused.run(s));
            } else {
                CK(launch_row_stats(vf, n_vrows, cf.v_feat_dim, cf.v_feat_dim, p.vstat, s, packed ? t_vin : nullptr));
                CK(Lin((int)n_vrows, ctx->vid0, vf, cf.v_feat_dim, p.v1, D).amap(inmap).fold(p.vstat, ctx->vid0.colsum).act(ACT_RELU).ln(ctx->vid1_ln).run(s));
            }
            ctx->last_feature_bytes = n_vrows * (long long)cf.v_feat_dim * 4;
        }
        Lin second((int)Rall, ctx->vid1, p.v1, D, projV_all, D);
        if (shared_video) second.amap(table_map(p.t_p2v));                    // pair row -> its video's row
        if (projV_padded_out) {           // the caller's [B, Lv, 256] output: valid rows scattered, pad rows zero
            second.op.out2 = projV_padded_out; second.op.ldo2 = D; second.op.o2map = outmap;
        }
        CK(second.run(s));
        if (projV_padded_out) CK(launch_zero_masked_rows(projV_padded_out, in->video_mask, (long long)B * Lv, D, s));
    }
    // ---- whole batch: SS-MESM sentence reconstruction (model/model.py:184-222, 467-488).  One masked sentence slot per
    //      pair -> M = B rows; the per-head back-projections are batched over blockIdx.z. ------------------------------
    {
        float* S = p.rS; float* S2 = p.rS2;
        CK(launch_broadcast_row(ctx->msent, S, B, s));
        for (size_t l = 0; l < ctx->rec.size(); ++l) {
            const AttnFfn& L = ctx->rec[l];
            CK(Lin(B, L.q, S, D, p.rq, D).scale(kScale32).run(s));
            if (L.blk_qk.Wp && recon_blk()) {   // qk[b,h,:] = Wk_h^T q_h for all heads: one dense GEMM on the block-structured weights
                CK(Lin(B, L.blk_qk, p.rq, D, p.rqk, NH * D).run(s));
            } else {                            // per head [B,32] x [32,256], batched over blockIdx.z (fp32 SIMT kernel)
                PL w; w.Wt = L.in_w + (size_t)D * D; w.ldw = D; w.K = HD; w.N = D; w.bias = nullptr;
                CK(Lin(B, w, p.rq, D, p.rqk, NH * D).batch(NH, HD, (long long)HD * D, 0, D).run(s));
            }
            ReconPoolArgs ra;
            ra.x = projV_all; ra.ldx = D; ra.vmask = in->video_mask; ra.qk = p.rqk; ra.pooled = p.rpool;
            ra.pair_group = d_group; ra.pair_slot = d_slot; ra.group_start = d_gstart; ra.group_len = d_glen;
            ra.B = B; ra.Lv = Lv; ra.qvh = cf.qvh_grouping; ra.max_keys = recon_max_keys; ra.b0 = 0; ra.Btot = B;
            ra.x_start = d_cu;
            CK(launch_recon_pool(ra, s));
            if (L.blk_v.Wp && recon_blk()) {    // attn_out = Wvo pooled + bv (all heads at once)
                CK(Lin(B, L.blk_v, p.rpool, NH * D, p.rao, D).run(s));
            } else {                            // attn_out[:, h*32:+32] = Wv_h pooled_h + bv_h
                PL w; w.Wt = L.vT; w.ldw = L.v.ldw; w.K = D; w.N = HD; w.bias = L.v.bias;
                CK(Lin(B, w, p.rpool, NH * D, p.rao, D).batch(NH, D, HD, HD, HD).run(s));
            }
            CK(Lin(B, L.out, p.rao, D, p.rY1, D).res(S, D).pre_ln(p.rX1).ln(L.n1).run(s));
            CK(Lin(B, L.l1, p.rY1, D, p.rH, FF).act(ACT_PRELU, L.prelu).run(s));
            CK(Lin(B, L.l2, p.rH, FF, S2, D).res(p.rX1, D).ln(L.n2).run(s));
            std::swap(S, S2);
        }
        // recon_feat = F.normalize(.) -> word slot 0 of the expanded text (model/model.py:217, 486)
        CK(launch_l2norm_rows(S, B, recon_all, expw, D, RowMap{1, Lk, 0}, s));
        if (out->projed_recon_feat) {           // output_sent_proj (model/model.py:487)
            CK(launch_layernorm_rows(recon_all, B, ctx->osp0_ln.g, ctx->osp0_ln.b, p.rtmp, s));
            CK(Lin(B, ctx->osp0, p.rtmp, D, p.rY1, D).act(ACT_RELU).ln(ctx->osp1_ln).run(s));
            CK(Lin(B, ctx->osp1, p.rY1, D, out->projed_recon_feat, D).run(s));
        }
    }

    // One chunk of whole video groups through enhance -> align -> encoder -> (decoder, heads).
    auto video_chunk = [&](int b0, int b1, bool neg) -> int {
        // Row layout of the chunk: uniform ([Bc, Lv] clips, [Bc, Lv+1] encoder rows) or packed (pair b owns
        // cu[b+1]-cu[b] clip rows; its encoder rows are those plus a leading global-token row).
        const int Bc = b1 - b0;
        const int Rv = packed ? h_cu[b1] - h_cu[b0] : Bc * Lv, Re = Rv + Bc;
        const int* cu = packed ? d_cu + b0 : nullptr;
        const long long v0 = packed ? h_cu[b0] : (long long)b0 * Lv;          // first clip row of the chunk in projV_all
        const RowMap c2e = packed ? table_map(p.t_c2e) : RowMap{Lv, L1, 1};  // clip row -> encoder-buffer row
        const uint8_t* vmask = in->video_mask + (size_t)b0 * Lv;
        float* projV = projV_all + (size_t)v0 * D;
        if (packed && !neg) CK(launch_chunk_tables(cu, Bc, p.t_c2e, p.t_g, s, ptab ? d_lenoff : nullptr, p.t_posV, p.t_posE));
        PosArgs pa; pa.vmask = vmask; pa.B = Bc; pa.Lv = Lv; pa.gtok = ctx->gtok; pa.gpos = ctx->gpos; pa.cu = cu;
        pa.posV = p.posV; pa.posE = p.posE; pa.encbuf = p.E; pa.padV = p.padV; pa.padE = p.padE;
        pa.enc_hi = p.EP.hi; pa.enc_lo = p.EP.lo;
        if (ptab) { pa.posV = nullptr; pa.posE = nullptr; }                                        // positions come from the table
        if (neg) { pa.posV = nullptr; pa.posE = nullptr; pa.padV = nullptr; pa.padE = nullptr; }   // positions / pads kept from the main pass
        CK(launch_pos_embed(pa, s));
        const float* words_c = (neg ? p.negw : expw) + (size_t)b0 * Lk * D;
        const uint8_t* epad_all = neg ? p.neg_epad : p.epad;
        const uint8_t* wpad_all = neg ? p.neg_wpad : p.wpad;
        // ---- FW-MESM enhance encoder (model/model.py:175-182; neg: 281-286): keys = the Lt projected words ----
        const float* x = projV;
        float* enh_out = (!neg && out->enhanced_video_feat) ? out->enhanced_video_feat + (size_t)b0 * Lv * D : nullptr;
        float* enh = (enh_out && !packed) ? enh_out : p.enh;
        Planes xP;                                      // planes of `x` (empty: the producer of x wrote fp32 only)
        static int po_on = -1;
        if (po_on < 0) { const char* e = getenv("MESM_PLANES_ONLY"); po_on = (e && e[0] == '0') ? 0 : 1; }
        const bool planes_only_ok = po_on && ptab;       // intermediate layer outputs as planes only (next layer: TMA operand + planes residual)
        for (size_t l = 0; l < ctx->enh.size(); ++l) {
            const bool lastl = (l + 1 == ctx->enh.size());
            float* dst = lastl ? enh : (l % 2 == 0 ? p.xa : p.xb);
            PlaneIO pio;
            pio.in = xP; pio.out = lastl ? p.enhP : (l % 2 == 0 ? p.xaP : p.xbP); pio.ao = p.aoP; pio.y1 = p.y1P;
            pio.planes_only = planes_only_ok && !lastl;            // the last layer's fp32 rows are an output (enhanced_video_feat)
            T2VBuffers tb = p.t2v;
            if (l == 0) tb.Q = p.Qenh0;                 // layer 0's Q = (projV + pos) Wq is identical in the negative pass
            CK(t2v_layer(ctx->enh[l], words_c, wordsMap, nullptr, Lt, x, p.posV, Lv, Bc, b0, B, p.padV_all, wpad_all, tb,
                         dst, D, identity_map(), s, l == 0 && neg, cu, Rv, Lv, PWq_enh[l], p.t_posV, &pio));
            x = pio.wrote_fp32 ? dst : nullptr;         // nullptr: this activation only exists as planes
            xP = pio.wrote_out ? pio.out : Planes();
        }
        if (ctx->enh.empty() && enh_out && !packed)
            CK(cudaMemcpyAsync(enh, projV, (size_t)Rv * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
        const float* xin = ctx->enh.empty() ? projV : enh;
        if (enh_out && packed) {            // the caller's [B, Lv, 256] output: valid rows scattered, pad rows zero
            CK(launch_copy_rows(xin, D, identity_map(), out->enhanced_video_feat, D, table_map(p.t_pad + v0), Rv, s));
            CK(launch_zero_masked_rows(enh_out, vmask, (long long)Bc * Lv, D, s));
        }
        // ---- aligner (model/model.py:230-234; neg: 290-294): keys = recon token + words; last layer writes the
        //      encoder buffer [Bc, Lv+1, 256] behind the global token ----
        Planes xinP = ctx->enh.empty() ? Planes() : xP;
        bool E_fp32 = true;                              // the fp32 encoder buffer holds the aligner's result
        for (size_t l = 0; l < ctx->aln.size(); ++l) {
            const bool last = (l + 1 == ctx->aln.size());
            float* dst = last ? p.E : (l % 2 == 0 ? p.xa : p.xb);
            PlaneIO pio;
            pio.in = xinP; pio.out = last ? p.EP : (l % 2 == 0 ? p.xaP : p.xbP); pio.ao = p.aoP; pio.y1 = p.y1P;
            pio.planes_only = planes_only_ok && (!last || !ctx->enc.empty());      // the encoder reads E as planes (operand + residual)
            CK(t2v_layer(ctx->aln[l], words_c, identity_map(), nullptr, Lk, xin, p.posV, Lv, Bc, b0, B, p.padV_all, epad_all,
                         p.t2v, dst, D, last ? c2e : identity_map(), s, false, cu, Rv, Lv, PWq_aln[l], p.t_posV, &pio));
            xin = pio.wrote_fp32 ? dst : nullptr;
            E_fp32 = pio.wrote_fp32;
            xinP = pio.wrote_out ? pio.out : Planes();
        }
        Planes EcurP = ctx->aln.empty() ? Planes() : xinP, EnextP = p.E2P;      // planes of the encoder buffer (empty: fp32 only)
        if (ctx->aln.empty())
            CK(launch_copy_rows(xin, D, identity_map(), p.E, D, c2e, Rv, s));
        // ---- transformer encoder (model/transformer.py:185-197) ----
        float* Ecur = p.E; float* Enext = p.E2;
        for (size_t l = 0; l < ctx->enc.size(); ++l) {
            PlaneIO pio;
            pio.in = EcurP; pio.out = EnextP; pio.ao = p.aoP; pio.y1 = p.y1P;
            pio.planes_only = planes_only_ok && (l + 1 < ctx->enc.size());        // the last layer's fp32 rows feed the heads / outputs
            CK(enc_layer(ctx->enc[l], E_fp32 ? Ecur : nullptr, p.posE, p.padE, L1, Bc, p.encb, Enext, s, cu, Re, PWqk[l], p.t_posE, &pio));
            E_fp32 = pio.wrote_fp32;
            std::swap(Ecur, Enext);
            const Planes done = pio.wrote_out ? pio.out : Planes();
            EnextP = (pio.out.hi == p.E2P.hi) ? p.EP : p.E2P;
            EcurP = done;
        }
        // ---- saliency head (model/model.py:301-302) ----
        float* sal = neg ? out->neg_saliency_scores : out->saliency_scores;
        if (sal) {
            {
                Lin s1(Re, ctx->sal1, Ecur, D, p.P1, D);
                if (EcurP && ctx->sal1.Wtm) s1.aplanes(EcurP.hi, EcurP.lo, D);
                CK(s1.run(s));
            }
            if (packed) {
                CK(Lin(Bc, ctx->sal2, Ecur, D, p.P2, D).amap(table_map(p.t_g)).run(s));
                CK(launch_saliency_packed(p.P1, cu, p.P2, Bc, Lv, sal + (size_t)b0 * Lv, s));
            } else {
                CK(Lin(Bc, ctx->sal2, Ecur, L1 * D, p.P2, D).run(s));
                CK(launch_saliency(p.P1, c2e, p.P2, Bc, Lv, sal + (size_t)b0 * Lv, s));
            }
        }
        if (neg) return 0;
        if (out->memory) {
            if (packed) {
                CK(launch_copy_rows(Ecur, D, c2e, out->memory, D, table_map(p.t_pad + v0), Rv, s));
                CK(launch_zero_masked_rows(out->memory + (size_t)b0 * Lv * D, vmask, (long long)Bc * Lv, D, s));
            } else {
                CK(launch_copy_rows(Ecur, D, c2e, out->memory + (size_t)b0 * Lv * D, D, identity_map(), Rv, s));
            }
        }
        if (out->memory_global)
            CK(launch_copy_rows(Ecur, D, packed ? table_map(p.t_g) : RowMap{1, L1, 0}, out->memory_global + (size_t)b0 * D, D, identity_map(), Bc, s));
        // ---- DAB-DETR decoder + class / span heads ----
        if (out->pred_logits || out->pred_spans || out->aux_logits || out->aux_spans || out->hs) {
            cudaError_t e = run_decoder(ctx, ctx->qembed, Ecur, p.posE, p.padV, Lv, Bc, p.dec,
                                        out->pred_logits ? out->pred_logits + (size_t)b0 * nq * 2 : nullptr,
                                        out->pred_spans ? out->pred_spans + (size_t)b0 * nq * 2 : nullptr,
                                        out->aux_logits ? out->aux_logits + (size_t)b0 * nq * 2 : nullptr,
                                        out->aux_spans ? out->aux_spans + (size_t)b0 * nq * 2 : nullptr, (long long)B * nq * 2,
                                        out->hs ? out->hs + (size_t)b0 * nq * D : nullptr, (long long)B * nq * D, nullptr, 0, s,
                                        cu, Re, ptab ? PWkp.data() : nullptr, p.t_posE, EcurP);
            CK(e);
        }
        return 0;
    };

    // ---- negative branch (model/model.py:260-302): every pair re-scored against the text of another video group.  The
    //      gathered text is needed up front so that each chunk runs its negative pass right after its main pass (the
    //      chunk's projected clips, positions and layer-0 queries are still in L2 / reused).
    const bool do_neg = in->neg_index && out->neg_saliency_scores;
    if (do_neg) {
        CK(launch_gather_blocks(expw, p.negw, in->neg_index, B, (long long)Lk * D, p.epad, p.neg_epad, Lk, s));
        CK(launch_expand_mask_from_epad(p.neg_epad, B, Lt, p.neg_wpad, s));
    }
    for (auto& ch : chunks) {
        int rc = video_chunk(ch.first, ch.second, false);
        if (rc) return rc;
        if (do_neg) { rc = video_chunk(ch.first, ch.second, true); if (rc) return rc; }
    }
    CK(cudaGetLastError());
    ctx->last_launches = g_stats.launches - launches0;
    return 0;
}
