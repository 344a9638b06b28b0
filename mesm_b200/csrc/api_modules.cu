// Sub-module entry points of the C ABI: alignment scores, projection-free MHA, T2VEncoder and Transformer
// (drop-in classes of mesm_b200/model.py call these; the fused mesm_forward does not go through them).
#include "ctx.h"

#include <cstdlib>
#include <cstring>

using namespace mesm;
namespace mesm { void tc_read_watchdog(unsigned long long* out8); void tc_read_watchdog_linear(unsigned long long* out8); int tc_read_attn_trace(long long* out128); }

namespace {
__global__ void build_enc_kernel(const float* __restrict__ src, const float* __restrict__ pos, const uint8_t* __restrict__ pad,
                                 const float* __restrict__ gtok, const float* __restrict__ gpos, int B, int L,
                                 float* __restrict__ E, float* __restrict__ posE, uint8_t* __restrict__ padE) {
    const int b = blockIdx.y, i = blockIdx.x, c = threadIdx.x;      // i in [0, L]
    const long long dst = ((long long)b * (L + 1) + i) * D + c;
    if (i == 0) {
        E[dst] = gtok[c]; posE[dst] = gpos[c];
        if (c == 0) padE[(long long)b * (L + 1)] = 1;
    } else {
        const long long srow = ((long long)b * L + i - 1);
        E[dst] = src[srow * D + c]; posE[dst] = pos[srow * D + c];
        if (c == 0) padE[(long long)b * (L + 1) + i] = pad[srow];
    }
}

// Zero the clip rows the ragged upload skipped (pad rows, mask == 0): one CTA per row of `row_bytes` bytes (a multiple of 2).
__global__ void zero_pad_rows_kernel(uint8_t* __restrict__ feat, const uint8_t* __restrict__ mask, long long rows, int row_bytes) {
    const long long r = blockIdx.x;
    if (r >= rows || mask[r]) return;
    uint8_t* q = feat + r * (long long)row_bytes;
    if ((row_bytes & 7) == 0 && (reinterpret_cast<uintptr_t>(feat) & 7) == 0) {
        float2* p = reinterpret_cast<float2*>(q);
        for (int c = threadIdx.x; c < row_bytes / 8; c += blockDim.x) p[c] = make_float2(0.f, 0.f);
    } else if ((row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 3) == 0) {
        float* p = reinterpret_cast<float*>(q);
        for (int c = threadIdx.x; c < row_bytes / 4; c += blockDim.x) p[c] = 0.f;
    } else {
        uint16_t* p = reinterpret_cast<uint16_t*>(q);
        for (int c = threadIdx.x; c < row_bytes / 2; c += blockDim.x) p[c] = 0;
    }
}
}  // namespace

extern "C" {

size_t mesm_align_workspace_bytes(int32_t B) {
    const size_t Bp = ((size_t)B + 3) / 4 * 4;
    return ((size_t)B * D + (size_t)D * Bp) * sizeof(float) + 1024;
}

int mesm_align_scores(const float* projed_video_feat, const uint8_t* clip_mask, const float* expanded_words_feat,
                      const uint8_t* expanded_words_mask, int32_t B, int32_t Lv, int32_t Lw, float tau, float* scores,
                      void* workspace, size_t workspace_bytes, void* stream) {
    mesm_ctx* ctx = nullptr;
    if (!projed_video_feat || !clip_mask || !expanded_words_feat || !expanded_words_mask || !scores || !workspace || B < 1)
        return fail(ctx, 1, "mesm_align_scores: bad argument");
    if (workspace_bytes < mesm_align_workspace_bytes(B)) return fail(ctx, 1, "mesm_align_scores: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int Bp = (B + 3) / 4 * 4;
    Arena ar(workspace, workspace_bytes);
    float* clipn = ar.get<float>((size_t)B * D);
    float* wordsT = ar.get<float>((size_t)D * Bp);
    CK(launch_masked_mean_norm(projed_video_feat, clip_mask, B, Lv, clipn, D, 0, s));
    CK(launch_masked_mean_norm(expanded_words_feat, expanded_words_mask, B, Lw, wordsT, Bp, 1, s));
    LinearOp op = make_linear(B, B, D, clipn, D, wordsT, Bp, nullptr, scores, B);
    op.out_scale = 1.f / tau;
    CK(launch_linear(op, s));
    return 0;
}

// cudaMemcpyBatchAsync (CUDA 12.8+): one driver call for the ~800 ragged copies of a batch instead of one each.  Called
// directly through the prototype of the toolkit this library is compiled AND linked against (libcudart.so.12 is a versioned
// dependency of the .so), never through dlsym: CUDA 13 dropped the `failIdx` argument, so a run-time lookup could bind a
// function with a different signature.  MESM_UPLOAD_BATCH=0 forces the per-copy path.
#if defined(CUDART_VERSION) && CUDART_VERSION >= 12080
#define MESM_HAVE_MEMCPY_BATCH 1
static cudaError_t memcpy_batch_h2d(std::vector<void*>& dst, std::vector<void*>& src, std::vector<size_t>& size, cudaStream_t s) {
    cudaMemcpyAttributes at;
    memset(&at, 0, sizeof(at));
    at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    size_t at_idx = 0;
#if CUDART_VERSION >= 13000
    return cudaMemcpyBatchAsync(dst.data(), src.data(), size.data(), dst.size(), &at, &at_idx, 1, s);
#else
    size_t fail_idx = 0;
    return cudaMemcpyBatchAsync(dst.data(), src.data(), size.data(), dst.size(), &at, &at_idx, 1, &fail_idx, s);
#endif
}
#else
#define MESM_HAVE_MEMCPY_BATCH 0
static cudaError_t memcpy_batch_h2d(std::vector<void*>&, std::vector<void*>&, std::vector<size_t>&, cudaStream_t) { return cudaErrorNotSupported; }
#endif
static bool use_memcpy_batch() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MESM_UPLOAD_BATCH"); v = (MESM_HAVE_MEMCPY_BATCH && !(e && e[0] == '0')) ? 1 : 0; }
    return v == 1;
}

static int upload_clips_impl(const uint8_t* host_feat, const uint8_t* host_mask, int32_t B, int32_t L, int32_t Dv, int elem, uint8_t* dev_feat,
                             uint8_t* dev_mask, const int64_t* num_clips, int32_t G, int64_t* bytes_copied, void* stream) {
    mesm_ctx* ctx = nullptr;
    if (!host_feat || !host_mask || !dev_feat || !dev_mask || B < 1 || L < 1 || Dv < 1) return fail(ctx, 1, "mesm_upload_clips: bad argument");
    if (num_clips) {
        long long tot = 0;
        for (int g = 0; g < G; ++g) { if (num_clips[g] < 1) return fail(ctx, 1, "mesm_upload_clips: num_clips entries must be >= 1"); tot += num_clips[g]; }
        if (tot != B) return fail(ctx, 1, "mesm_upload_clips: sum(num_clips) != B");
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t row = (size_t)Dv * elem;
    int64_t total = (int64_t)B * L;
    CK(cudaMemcpyAsync(dev_mask, host_mask, (size_t)B * L, cudaMemcpyHostToDevice, s));
    // valid rows are a prefix of each pair (utils/data_utils.py:78-82); a pair whose mask is not a prefix is copied up to its
    // last valid row.  Adjacent spans (a full-length pair followed by the next pair's prefix) are merged into one copy.
    long long run_start = -1, run_rows = 0;        // in rows of the flat [B*L] row index
    const bool batch = use_memcpy_batch();
    std::vector<void*> b_dst, b_src;
    std::vector<size_t> b_size;
    auto flush = [&]() -> cudaError_t {
        if (run_rows <= 0) return cudaSuccess;
        cudaError_t e = cudaSuccess;
        if (batch) {
            b_dst.push_back(dev_feat + run_start * row); b_src.push_back(const_cast<uint8_t*>(host_feat) + run_start * row);
            b_size.push_back((size_t)run_rows * row);
        } else {
            e = cudaMemcpyAsync(dev_feat + run_start * row, host_feat + run_start * row, (size_t)run_rows * row, cudaMemcpyHostToDevice, s);
        }
        total += run_rows * (long long)row;
        run_rows = 0; run_start = -1;
        return e;
    };
    int g = 0; long long g_first = 0;               // current video group and its first pair
    for (int b = 0; b < B; ++b) {
        if (num_clips) {
            while (b >= g_first + num_clips[g]) { g_first += num_clips[g]; ++g; }
            if (b != g_first) continue;             // shared video: only the group's first pair crosses the bus
        }
        const uint8_t* m = host_mask + (size_t)b * L;
        int n = L;
        while (n > 0 && !m[n - 1]) --n;            // rows up to the last valid one
        const long long first = (long long)b * L;
        if (run_rows > 0 && run_start + run_rows != first) CK(flush());
        if (n > 0) {
            if (run_rows == 0) run_start = first;
            run_rows += n;
        }
        if (n < L) CK(flush());
    }
    CK(flush());
    if (batch && !b_dst.empty()) {
        cudaError_t e = memcpy_batch_h2d(b_dst, b_src, b_size, s);
        if (e != cudaSuccess) {                      // e.g. an older driver: fall back to one copy per run
            (void)cudaGetLastError();
            for (size_t i = 0; i < b_dst.size(); ++i) CK(cudaMemcpyAsync(b_dst[i], b_src[i], b_size[i], cudaMemcpyHostToDevice, s));
        }
    }
    zero_pad_rows_kernel<<<(unsigned)((long long)B * L), 128, 0, s>>>(dev_feat, dev_mask, (long long)B * L, (int)row);
    CK(cudaGetLastError());
    if (bytes_copied) *bytes_copied = total;
    return 0;
}

int mesm_upload_clips(const float* host_feat, const uint8_t* host_mask, int32_t B, int32_t L, int32_t Dv, float* dev_feat,
                      uint8_t* dev_mask, const int64_t* num_clips, int32_t G, int64_t* bytes_copied, void* stream) {
    return upload_clips_impl((const uint8_t*)host_feat, host_mask, B, L, Dv, 4, (uint8_t*)dev_feat, dev_mask, num_clips, G, bytes_copied, stream);
}

int mesm_upload_clips_f16(const void* host_feat, const uint8_t* host_mask, int32_t B, int32_t L, int32_t Dv, void* dev_feat,
                          uint8_t* dev_mask, const int64_t* num_clips, int32_t G, int64_t* bytes_copied, void* stream) {
    return upload_clips_impl((const uint8_t*)host_feat, host_mask, B, L, Dv, 2, (uint8_t*)dev_feat, dev_mask, num_clips, G, bytes_copied, stream);
}

int mesm_memcpy_batch_h2d(const void* const* src, void* const* dst, const size_t* bytes, int64_t n, void* stream) {
    mesm_ctx* ctx = nullptr;
    if (n < 0 || (n > 0 && (!src || !dst || !bytes))) return fail(ctx, 1, "mesm_memcpy_batch_h2d: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<void*> d, h;
    std::vector<size_t> sz;
    for (int64_t i = 0; i < n; ++i)
        if (bytes[i]) { d.push_back(dst[i]); h.push_back(const_cast<void*>(src[i])); sz.push_back(bytes[i]); }
    if (d.empty()) return 0;
    if (use_memcpy_batch() && memcpy_batch_h2d(d, h, sz, s) == cudaSuccess) return 0;
    (void)cudaGetLastError();
    for (size_t i = 0; i < d.size(); ++i) CK(cudaMemcpyAsync(d[i], h[i], sz[i], cudaMemcpyHostToDevice, s));
    return 0;
}

size_t mesm_mha_workspace_bytes(int32_t L, int32_t S, int32_t B, int32_t E, int32_t Ev) {
    (void)S; (void)E;
    const size_t Kp = ((size_t)Ev + 15) / 16 * 16, ldw = ((size_t)Ev + 3) / 4 * 4;
    return ((size_t)L * B * Ev + Kp * ldw) * sizeof(float) + 1024;
}

int mesm_mha_noproj(const float* q, const float* k, const float* v, int32_t L, int32_t S, int32_t B, int32_t E, int32_t Ev,
                    int32_t nheads, const float* out_w, const float* out_b, const uint8_t* key_padding_mask, float* out,
                    float* attn_weights, void* workspace, size_t workspace_bytes, void* stream) {
    mesm_ctx* ctx = nullptr;
    if (!q || !k || !v || !out_w || !out || !workspace || L < 1 || S < 1 || B < 1 || nheads < 1 || E % nheads || Ev % nheads ||
        (E / nheads) % 4)
        return fail(ctx, 1, "mesm_mha_noproj: bad argument");
    if (workspace_bytes < mesm_mha_workspace_bytes(L, S, B, E, Ev)) return fail(ctx, 1, "mesm_mha_noproj: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    float* attn = ar.get<float>((size_t)L * B * Ev);
    const int Kp = (Ev + 15) / 16 * 16, ldw = (Ev + 3) / 4 * 4;
    float* Wt = ar.get<float>((size_t)Kp * ldw);
    CK(launch_transpose_pack(out_w, 0, Ev, Ev, Wt, ldw, Kp, s));
    if (attn_weights) CK(launch_fill(attn_weights, (long long)B * L * S, 0.f, s));
    MhaSmallArgs a;
    a.q = q; a.ldq = E; a.q2 = nullptr; a.ldq2 = 0; a.k = k; a.ldk = E; a.k2 = nullptr; a.ldk2 = 0; a.v = v; a.ldv = Ev;
    a.k_pad = key_padding_mask; a.out = attn; a.ldo = Ev; a.attn_w = attn_weights;
    a.B = B; a.L = L; a.S = S; a.nheads = nheads; a.hq = E / nheads; a.hv = Ev / nheads;
    a.scale = 1.f / sqrtf((float)(E / nheads));
    a.q_bs = 1; a.q_is = B; a.k_bs = 1; a.k_is = B; a.k_off = 0;          // seq-first [L,B,E] like the reference
    CK(launch_mha_small(a, s));
    LinearOp op = make_linear(L * B, Ev, Ev, attn, Ev, Wt, ldw, out_b, out, Ev);
    CK(launch_linear(op, s));
    return 0;
}

size_t mesm_t2v_workspace_bytes(int32_t B, int32_t Lt, int32_t Lv) {
    const size_t Rv = (size_t)B * Lv, Rt = (size_t)B * Lt;
    return (Rt * 2 * D + Rv * D * 6 + Rv * FF) * sizeof(float) + 4096;
}

int mesm_t2v_encoder(mesm_ctx* ctx, const char* prefix, const float* src_txt, const float* src_vid, const uint8_t* txt_pad,
                     const uint8_t* vid_pad, const float* pos_txt, const float* pos_vid, int32_t B, int32_t Lt, int32_t Lv,
                     float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!ctx) return 1;
    if (!ctx->finalized) return fail(ctx, 1, "mesm_t2v_encoder: weights not finalized");
    if (!prefix || !src_txt || !src_vid || !txt_pad || !vid_pad || !out || !workspace || B < 1 || Lt < 1 || Lv < 1 || Lv > 1024)
        return fail(ctx, 1, "mesm_t2v_encoder: bad argument");
    const std::vector<AttnFfn>* layers = nullptr;
    const std::string p(prefix);
    if (p == "enhance_encoder") layers = &ctx->enh;
    else if (p == "t2v_encoder") layers = &ctx->aln;
    else return fail(ctx, 1, "mesm_t2v_encoder: prefix must be enhance_encoder or t2v_encoder");
    if ((int)layers->size() != (p == "enhance_encoder" ? ctx->cfg.num_recfw_layers : ctx->cfg.t2v_layers))
        return fail(ctx, 2, "mesm_t2v_encoder: layer weights incomplete; missing: " + ctx->missing.substr(0, 600));
    if (workspace_bytes < mesm_t2v_workspace_bytes(B, Lt, Lv)) return fail(ctx, 1, "mesm_t2v_encoder: workspace too small");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    const size_t Rv = (size_t)B * Lv, Rt = (size_t)B * Lt;
    T2VBuffers t;
    t.KV = ar.get<float>(Rt * 2 * D); t.Q = ar.get<float>(Rv * D); t.AO = ar.get<float>(Rv * D); t.X1 = ar.get<float>(Rv * D);
    t.Y1 = ar.get<float>(Rv * D); t.H = ar.get<float>(Rv * FF);
    float* xa = ar.get<float>(Rv * D); float* xb = ar.get<float>(Rv * D);
    const float* x = src_vid;
    for (size_t l = 0; l < layers->size(); ++l) {
        float* dst = (l + 1 == layers->size()) ? out : (l % 2 == 0 ? xa : xb);
        CK(t2v_layer((*layers)[l], src_txt, identity_map(), pos_txt, Lt, x, pos_vid, Lv, B, 0, B, vid_pad, txt_pad, t, dst, D,
                     identity_map(), s));
        x = dst;
    }
    if (layers->empty()) CK(cudaMemcpyAsync(out, src_vid, Rv * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

size_t mesm_transformer_workspace_bytes(const mesm_ctx* ctx, int32_t B, int32_t L) {
    if (!ctx) return 0;
    Arena ar(nullptr, 0);
    const size_t Re = (size_t)B * (L + 1);
    ar.get<float>(Re * D); ar.get<float>(Re * D); ar.get<float>(Re * D); ar.get<uint8_t>(Re);
    ar.get<float>(Re * 3 * D); ar.get<float>(Re * D); ar.get<float>(Re * D); ar.get<float>(Re * FF);
    if (attn_split_floats((long long)Re, L + 1)) ar.get<float>(attn_split_floats((long long)Re, L + 1));
    DecBuffers d;
    dec_alloc(ar, d, B, ctx->cfg.num_queries, L + 1, ctx->cfg.dec_layers);
    return ar.off + 4096;
}

int mesm_transformer(mesm_ctx* ctx, const float* src, const uint8_t* pad, const float* query_embed, const float* pos_embed,
                     const float* global_token, const float* global_token_pos, int32_t B, int32_t L, float* hs,
                     float* references, float* memory_local, float* memory_global, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (!ctx) return 1;
    if (!ctx->finalized) return fail(ctx, 1, "mesm_transformer: weights not finalized");
    if ((int)ctx->enc.size() != ctx->cfg.enc_layers || (int)ctx->dec.size() != ctx->cfg.dec_layers || !ctx->dec_norm.g ||
        !ctx->bb2.Wt || !ctx->ra1.Wt || !ctx->qs1.Wt || !ctx->rph1.Wt)
        return fail(ctx, 2, "mesm_transformer: transformer weights incomplete; missing: " + ctx->missing.substr(0, 600));
    if (!src || !pad || !query_embed || !pos_embed || !global_token || !global_token_pos || !workspace || B < 1 || L < 1 || L > 1023)
        return fail(ctx, 1, "mesm_transformer: bad argument");
    if (workspace_bytes < mesm_transformer_workspace_bytes(ctx, B, L)) return fail(ctx, 1, "mesm_transformer: workspace too small");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int L1 = L + 1, nq = ctx->cfg.num_queries;
    const size_t Re = (size_t)B * L1;
    Arena ar(workspace, workspace_bytes);
    float* E = ar.get<float>(Re * D); float* E2 = ar.get<float>(Re * D); float* posE = ar.get<float>(Re * D);
    uint8_t* padE = ar.get<uint8_t>(Re);
    EncBuffers eb;
    eb.QKV = ar.get<float>(Re * 3 * D); eb.AO = ar.get<float>(Re * D); eb.Y1 = ar.get<float>(Re * D); eb.H = ar.get<float>(Re * FF);
    if (attn_split_floats((long long)Re, L1)) eb.split = ar.get<float>(attn_split_floats((long long)Re, L1));
    DecBuffers d;
    dec_alloc(ar, d, B, nq, L1, ctx->cfg.dec_layers);
    build_enc_kernel<<<dim3(L1, B), D, 0, s>>>(src, pos_embed, pad, global_token, global_token_pos, B, L, E, posE, padE);
    g_stats.launches++;
    float* cur = E; float* nxt = E2;
    for (size_t l = 0; l < ctx->enc.size(); ++l) { CK(enc_layer(ctx->enc[l], cur, posE, padE, L1, B, eb, nxt, s)); std::swap(cur, nxt); }
    if (memory_local) CK(launch_copy_rows(cur, D, RowMap{L, L1, 1}, memory_local, D, identity_map(), (long long)B * L, s));
    if (memory_global) CK(launch_copy_rows(cur, D, RowMap{1, L1, 0}, memory_global, D, identity_map(), B, s));
    if (hs || references)
        CK(run_decoder(ctx, query_embed, cur, posE, pad, L, B, d, nullptr, nullptr, nullptr, nullptr, 0, hs, (long long)B * nq * D,
                       references, (long long)B * nq * 2, s));
    return 0;
}

/* Barrier watchdog records of the tcgen05 kernels: out16 = {attention[8], linear[8]}; [0] != 0 means a wait gave up
 * (then [1] = tag, [2] = block, [3] = thread, [4] = barrier address, [5] = parity).  Synchronises the device; clears. */
int mesm_debug_attn_trace(long long* out128) { return tc_read_attn_trace(out128); }

int mesm_debug_watchdog(unsigned long long* out16) {
    cudaDeviceSynchronize();
    tc_read_watchdog(out16);
    tc_read_watchdog_linear(out16 + 64);
    return 0;
}

/* Test hook: `iters` back-to-back self-attention launches (tcgen05 kernel when use_tc) on q/k/v = columns [0,256), [256,512),
 * [512,768) of qkv [B*L, 768]; returns the watchdog record (all zero = no stalled barrier wait). */
int mesm_debug_attention(const float* qkv, const uint8_t* k_pad, int32_t B, int32_t L, float* out, int32_t use_tc, int32_t iters,
                         unsigned long long* watchdog8, void* stream) {
    mesm_ctx* ctx = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    MhaRowsArgs a;
    a.q = qkv; a.ldq = 3 * D; a.k = qkv + D; a.ldk = 3 * D; a.v = qkv + 2 * D; a.ldv = 3 * D;
    a.k_pad = k_pad; a.q_pad = nullptr; a.out = out; a.ldo = D; a.B = B; a.Lq = L; a.Lk = L; a.b0 = 0; a.Btot = B;
    a.q_scale = kScale32;
    float* split = nullptr;
    if (use_tc == 2) {                 // key-split tensor-core path (more than 224 keys)
        const size_t nf = attn_split_floats((long long)B * L, L);
        if (!nf) return fail(ctx, 3, "key-split needs more than 224 keys");
        CK(cudaMalloc((void**)&split, nf * sizeof(float)));
        a.split_ws = split; a.split_rows = (long long)B * L;
    }
    for (int i = 0; i < iters; ++i) {
        if (use_tc == 3) { if (!attn_mma_eligible(a)) return fail(ctx, 3, "not eligible"); CK(launch_attn_mma(a, s)); }
        else if (use_tc == 2) { if (!attn_tc_split_eligible(a)) { cudaFree(split); return fail(ctx, 3, "not eligible"); } CK(launch_attn_tc_split(a, a.split_rows, s)); }
        else if (use_tc) { if (!attn_tc_eligible(a)) return fail(ctx, 3, "not eligible"); CK(launch_attn_tc(a, s)); }
        else { CK(launch_mha_rows(a, s, true)); }
    }
    CK(cudaStreamSynchronize(s));
    cudaFree(split);
    if (watchdog8) { unsigned long long tmp[64]; tc_read_watchdog(tmp); for (int i = 0; i < 8; ++i) watchdog8[i] = tmp[i]; }
    return 0;
}

/* Test hook: one fused linear through either kernel.  W [N,K] fp32 (out,in); every optional pointer may be NULL.
 * use_tc: 0 = fp32 SIMT kernel, 1 = tcgen05 kernel (error if the shape is not eligible). */
int mesm_debug_linear(const float* A, const float* Apos, const float* W, const float* bias, const float* residual,
                      const float* ln_g, const float* ln_b, const float* rowstat, const float* prelu, int32_t M, int32_t N,
                      int32_t K, int32_t lda, int32_t act, float out_scale, float* out, float* pre_ln, int32_t use_tc,
                      void* stream) {
    mesm_ctx* ctx = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    const int Kp = (K + 15) / 16 * 16, ldw = (N + 3) / 4 * 4;
    float* Wt = nullptr; void* Wp = nullptr; float* cs = nullptr;
    CK(cudaMalloc((void**)&Wt, (size_t)Kp * ldw * sizeof(float)));
    CK(cudaMalloc(&Wp, tc_packed_bytes(N, K)));
    CK(cudaMalloc((void**)&cs, (size_t)N * sizeof(float) + 16));
    CK(launch_transpose_pack(W, 0, N, K, Wt, ldw, Kp, s));
    CK(launch_pack_tc(W, 0, N, K, nullptr, Wp, s));
    CK(cudaStreamSynchronize(s));      // the tcgen05 kernels fetch weights ahead of their stream predecessors (dependent launch)
    LinearOp op = make_linear(M, N, K, A, lda, Wt, ldw, bias, out, N);
    op.Apos = Apos; op.Wp = Wp; op.act = act; op.prelu = prelu; op.out_scale = out_scale;
    op.residual = residual; op.ldr = N; op.ln_g = ln_g; op.ln_b = ln_b; op.pre_ln = pre_ln;
    if (rowstat) {      // colsum of the (unfolded) weights: the caller folds gamma into W beforehand
        CK(launch_colsum(Wt, Kp, ldw, N, cs, s));
        op.rowstat = rowstat; op.colsum = cs;
    }
    int rc = 0;
    uint16_t* planes = nullptr; void* wtm = nullptr;
    if (use_tc >= 2) {
        // TMA-fed kernel: A pre-split into bf16 hi/lo planes (use_tc 2: fp32 result; 3: result as planes, merged back into `out`;
        // 4: A as ONE exact fp16 plane - the caller passes fp16-representable values - against fp16 hi/lo weights)
        const bool f16 = use_tc == 4;
        const int ldp = (K + 7) / 8 * 8;
        CK(cudaMalloc((void**)&planes, ((size_t)2 * M * ldp + (size_t)2 * M * N) * sizeof(uint16_t)));
        uint16_t* a_hi = planes; uint16_t* a_lo = planes + (size_t)M * ldp; uint16_t* o_hi = a_lo + (size_t)M * ldp; uint16_t* o_lo = o_hi + (size_t)M * N;
        CK(cudaMemsetAsync(planes, 0, (size_t)2 * M * ldp * sizeof(uint16_t), s));
        if (Apos || (K & 1) || (lda & 1)) rc = fail(ctx, 3, "mesm_debug_linear: the TMA kernel takes no Apos / odd K");
        else if (f16) CK(launch_f32_to_f16(A, M, K, lda, a_hi, ldp, s));
        else CK(launch_split_planes(A, M, K, lda, a_hi, a_lo, ldp, s));
        wtm = tma_pack_weights(W, 0, N, K, nullptr, f16, s);
        CK(cudaStreamSynchronize(s));
        op.a_hi = a_hi; op.a_lo = f16 ? nullptr : a_lo; op.lda_p = ldp; op.Wtm = wtm;
        if (use_tc == 3) { op.out = nullptr; op.out_hi = o_hi; op.out_lo = o_lo; op.ldp = N; }
        if (rc == 0) {
            if (!wtm || !linear_tma_eligible(op)) rc = fail(ctx, 3, "mesm_debug_linear: shape not eligible for the TMA-fed tcgen05 kernel");
            else { cudaError_t e = launch_linear_tma(op, s); if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e)); }
        }
        if (rc == 0 && use_tc == 3) { cudaError_t e = launch_merge_planes(o_hi, o_lo, (long long)M * N, out, s); if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e)); }
    } else if (use_tc) {
        if (!linear_tc_eligible(op)) rc = fail(ctx, 3, "mesm_debug_linear: shape not eligible for the tcgen05 kernel");
        else { cudaError_t e = launch_linear_tc(op, s); if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e)); }
    } else {
        cudaError_t e = launch_linear_simt(op, s); if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e));
    }
    if (rc == 0 && getenv("MESM_DEBUG_LINEAR_ITERS")) {      // developer timing loop (tools/linear_bench.py): same launch repeated, CUDA events
        const int iters = atoi(getenv("MESM_DEBUG_LINEAR_ITERS"));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        for (int i = 0; i < iters; ++i) {
            if (use_tc >= 2) launch_linear_tma(op, s); else if (use_tc) launch_linear_tc(op, s); else launch_linear_simt(op, s);
        }
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[linear_bench] mode %d M %d N %d K %d: %.1f us per launch\n", use_tc, M, N, K, 1e3f * ms / iters);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess && rc == 0) rc = fail(ctx, (int)e, std::string("mesm_debug_linear: ") + cudaGetErrorString(e));
    cudaFree(Wt); cudaFree(Wp); cudaFree(cs); cudaFree(planes); tma_free_weights(wtm);
    return rc;
}

}  // extern "C"
