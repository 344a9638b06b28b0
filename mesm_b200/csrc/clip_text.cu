// CLIP text tower: CLIPTextEncoder.forward (model/text_encoder.py:240-354) - token + positional embedding,
// `layers` pre-norm residual blocks (ln_1 -> nn.MultiheadAttention with the causal mask of :321-327 -> + x; ln_2 -> c_fc ->
// QuickGELU -> c_proj -> + x; :165-186), ln_final, and the pooled output (eot row @ text_projection, :346-348).  The reference
// runs it in fp16 on the GPU; here every product goes through the same fp32-in / fp32-out bf16x3 linears as the rest of the path
// (more accurate than the reference's own fp16 run; the test compares against the reference module in fp32 and in fp16).
// Built from the path's existing kernels: fused linear (tcgen05 when the shape is eligible), mha_small_kernel (head_dim 64,
// causal), plus four row-wise kernels below.  A front-end: correctness first, not on the benchmark path.
#include "ctx.h"

#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

using namespace mesm;

struct mesm_clip {
    int width = 0, heads = 0, layers = 0, ctx = 0, vocab = 0, embed = 0, device = 0;
    std::string err, missing;
    std::unordered_map<std::string, Tensor> w;
    std::vector<void*> owned;
    bool finalized = false;
    struct Block { PL qkv, out, fc, proj; Norm ln1, ln2; };
    std::vector<Block> blocks;
    Norm ln_final;
    PL text_proj;
    const float *tok = nullptr, *pos = nullptr;
};

namespace {
thread_local std::string g_clip_create_error;
int cfail(mesm_clip* c, int code, const std::string& msg) { if (c) c->err = msg; else g_clip_create_error = msg; return code ? code : 1; }
#define CCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cfail(ctx, (int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)

__global__ void clip_embed_kernel(const int64_t* __restrict__ text, const float* __restrict__ tok, const float* __restrict__ pos, int ctx_len, int W,
                                  int vocab, float* __restrict__ x) {
    const long long r = blockIdx.x;                       // row = b * ctx_len + t
    long long id = text[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int t = (int)(r % ctx_len);
    for (int c = threadIdx.x; c < W; c += blockDim.x) x[r * W + c] = tok[id * W + c] + pos[(long long)t * W + c];
}
// LayerNorm over rows of any width (eps 1e-5, two passes), one warp per row
__global__ void ln_rows_any_kernel(const float* __restrict__ x, long long R, int W, const float* __restrict__ g, const float* __restrict__ b,
                                   float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + r * W;
    float s = 0.f;
    for (int c = lane; c < W; c += 32) s += xr[c];
    const float mu = warp_sum(s) / W;
    float q = 0.f;
    for (int c = lane; c < W; c += 32) { const float d = xr[c] - mu; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) / W + 1e-5f);
    for (int c = lane; c < W; c += 32) out[r * W + c] = (xr[c] - mu) * rs * g[c] + b[c];
}
__global__ void quick_gelu_kernel(float* __restrict__ x, long long n) {           // x * sigmoid(1.702 x), text_encoder.py:160-162
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i]; x[i] = v / (1.f + __expf(-1.702f * v)); }
}
// eot row of every sequence = position of the largest token id (text.argmax(-1), first maximum like torch)
__global__ void clip_eot_gather_kernel(const int64_t* __restrict__ text, const float* __restrict__ x, int ctx_len, int W, float* __restrict__ out) {
    const int b = blockIdx.x;
    __shared__ int s_pos;
    if (threadIdx.x == 0) {
        int best = 0; long long bv = text[(long long)b * ctx_len];
        for (int t = 1; t < ctx_len; ++t) { const long long v = text[(long long)b * ctx_len + t]; if (v > bv) { bv = v; best = t; } }
        s_pos = best;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) out[(long long)b * W + c] = x[((long long)b * ctx_len + s_pos) * W + c];
}

struct ClipPacker {
    mesm_clip* c; cudaStream_t s; cudaError_t cerr = cudaSuccess;
    float* alloc(size_t n) { void* p = nullptr; cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(float)); if (e != cudaSuccess) { cerr = e; return nullptr; } c->owned.push_back(p); return (float*)p; }
    const Tensor* get(const std::string& k, std::initializer_list<int64_t> shape) {
        auto it = c->w.find(k);
        if (it == c->w.end()) { c->missing += k + " "; return nullptr; }
        if (std::vector<int64_t>(shape) != it->second.shape) { c->missing += k + "(shape) "; return nullptr; }
        return &it->second;
    }
    Norm norm(const std::string& p, int64_t n) { Norm r; const Tensor* g = get(p + ".weight", {n}); const Tensor* b = get(p + ".bias", {n}); r.g = g ? g->p : nullptr; r.b = b ? b->p : nullptr; return r; }
    // W [N, K] (nn.Linear) or, with transposed = true, W [K, N] (text_projection: x @ W)
    PL lin(const std::string& wk, const std::string& bk, int64_t N, int64_t K, bool transposed = false) {
        PL r;
        const Tensor* W = transposed ? get(wk, {K, N}) : get(wk, {N, K});
        const Tensor* b = bk.empty() ? nullptr : get(bk, {N});
        if (!W || (!bk.empty() && !b)) return r;
        const int Kp = (int)((K + 15) / 16 * 16), ldw = (int)((N + 3) / 4 * 4);
        float* Wt = alloc((size_t)Kp * ldw);
        if (!Wt) return r;
        if (transposed) {               // already [K, N]: copy rows into the padded k-major layout
            cudaMemsetAsync(Wt, 0, (size_t)Kp * ldw * sizeof(float), s);
            cudaMemcpy2DAsync(Wt, (size_t)ldw * sizeof(float), W->p, (size_t)N * sizeof(float), (size_t)N * sizeof(float), (size_t)K, cudaMemcpyDeviceToDevice, s);
        } else {
            launch_transpose_pack(W->p, 0, (int)N, (int)K, Wt, ldw, Kp, s);
            if (N >= 64) {
                void* wp = nullptr;
                if (cudaMalloc(&wp, tc_packed_bytes((int)N, (int)K)) == cudaSuccess) { c->owned.push_back(wp); launch_pack_tc(W->p, 0, (int)N, (int)K, nullptr, wp, s); r.Wp = wp; }
            }
        }
        r.Wt = Wt; r.ldw = ldw; r.K = (int)K; r.N = (int)N; r.bias = b ? b->p : nullptr;
        return r;
    }
};
}  // namespace

extern "C" {

const char* mesm_clip_last_error(const mesm_clip* c) { return c ? c->err.c_str() : g_clip_create_error.c_str(); }

mesm_clip* mesm_clip_create(int32_t width, int32_t heads, int32_t layers, int32_t context_length, int32_t vocab_size, int32_t embed_dim, int32_t device) {
    if (width < 64 || heads < 1 || width % heads || (width / heads) % 4 || layers < 1 || context_length < 1 || vocab_size < 1 || embed_dim < 1 || (width & 3)) {
        g_clip_create_error = "mesm_clip_create: bad shape (width % heads == 0, head_dim % 4 == 0, width % 4 == 0)"; return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_clip_create_error = "no CUDA device (mesm_b200 has no CPU fallback)"; return nullptr; }
    if (device < 0 || device >= ndev) { g_clip_create_error = "bad device index"; return nullptr; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10) { g_clip_create_error = "mesm_b200 is built for sm_100a only"; return nullptr; }
    mesm_clip* c = new mesm_clip();
    c->width = width; c->heads = heads; c->layers = layers; c->ctx = context_length; c->vocab = vocab_size; c->embed = embed_dim; c->device = device;
    return c;
}

void mesm_clip_destroy(mesm_clip* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (void* p : c->owned) cudaFree(p);
    for (auto& kv : c->w) cudaFree(kv.second.p);
    delete c;
}

int mesm_clip_load_weight(mesm_clip* ctx, const char* key, const float* data, const int64_t* shape, int ndim, int is_device, void* stream) {
    if (!ctx || !key || !data || ndim < 0 || ndim > 4) return cfail(ctx, 1, "mesm_clip_load_weight: bad argument");
    CCK(cudaSetDevice(ctx->device));
    size_t n = 1;
    std::vector<int64_t> shp;
    for (int i = 0; i < ndim; ++i) { n *= (size_t)shape[i]; shp.push_back(shape[i]); }
    Tensor& t = ctx->w[key];
    if (t.p && t.n != n) { cudaFree(t.p); t.p = nullptr; }
    if (!t.p) CCK(cudaMalloc((void**)&t.p, std::max<size_t>(n, 4) * sizeof(float)));
    t.n = n; t.shape = shp;
    CCK(cudaMemcpyAsync(t.p, data, n * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    ctx->finalized = false;
    return 0;
}

int mesm_clip_finalize(mesm_clip* ctx, void* stream) {
    if (!ctx) return 1;
    CCK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    CCK(cudaStreamSynchronize(s));
    for (void* p : ctx->owned) cudaFree(p);
    ctx->owned.clear(); ctx->blocks.clear(); ctx->missing.clear();
    ClipPacker P{ctx, s};
    const int64_t W = ctx->width;
    for (int l = 0; l < ctx->layers; ++l) {
        const std::string p = "transformer.resblocks." + std::to_string(l) + ".";
        mesm_clip::Block b;
        b.ln1 = P.norm(p + "ln_1", W); b.ln2 = P.norm(p + "ln_2", W);
        b.qkv = P.lin(p + "attn.in_proj_weight", p + "attn.in_proj_bias", 3 * W, W);
        b.out = P.lin(p + "attn.out_proj.weight", p + "attn.out_proj.bias", W, W);
        b.fc = P.lin(p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", 4 * W, W);
        b.proj = P.lin(p + "mlp.c_proj.weight", p + "mlp.c_proj.bias", W, 4 * W);
        ctx->blocks.push_back(b);
    }
    ctx->ln_final = P.norm("ln_final", W);
    ctx->text_proj = P.lin("text_projection", "", ctx->embed, W, true);
    const Tensor* tok = P.get("token_embedding.weight", {ctx->vocab, W});
    const Tensor* pos = P.get("positional_embedding", {ctx->ctx, W});
    ctx->tok = tok ? tok->p : nullptr; ctx->pos = pos ? pos->p : nullptr;
    if (P.cerr != cudaSuccess) return cfail(ctx, (int)P.cerr, std::string("weight packing: ") + cudaGetErrorString(P.cerr));
    CCK(cudaGetLastError());
    CCK(cudaStreamSynchronize(s));
    if (!ctx->missing.empty()) return cfail(ctx, 2, "mesm_clip_finalize: state_dict incomplete or mis-shaped; missing: " + ctx->missing.substr(0, 600));
    ctx->finalized = true;
    return 0;
}

size_t mesm_clip_workspace_bytes(const mesm_clip* c, int32_t B) {
    if (!c || B < 1) return 0;
    const size_t R = (size_t)B * c->ctx, W = c->width;
    return (R * W * 4 + R * 3 * W + R * 4 * W + (size_t)B * W) * sizeof(float) + 8192;      // x, y, ao, tmp | qkv | hidden | eot rows
}

int mesm_clip_forward(mesm_clip* ctx, const int64_t* text, int32_t B, float* last_hidden_state, float* pooler_output, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!ctx) return 1;
    if (!text || !last_hidden_state || !workspace || B < 1) return cfail(ctx, 1, "mesm_clip_forward: bad argument");
    if (!ctx->finalized) return cfail(ctx, 1, "mesm_clip_forward: weights not finalized");
    if (workspace_bytes < mesm_clip_workspace_bytes(ctx, B)) return cfail(ctx, 1, "mesm_clip_forward: workspace too small");
    CCK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int W = ctx->width, L = ctx->ctx, H = ctx->heads, hd = W / H;
    const long long R = (long long)B * L;
    Arena ar(workspace, workspace_bytes);
    float* x = ar.get<float>((size_t)R * W); float* y = ar.get<float>((size_t)R * W); float* ao = ar.get<float>((size_t)R * W);
    float* x2 = ar.get<float>((size_t)R * W); float* qkv = ar.get<float>((size_t)R * 3 * W); float* hid = ar.get<float>((size_t)R * 4 * W);
    float* eot = ar.get<float>((size_t)B * W);
    clip_embed_kernel<<<(unsigned)R, 128, 0, s>>>(text, ctx->tok, ctx->pos, L, W, ctx->vocab, x);
    g_stats.launches++;
    for (const mesm_clip::Block& b : ctx->blocks) {
        ln_rows_any_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(x, R, W, b.ln1.g, b.ln1.b, y);
        CCK(Lin((int)R, b.qkv, y, W, qkv, 3 * W).run(s));
        MhaSmallArgs a;
        a.q = qkv; a.ldq = 3 * W; a.q2 = nullptr; a.ldq2 = 0; a.k = qkv + W; a.ldk = 3 * W; a.k2 = nullptr; a.ldk2 = 0; a.v = qkv + 2 * W; a.ldv = 3 * W;
        a.k_pad = nullptr; a.out = ao; a.ldo = W; a.attn_w = nullptr; a.B = B; a.L = L; a.S = L; a.nheads = H; a.hq = hd; a.hv = hd;
        a.scale = 1.f / sqrtf((float)hd); a.q_bs = L; a.q_is = 1; a.k_bs = L; a.k_is = 1; a.k_off = 0; a.causal = 1;
        CCK(launch_mha_small(a, s));
        CCK(Lin((int)R, b.out, ao, W, x2, W).res(x, W).run(s));                       // x + attn(ln_1(x))
        ln_rows_any_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(x2, R, W, b.ln2.g, b.ln2.b, y);
        CCK(Lin((int)R, b.fc, y, W, hid, 4 * W).run(s));
        quick_gelu_kernel<<<(unsigned)((R * 4 * W + 255) / 256), 256, 0, s>>>(hid, R * 4 * W);
        CCK(Lin((int)R, b.proj, hid, 4 * W, x, W).res(x2, W).run(s));                 // x + mlp(ln_2(x))
        g_stats.launches += 3;
    }
    ln_rows_any_kernel<<<(unsigned)((R + 7) / 8), 256, 0, s>>>(x, R, W, ctx->ln_final.g, ctx->ln_final.b, last_hidden_state);
    g_stats.launches++;
    if (pooler_output) {
        clip_eot_gather_kernel<<<B, 128, 0, s>>>(text, last_hidden_state, L, W, eot);
        g_stats.launches++;
        CCK(Lin(B, ctx->text_proj, eot, W, pooler_output, ctx->embed).run(s));
    }
    CCK(cudaGetLastError());
    return 0;
}

}  // extern "C"
