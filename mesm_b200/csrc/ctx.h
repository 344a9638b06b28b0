// Internal: context layout and the layer helpers shared by api.cu (fused forward) and api_modules.cu (sub-module
// entry points).  Not part of the C ABI.
#pragma once
#include "../../include/mesm_b200.h"
#include "kernels.h"

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace mesm {

struct Tensor {
    float* p = nullptr;
    std::vector<int64_t> shape;
    size_t n = 0;
};

struct PL {              // packed linear: Wt[Kp, ldw] (k-major), bias[N]
    const float* Wt = nullptr;
    int ldw = 0, K = 0, N = 0;
    const float* bias = nullptr;
    const float* colsum = nullptr;   // LayerNorm-folded layers only
    const void* Wp = nullptr;        // tcgen05 packed tiles (null for tiny layers)
    const void* Wtm = nullptr;       // the same weights as TMA-addressable bf16 hi/lo planes (linear_tma.cu; N % 256 == 0 only)
};

struct Norm { const float* g = nullptr; const float* b = nullptr; };

struct AttnFfn {         // T2V / recon / encoder layer
    PL q, kv, qk, v, out, l1, l2;
    const float* k_only_Wt = nullptr;      // K rows of in_proj (for the explicit pos_txt path)
    const float* in_w = nullptr;           // raw in_proj_weight [768,256] (recon back-projection uses row slices)
    const float* in_b = nullptr;
    const float* vT = nullptr;             // Wv^T packed [256,256] (recon)
    PL blk_qk, blk_v;                      // recon back-projections as ONE dense GEMM each: block-structured weights [8*256, 256] / [256, 8*256]
    Norm n1, n2;
    const float* prelu = nullptr;
    const void* ffn_w1 = nullptr;          // fused-FFN weight images (ffn_tc.cu); null when the shape is not 256 / 1024
    const void* ffn_w2 = nullptr;
    const void* ffn_maps = nullptr;        // host-side tensor maps of the two images
};

struct DecLayer {
    PL sa_qc, sa_qp, sa_kc, sa_kp, sa_v, sa_out;
    PL ca_qc, ca_qp, ca_kc, ca_kp, ca_v, ca_sine, ca_out, l1, l2;
    const float* sa_q_bias = nullptr;      // b(qcontent)+b(qpos)
    const float* sa_k_bias = nullptr;
    const float* ca_q_bias0 = nullptr;     // b(ca_qcontent)+b(ca_qpos)   (layer 0)
    const float* ca_k_bias0 = nullptr;     // b(ca_kcontent)+b(ca_kpos)   (layer 0)
    Norm n1, n2, n3;
    const float* prelu = nullptr;
};

}  // namespace mesm

using mesm::Tensor; using mesm::PL; using mesm::Norm; using mesm::AttnFfn; using mesm::DecLayer;

struct mesm_ctx {
    mesm_cfg cfg{};
    int device = 0;
    std::string err;
    std::string missing;               // state_dict keys absent / mis-shaped at the last finalize
    std::unordered_map<std::string, Tensor> w;
    std::vector<void*> owned;
    std::vector<void*> owned_host;      // malloc'ed host objects (tensor maps)
    std::vector<void*> owned_tma;       // TmaWeights objects (tma_free_weights)
    const void* vid0_f16 = nullptr;     // input_vid_proj.0 (LayerNorm-folded) as fp16 hi/lo planes: 16-bit stored features
    bool finalized = false;
    int chunk_pairs = 256;
    long long last_launches = 0;
    long long last_feature_bytes = 0;   // clip-feature bytes the last forward's first projection streamed (roofline of that stage)

    PL vid0, vid1, txt0, txt1;
    Norm vid1_ln, txt1_ln;
    std::vector<AttnFfn> enh, aln, rec, enc;
    std::vector<DecLayer> dec;
    Norm dec_norm;
    PL qs0, qs1, rph0, rph1, bb0, bb1, bb2, ra0, ra1;
    PL span0, span1, span2, cls, sal1, sal2, osp0, osp1;
    Norm osp0_ln, osp1_ln;
    const float *gtok = nullptr, *gpos = nullptr, *msent = nullptr, *qembed = nullptr;

    // pinned host tables of a forward (pulled by a kernel at its start).  A small ring of them lets the host enqueue a few
    // forwards ahead of the device: with a single table every forward had to wait for the previous one to START on the GPU,
    // so any host hiccup longer than one forward's GPU time stalled the device.
    static constexpr int kTabSlots = 3;
    int* h_tab[kTabSlots] = {nullptr, nullptr, nullptr};
    size_t h_tab_cap[kTabSlots] = {0, 0, 0};
    cudaEvent_t tab_event[kTabSlots] = {nullptr, nullptr, nullptr};
    bool tab_event_pending[kTabSlots] = {false, false, false};
    int tab_turn = 0;
    std::vector<int*> graph_tabs;      // pinned tables of forwards recorded into CUDA graphs (one per capture, kept until destroy)
};


namespace mesm {

int fail(mesm_ctx* c, int code, const std::string& msg);

#define CK(expr)                                                                                             \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return mesm::fail(ctx, (int)_e, std::string(#expr) + " failed at " + __FILE__ + ":" + std::to_string(__LINE__) + \
                                                ": " + cudaGetErrorString(_e));                              \
    } while (0)

// ---- workspace arena -----------------------------------------------------------------------------------------------
struct Arena {
    char* base; size_t off = 0, cap;
    Arena(void* b, size_t c) : base((char*)b), cap(c) {}
    template <class T> T* get(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : (T*)nullptr;
        off += n * sizeof(T);
        return p;
    }
    bool fits() const { return base == nullptr || off <= cap; }
};

struct Lin {    // thin builder around LinearOp
    LinearOp op;
    Lin(int M, const PL& w, const float* A, int lda, float* out, int ldo) {
        op = make_linear(M, w.N, w.K, A, lda, w.Wt, w.ldw, w.bias, out, ldo);
        op.Wp = w.Wp; op.Wtm = w.Wtm;
    }
    // A / result as pre-split 16-bit planes (linear_tma.cu); lo == nullptr: one exact fp16 plane
    Lin& aplanes(const uint16_t* hi, const uint16_t* lo, int ld) { op.a_hi = hi; op.a_lo = lo; op.lda_p = ld; return *this; }
    Lin& oplanes(uint16_t* hi, uint16_t* lo, int ld) { op.out_hi = hi; op.out_lo = lo; op.ldp = ld; return *this; }
    Lin& wtm(const void* w) { op.Wtm = w; return *this; }
    Lin& res_planes(const uint16_t* hi, const uint16_t* lo) { op.res_hi = hi; op.res_lo = lo; return *this; }
    Lin& bias(const float* b) { op.bias = b; return *this; }
    Lin& amap(RowMap m) { op.amap = m; return *this; }
    Lin& omap(RowMap m) { op.omap = m; return *this; }
    Lin& apos(const float* p) { op.Apos = p; return *this; }
    Lin& second(const float* A2, int lda2, const PL& w2, RowMap m = identity_map()) {
        op.A2 = A2; op.lda2 = lda2; op.K2 = w2.K; op.Wt2 = w2.Wt; op.Wp2 = w2.Wp; op.a2map = m; return *this;
    }
    Lin& act(int a, const float* slope = nullptr) { op.act = a; op.prelu = slope; return *this; }
    Lin& scale(float s) { op.out_scale = s; return *this; }
    Lin& fold(const float* rowstat, const float* colsum) { op.rowstat = rowstat; op.colsum = colsum; return *this; }
    Lin& fold_fused(const float* colsum) { op.rowstat = nullptr; op.fuse_rowstat = 1; op.colsum = colsum; return *this; }
    Lin& res(const float* r, int ldr, RowMap m = identity_map()) { op.residual = r; op.ldr = ldr; op.rmap = m; return *this; }
    Lin& ln(const Norm& n) { op.ln_g = n.g; op.ln_b = n.b; return *this; }
    Lin& pre_ln(float* p) { op.pre_ln = p; return *this; }
    Lin& ln_stats(float* p) { op.ln_stats = p; return *this; }
    Lin& batch(int n, long long sA, long long sW, long long sBias, long long sOut) {
        op.nbatch = n; op.bsA = sA; op.bsW = sW; op.bsBias = sBias; op.bsOut = sOut; return *this;
    }
    cudaError_t run(cudaStream_t s) { return launch_linear(op, s); }
};


constexpr float kScale32 = 0.17677669529663687f;   // 32^-0.5
constexpr float kScale64 = 0.125f;                 // 64^-0.5

// Pre-split activations (bf16 hi / lo planes, [rows, 256], pitch 256) around one layer call: the layer input as planes (A operand of
// its Q / QK / V projections on linear_tma.cu), where to store its output as planes (same row mapping as the fp32 output), and the two
// plane scratch buffers every layer needs (attention output, LayerNorm-1 output).  Empty members fall back to the fp32 operands.
struct Planes { uint16_t* hi = nullptr; uint16_t* lo = nullptr; explicit operator bool() const { return hi != nullptr; } };
struct PlaneIO {
    Planes in, out, ao, y1;
    bool planes_only = false;   // request: store the layer's result as planes only (no fp32 copy) - honoured on the plane path
    bool wrote_out = false;     // set by the layer: `out` now holds its result
    bool wrote_fp32 = true;     // set by the layer: the fp32 output buffer holds its result
};

struct T2VBuffers { float *KV, *Q, *AO, *X1, *Y1, *H; };
struct EncBuffers { float *QKV, *AO, *Y1, *H; float* split = nullptr; /* key-split attention scratch (attn_split_floats), Lv + 1 > 224 only */ };
struct DecBuffers {
    float *tgtA, *tgtB, *ref, *refs, *sine, *sine_s, *h1, *h2, *qpos, *ptrans, *anc, *qsa, *ksa, *vsa, *ao, *t1, *qca,
        *sinep, *t2, *hff, *d2, *hs, *Kc, *Kp, *Vd, *tmpref;
};

cudaError_t t2v_layer(const AttnFfn& L, const float* txt, RowMap tmap, const float* pos_txt, int Lk, const float* vid,
                      const float* pos_vid, int Lq, int Bc, int b0, int Btot, const uint8_t* q_pad, const uint8_t* k_pad,
                      const T2VBuffers& t, float* out, int ldo, RowMap omap, cudaStream_t s, bool reuse_q = false,
                      const int* cu = nullptr, int Rv_packed = 0, int q_pad_ld = 0, const float* posW = nullptr,
                      const int* t_pos = nullptr, PlaneIO* pio = nullptr);
// posW / t_pos: the position term of a projection as a gathered residual - row r of the GEMM adds posW[t_pos[r]], where
// posW = (position table) . W^T was computed once per distinct (clip count, clip index); replaces the (x + pos) K-sweep
// cu != nullptr (device, first pair of the chunk): packed variable-length rows, see pair_rows() in common.cuh
cudaError_t enc_layer(const AttnFfn& L, const float* src, const float* pos, const uint8_t* pad, int L1, int Bc,
                      const EncBuffers& t, float* out, cudaStream_t s, const int* cu = nullptr, int R_packed = 0,
                      const float* posW = nullptr, const int* t_pos = nullptr, PlaneIO* pio = nullptr);
size_t dec_alloc(Arena& ar, DecBuffers& d, int Bc, int nq, int L1, int nl);
cudaError_t launch_colsum(const float* Wt, int Kp, int ldw, int N, float* out, cudaStream_t s);
cudaError_t launch_transpose_pack(const float* W, int row0, int nrows, int K, float* Wt, int ldw, int Kp, cudaStream_t s);
cudaError_t run_decoder(const mesm_ctx* c, const float* qembed, const float* E, const float* posE, const uint8_t* padV, int Lv, int Bc,
                        const DecBuffers& d, float* logits_out, float* spans_out, float* aux_logits, float* aux_spans,
                        long long aux_layer_stride, float* hs_out, long long hs_layer_stride, float* refs_out,
                        long long refs_layer_stride, cudaStream_t s, const int* cu = nullptr, int Re_packed = 0,
                        const float* const* posWkp = nullptr, const int* t_posE = nullptr, Planes Ep = Planes());

}  // namespace mesm
