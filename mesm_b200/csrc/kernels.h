// Internal launcher prototypes (definitions in attention.cu / elementwise.cu / linear_*.cu).
#pragma once
#include "common.cuh"

namespace mesm {

struct MhaRowsArgs {
    const float* q; int ldq;
    const float* k; int ldk;
    const float* v; int ldv;
    const uint8_t* k_pad;
    const uint8_t* q_pad;
    float* out; int ldo;
    int B, Lq, Lk;             // uniform layout: Lq / Lk rows per pair; packed layout: the maxima over the pairs
    int b0, Btot;
    float q_scale;
    // packed (variable-length) layouts, see pair_rows() in common.cuh; null = uniform.  k_pad is indexed like the key rows
    // when k_cu is set, else [Btot, Lk] by global pair; q_pad (quirk partner) is always [Btot, q_pad_ld] by global pair.
    const int* q_cu = nullptr; int q_enc = 0;
    const int* k_cu = nullptr; int k_enc = 0;
    int q_pad_ld = 0;          // 0 = Lq
    // optional: the result as bf16 hi / lo planes [rows, 256] (the A operand of the output projection on linear_tma.cu); `out` may then be null
    uint16_t* out_hi = nullptr; uint16_t* out_lo = nullptr;
    // Key-split pass (tcgen05 kernel, self-attention with more keys than one 224-key tile, e.g. the shipped max_video_l = 600):
    // only keys [k_begin, k_begin + k_count) of every pair are visited, the output is the chunk's own softmax-normalised result and
    // split_stats[(row * 8 + head) * 2 + {0, 1}] = (row maximum, row sum of exp) lets launch_attn_combine merge the chunks exactly.
    int k_begin = 0, k_count = 0;
    float* split_stats = nullptr;
    // caller-provided scratch for the key-split path: nchunks x (rows x 256 + rows x 16) floats (see attn_split_floats)
    float* split_ws = nullptr;
    long long split_rows = 0;   // total query rows of the launch (uniform: B * Lq; packed: the packed row count)
};
bool attn_tc_split_eligible(const MhaRowsArgs& a);
cudaError_t launch_attn_tc_split(const MhaRowsArgs& a, long long rows, cudaStream_t s);
size_t attn_split_floats(long long rows, int Lk);         // 0 when one key tile suffices
cudaError_t launch_mha_rows(const MhaRowsArgs& a, cudaStream_t s, bool force_simt = false);      // dispatcher: tcgen05 kernel when eligible
bool attn_tc_eligible(const MhaRowsArgs& a);
bool attn_mma_eligible(const MhaRowsArgs& a);                       // attn_mma.cu: warp-level mma.sync kernel, any Lk whose operands fit shared memory
cudaError_t launch_attn_mma(const MhaRowsArgs& a, cudaStream_t s);
cudaError_t launch_attn_tc(const MhaRowsArgs& a, cudaStream_t s);

struct MhaSmallArgs {
    const float* q; int ldq;  const float* q2; int ldq2;
    const float* k; int ldk;  const float* k2; int ldk2;
    const float* v; int ldv;
    const uint8_t* k_pad;
    float* out; int ldo;
    float* attn_w;
    int B, L, S, nheads, hq, hv;
    float scale;
    int q_bs, q_is;            // row of query (b,i) = b*q_bs + i*q_is  (same for out; attn_w is always [B,L,S])
    int k_bs, k_is, k_off;     // row of key/value (b,j) = b*k_bs + j*k_is + k_off
    // packed keys: pair b owns the S_b = k_cu[b+1]-k_cu[b] key rows (k_cu[b]-k_cu[0]) + b*k_enc + k_off + j; k_pad is then
    // indexed (k_cu[b]-k_cu[0]) + j.  S stays the maximum over the pairs (shared-memory sizing).
    const int* k_cu = nullptr; int k_enc = 0;
    const int* k2_table = nullptr;   // k2 row of key row r is k2_table[r] (position-term table) instead of r
    int causal = 0;                  // 1: key j is masked for query i when j > i (CLIP text tower, model/text_encoder.py:321-327)
};
cudaError_t launch_mha_small(const MhaSmallArgs& a, cudaStream_t s);
bool dec_cross_mma_eligible(const MhaSmallArgs& a);                    // attn_mma.cu: decoder cross-attention on mma.sync
cudaError_t launch_dec_cross_mma(const MhaSmallArgs& a, cudaStream_t s);

struct ReconPoolArgs {
    const float* x; int ldx;
    const uint8_t* vmask;
    const float* qk;
    float* pooled;
    const int* pair_group;
    const int* pair_slot;
    const int* group_start;
    const int* group_len;
    int B, Lv, qvh, max_keys;
    int b0, Btot;
    const int* x_start = nullptr;   // packed clip rows: pair p's clip i is row x_start[p] + i of x (null: (p - b0) * Lv + i)
};
cudaError_t launch_recon_pool(const ReconPoolArgs& a, cudaStream_t s);

struct PosArgs {
    const uint8_t* vmask; int B, Lv;
    const float* gtok; const float* gpos;
    float* posV; float* posE; float* encbuf; uint8_t* padV; uint8_t* padE;
    const int* cu = nullptr;        // packed layout (pair_rows): clip rows at cu[b]-cu[0], encoder rows at cu[b]-cu[0]+b
    uint16_t* enc_hi = nullptr; uint16_t* enc_lo = nullptr;     // optional planes of encbuf: the global-token row is written there too
};
cudaError_t launch_pos_embed(const PosArgs& a, cudaStream_t s);

cudaError_t launch_text_prep(const float* x, int R, int Dt, float* y, uint8_t* mask, float* rowstat, cudaStream_t s);
cudaError_t launch_row_stats(const float* x, long long R, int Dv, int ldx, float* rowstat, cudaStream_t s, const int* table = nullptr);
cudaError_t launch_invert_mask(const uint8_t* in, uint8_t* out, long long n, cudaStream_t s);
cudaError_t launch_expand_mask(const uint8_t* wmask, int B, int Lt, uint8_t* emask, uint8_t* epad, uint8_t* wpad, cudaStream_t s);
cudaError_t launch_expand_mask_from_epad(const uint8_t* epad, int B, int Lt, uint8_t* wpad, cudaStream_t s);
cudaError_t launch_gather_blocks(const float* src, float* dst, const int64_t* idx, int B, long long block_elems,
                                 const uint8_t* msrc, uint8_t* mdst, int mlen, cudaStream_t s);
cudaError_t launch_copy_rows(const float* src, int lds, RowMap imap, float* dst, int ldd, RowMap omap, long long R, cudaStream_t s);
cudaError_t launch_broadcast_row(const float* vec, float* dst, long long R, cudaStream_t s);
cudaError_t launch_l2norm_rows(const float* x, long long R, float* out1, float* out2, int ld2, RowMap map2, cudaStream_t s);
cudaError_t launch_saliency(const float* p1, RowMap map1, const float* p2, int B, int Lv, float* out, cudaStream_t s);
// packed encoder layout: p1 row of (b, i) = cu[b]-cu[0] + b + 1 + i for i < cu[b+1]-cu[b]; out [B, Lv] (0 at the pad clips)
cudaError_t launch_saliency_packed(const float* p1, const int* cu, const float* p2, int B, int Lv, float* out, cudaStream_t s);
// packed-layout tables: t_pad[cu[b] + i] = src(b) * Lv + i (packed clip row -> row of the zero-padded [B, Lv] layout);
// src(b) = b, or the first pair of b's video group when pair_group / group_start are given
cudaError_t launch_pack_table(const int* cu, int B, int Lv, int* t_pad, cudaStream_t s, const int* pair_group = nullptr,
                              const int* group_start = nullptr);
cudaError_t launch_video_tables(const int* cu, const int* vcu, const int* pair_group, const int* group_start, int B, int G, int Lv,
                                int* t_vin, int* t_p2v, cudaStream_t s);
// rows [r0, r0 + R) of the (gathered) fp16 input -> compact [R, ldo] fp16 rows + their LayerNorm statistics (elementwise.cu)
cudaError_t launch_repack_f16_rows(const uint16_t* x, const int* table, long long r0, long long R, int Dv, uint16_t* out, int ldo,
                                   float* rowstat, cudaStream_t s);
// per chunk: t_c2e[r] = encoder-buffer row of packed clip row r; t_g[b] = encoder-buffer row of pair b's global token
cudaError_t launch_chunk_tables(const int* cu, int Bc, int* t_c2e, int* t_g, cudaStream_t s, const int* len_off = nullptr,
                                int* t_posV = nullptr, int* t_posE = nullptr);
// Position table (packed layout): PositionEmbeddingSine depends only on (clip count n, clip index i), so the batch needs
// one row per distinct (n, i): row 0 = global_rep_pos, rows dl_off[j] + i (i < dl_len[j]) = the sine embedding of clip i
// in a video of dl_len[j] clips.  chunk tables t_posV / t_posE (len_off[n] = first row of length n) index it per row.
cudaError_t launch_pos_table(const int* dl_len, const int* dl_off, int nd, const float* gpos, float* PT, cudaStream_t s);
// rows r < R of x [R, width] with mask[r] == 0 are set to zero
cudaError_t launch_zero_masked_rows(float* x, const uint8_t* mask, long long R, int width, cudaStream_t s);
cudaError_t launch_dec_init_ref(const float* qe, int B, int nq, float* ref, cudaStream_t s);
cudaError_t launch_dec_sine(const float* ref, long long R, const float* pos_trans, const float* anchor, float* sine,
                            float* scaled, cudaStream_t s);
cudaError_t launch_ref_update(const float* delta, int ldd, const float* ref, long long rows, float* out, cudaStream_t s);
cudaError_t launch_layernorm_rows(const float* x, long long R, const float* g, const float* b, float* out, cudaStream_t s);
cudaError_t launch_group_len(const uint8_t* vmask, int Lv, const int* group_start, int G, int* group_len, cudaStream_t s);
cudaError_t launch_masked_mean_norm(const float* x, const uint8_t* mask, int B, int L, float* out, int ldo, int transposed, cudaStream_t s);
// Fused FFN block (ffn_tc.cu): out = LN2(R + W2 . PReLU(W1 . X + b1) + b2), d_model 256, hidden 1024, CTA pairs.
struct FfnArgs {
    const float* X; int ldx;                 // FFN input rows [M, 256]
    const float* R; int ldr;                 // residual rows [M, 256]
    float* out; int ldo; RowMap omap;
    int M;
    const void* W1f; const void* W2f;        // launch_pack_ffn images of linear1 / linear2
    const void* maps;                        // ffn_make_maps(W1f, W2f)
    const float* b1; const float* b2; const float* ln_g; const float* ln_b; const float* prelu;
    // pre-split planes (optional): X as bf16 hi / lo [M, 256] (then X may be null: no conversion in the kernel, TMA loads), and the
    // result also as planes with the row mapping of `out` (needs ldo == 256; `out` may be null)
    const uint16_t* x_hi = nullptr; const uint16_t* x_lo = nullptr;
    uint16_t* out_hi = nullptr; uint16_t* out_lo = nullptr;
    // LayerNorm-1 fused into the prologue (optional): X holds the PRE-LayerNorm rows, ln1_stats their (mean, rstd) [M,2] as written by the
    // output projection (LinearOp::ln_stats); the kernel normalises while it converts X, so LN1's output never exists in HBM.
    // res_ln1 = 1 (encoder layers): the residual is LN1(R) as well (R = the same pre-LN rows), recomputed in the epilogue.
    const float* ln1_g = nullptr; const float* ln1_b = nullptr; const float* ln1_stats = nullptr; int res_ln1 = 0;
};
size_t ffn_packed_bytes();
cudaError_t launch_pack_ffn(const float* W1, const float* W2, void* W1f, void* W2f, cudaStream_t s);
void* ffn_make_maps(const void* W1f, const void* W2f);      // host object (free() it); null when TMA descriptors are unavailable
bool ffn_fused_eligible(const FfnArgs& a);
cudaError_t launch_ffn_fused(const FfnArgs& a, cudaStream_t s);
cudaError_t launch_fill(float* p, long long n, float v, cudaStream_t s);
cudaError_t launch_pull_ints(const int* pinned_host_src, int* dst, long long n, cudaStream_t s);

}  // namespace mesm
