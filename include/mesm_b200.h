/*
 * mesm_b200 — C ABI of the B200-native (sm_100a) MESM per-pair inference path.
 *
 * The reference (lntzm/MESM) is pure Python/PyTorch and has no FFI surface; the drop-in boundary is therefore the set
 * of Python call sites listed beside each entry point below (file:line in the reference repo).  `mesm_b200/` binds
 * these symbols with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch types.  Every `dev` pointer is device memory owned by the CALLER
 *     (PyTorch); the library allocates device memory only inside mesm_create / mesm_load_weight (packed weights).
 *   - All calls are stream-ordered on the `stream` argument (a cudaStream_t passed as void*), never synchronise the
 *     device and never fall back to the CPU.  Non-zero return = error; text via mesm_last_error().
 *   - float tensors are fp32 row-major contiguous; masks are uint8 (1 = valid), as torch.bool storage.
 *   - One ctx per device; a ctx is not thread-safe.
 */
#ifndef MESM_B200_H
#define MESM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MESM_ABI_VERSION 3

typedef struct mesm_ctx mesm_ctx;

/* Hyper-parameters that shape the module tree — the argparse/JSON keys of utils/config.py:26-163 that
 * runner.py:255-298 (build_model) consumes. */
typedef struct mesm_cfg {
    int32_t v_feat_dim;        /* incl. the +2 tef columns (utils/config.py:242-243) */
    int32_t t_feat_dim;
    int32_t hidden_dim;        /* must be 256 */
    int32_t nheads;            /* must be 8   */
    int32_t dim_feedforward;   /* must be 1024 */
    int32_t num_queries;       /* <= 32 (10 in every shipped config) */
    int32_t num_recfw_layers;  /* enhance encoder (FW-MESM) */
    int32_t t2v_layers;        /* aligner */
    int32_t enc_layers;
    int32_t dec_layers;
    int32_t num_recss_layers;  /* SS-MESM reconstructor */
    int32_t n_input_proj;      /* must be 2 */
    int32_t rec_fw;            /* bool */
    int32_t rec_ss;            /* bool */
    int32_t share_mlp;         /* 0 -> T2VEncoder_TwoMLP key set for the enhance encoder (runner.py:190-210) */
    int32_t qvh_grouping;      /* 1 -> dataset_name == "qvhighlights" branch of model/model.py:190-195, else :185-189 */
    int32_t max_words_l;
    int32_t max_video_l;
} mesm_cfg;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
/* replaces runner.build_model (runner.py:255-298) + model.to(device) */
mesm_ctx*   mesm_create(const mesm_cfg* cfg, int device);
void        mesm_destroy(mesm_ctx* ctx);
const char* mesm_last_error(const mesm_ctx* ctx);      /* ctx may be NULL: error of the last failed mesm_create */
int         mesm_abi_version(void);

/* replaces model.load_state_dict (eval.py:513-522).  `key` is the reference state_dict key; `data` is fp32, device
 * (is_device=1) or host memory; shape/ndim as in the state_dict.  Unknown keys and training-only keys are accepted and
 * ignored (returns 0).  Call mesm_finalize_weights once after the last tensor: it checks completeness and packs. */
int mesm_load_weight(mesm_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim,
                     int is_device, void* stream);
int mesm_finalize_weights(mesm_ctx* ctx, void* stream);

/* ---- forward: replaces MESM.forward in eval mode (model/model.py:154-359), called at eval.py:63 ---------------- */
typedef struct mesm_inputs {
    int32_t B;                     /* pairs = sum(num_clips) */
    int32_t Lv;                    /* padded clip count */
    int32_t Lt;                    /* padded word count */
    int32_t G;                     /* video groups */
    const void*    video_feat;     /* dev  [B,Lv,v_feat_dim] fp32 (fp16 when video_feat_f16 = 1) */
    const uint8_t* video_mask;     /* dev  [B,Lv] 1 = valid */
    const float*   words_feat;     /* dev  [B,Lt,t_feat_dim] — `words_id` of the text_encoder=None path (model.py:160-161) */
    const int64_t* num_clips;      /* HOST [G]  (the reference calls .tolist() on it, model.py:191) */
    const int64_t* neg_index;      /* dev  [B] or NULL: NULL skips the negative branch (model.py:260-302) */
    const int32_t* video_len;      /* HOST [B] or NULL.  Clip count of every pair, known to the host from the collate
                                    * step (utils/data_utils.py:64: `lengths`): video_mask[b, i] must be 0 for
                                    * i >= video_len[b], 1 <= video_len[b] <= Lv.  When given, the forward runs on
                                    * packed variable-length rows and spends no work on the zero padding; outputs at
                                    * valid clips are unchanged, row outputs at pad clips (saliency, projed/enhanced
                                    * video features, memory) are 0 instead of the reference's don't-care values
                                    * (eval.py:70-72 truncates them).  NULL: every pair is processed at Lv rows. */
    int32_t shared_group_video;    /* 1: the clips of pair b are read from the FIRST pair of its video group (the
                                    * charades / tacos collate replicates the video per query, dataset/base.py:307-309;
                                    * see mesm_upload_clips).  Needs video_len; must be 0 for qvhighlights grouping.
                                    * The first input projection (K = v_feat_dim, the HBM-bound stage) then runs once per
                                    * video instead of once per pair. */
    int32_t video_feat_f16;        /* 1: video_feat holds IEEE fp16 values - the 16-bit feature-storage option of the ingest
                                    * front-end.  The values are used exactly (parity = the reference fed the
                                    * same values upcast to fp32): the first projection multiplies the exact fp16 plane by
                                    * fp16 hi/lo weights on the tensor pipe (2 MMAs per product).  Needs video_len. */
} mesm_inputs;

/* Every pointer may be NULL (that output is then not materialised).  Shapes follow model/model.py:334-351. */
typedef struct mesm_outputs {
    float*   pred_logits;          /* [B,nq,2] */
    float*   pred_spans;           /* [B,nq,2] (center,width) */
    float*   saliency_scores;      /* [B,Lv] */
    float*   neg_saliency_scores;  /* [B,Lv] (needs neg_index) */
    float*   aux_logits;           /* [dec_layers-1,B,nq,2] */
    float*   aux_spans;            /* [dec_layers-1,B,nq,2] */
    float*   projed_video_feat;    /* [B,Lv,256] */
    float*   enhanced_video_feat;  /* [B,Lv,256] */
    float*   recon_feat;           /* [B,256] */
    float*   projed_recon_feat;    /* [B,256] */
    float*   expanded_words_feat;  /* [B,Lt+1,256]; projed_words_feat == [:,1:] */
    uint8_t* expanded_words_mask;  /* [B,Lt+1] */
    float*   memory;               /* [B,Lv,256]  encoder memory_local (tap) */
    float*   memory_global;        /* [B,256] (tap) */
    float*   hs;                   /* [dec_layers,B,nq,256] (tap) */
} mesm_outputs;

/* Bytes of caller-owned scratch mesm_forward needs for this shape (upper bound; independent of the masks). */
size_t mesm_workspace_bytes(const mesm_ctx* ctx, int32_t B, int32_t Lv, int32_t Lt, int32_t G);
int    mesm_forward(mesm_ctx* ctx, const mesm_inputs* in, const mesm_outputs* out, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Pairs per internal chunk (whole video groups are never split; default 256).  Affects workspace size only. */
int    mesm_set_chunk_pairs(mesm_ctx* ctx, int32_t pairs);
/* Number of kernel launches the last mesm_forward issued (bench.py's gpu_launches). */
int64_t mesm_last_launch_count(const mesm_ctx* ctx);
/* Clip-feature bytes the first input projection of the last mesm_forward streamed from HBM (bench.py's HBM roofline entry). */
int64_t mesm_last_feature_bytes(const mesm_ctx* ctx);

/* Measurement aid for bench.py: between begin/end every fused-linear launch of the calling thread is bracketed by CUDA
 * events on its stream.  end() synchronises those events and writes {linear ms, algorithmic flops, algorithmic bytes,
 * launches, ms / flops / launches of the launches with M >= 16384}. */
void mesm_profile_begin(void);
void mesm_profile_end(double* out7);
/* per-kernel-class text report of the last profiled region: lines "name<TAB>launches<TAB>ms" */
const char* mesm_profile_report(void);

/* ---- host -> device ingest of the collated batch ----------------------------------------------------------------- */
/* replaces the `value.to(device, non_blocking=...)` of `video_feat` / `video_mask` in prepare_batch_input
 * (dataset/base.py:358-363, called at eval.py:62).  host_feat [B,L,Dv] fp32 and host_mask [B,L] (1 = valid) are HOST
 * buffers (pinned for asynchronous copies).  Only the valid rows of every pair cross PCIe: the collate function zero-pads
 * each video to the longest of the batch (utils/data_utils.py:66-82), so the rows after a pair's last valid clip are
 * zeros; contiguous runs are merged into one copy each (all of them submitted in a single cudaMemcpyBatchAsync call when the
 * runtime has it and the stream is not the legacy default stream), the mask is copied whole and a kernel zero-fills
 * the rows with mask == 0 on the device: dev_feat ends up bit-identical to a plain copy of the zero-padded tensor.
 * num_clips (HOST [G], may be NULL): the charades / tacos collate replicates one video for every query of its group
 * (dataset/base.py:307-309).  When given, only the FIRST pair of each group is copied; the rows of the other pairs of
 * the group are left untouched, which is what mesm_forward reads with mesm_inputs.shared_group_video = 1.  Do not pass
 * it for qvhighlights batches (a group there holds different segments).
 * Stream-ordered, no synchronisation; *bytes_copied (may be NULL) receives the host->device bytes enqueued. */
int mesm_upload_clips(const float* host_feat, const uint8_t* host_mask, int32_t B, int32_t L, int32_t Dv,
                      float* dev_feat, uint8_t* dev_mask, const int64_t* num_clips, int32_t G, int64_t* bytes_copied,
                      void* stream);
/* the same for features stored as fp16 (host_feat / dev_feat [B,L,Dv] IEEE half): half the bytes cross PCIe */
int mesm_upload_clips_f16(const void* host_feat, const uint8_t* host_mask, int32_t B, int32_t L, int32_t Dv,
                          void* dev_feat, uint8_t* dev_mask, const int64_t* num_clips, int32_t G, int64_t* bytes_copied,
                          void* stream);

/* n host -> device copies in ONE submission (cudaMemcpyBatchAsync when the runtime has it, else a loop of cudaMemcpyAsync):
 * dst[i] <- src[i], bytes[i] each; src pinned host memory, dst on the device the stream belongs to (the current device).
 * Part of the prepare_batch_input replacement (dataset/base.py:358-363): the relayed upload of mesm_b200/relay.py stages the
 * per-video row slabs of a batch on a peer GPU with it. */
int mesm_memcpy_batch_h2d(const void* const* src, void* const* dst, const size_t* bytes, int64_t n, void* stream);

/* ---- CLIP text tower: replaces CLIPTextEncoder.forward (model/text_encoder.py:240-354), called by MESM.CLIP_encode_text
 * (model/model.py:103-109).  State-dict keys as in the reference: token_embedding.weight [vocab,width], positional_embedding
 * [context,width], transformer.resblocks.N.{ln_1,ln_2}.{weight,bias}, .attn.in_proj_{weight [3w,w],bias}, .attn.out_proj.*,
 * .mlp.c_fc.* [4w,w], .mlp.c_proj.* [w,4w], ln_final.*, text_projection [width,embed_dim].  The reference computes in fp16; this
 * runs fp32-in / fp32-out (bf16x3 products), text dev int64 [B,context]; last_hidden_state dev [B,context,width],
 * pooler_output dev [B,embed_dim] (may be NULL) = ln_final(x)[b, argmax(text[b])] @ text_projection. */
typedef struct mesm_clip mesm_clip;
mesm_clip*  mesm_clip_create(int32_t width, int32_t heads, int32_t layers, int32_t context_length, int32_t vocab_size,
                             int32_t embed_dim, int32_t device);
void        mesm_clip_destroy(mesm_clip* c);
const char* mesm_clip_last_error(const mesm_clip* c);
int    mesm_clip_load_weight(mesm_clip* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_device,
                             void* stream);
int    mesm_clip_finalize(mesm_clip* c, void* stream);
size_t mesm_clip_workspace_bytes(const mesm_clip* c, int32_t B);
int    mesm_clip_forward(mesm_clip* c, const int64_t* text, int32_t B, float* last_hidden_state, float* pooler_output,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- feature-ingest front-end: replaces get_video_feat (dataset/charades.py:108-119, dataset/qvhighlights.py:201-211) +
 * sample_video_feat (dataset/base.py:100-114) + add_tef (dataset/base.py:225-230) for ONE video.  raw: HOST array of S device
 * pointers to the raw per-source clip features [raw_len[s], dims[s]] (fp32, or fp16 when raw_f16; the reference upcasts on load);
 * normalize = the config's normalize_video (per-source L2 normalisation over the feature dim, eps 1e-12); sources are truncated
 * to the shortest one, concatenated, mean-pooled to max_video_l clips when longer, and the two tef columns are appended
 * (use_tef).  out: dev [mesm_video_feat_rows(...), sum(dims) + 2 * use_tef] fp32, or fp16 when out_f16 (the 16-bit storage
 * option of mesm_forward).  qvhighlights truncates the raw arrays to max_video_l rows before anything else: pass
 * min(raw_len, max_video_l) there. */
int32_t mesm_video_feat_rows(const int32_t* raw_len, int32_t S, int32_t max_video_l);
size_t  mesm_video_feat_workspace_bytes(const int32_t* raw_len, int32_t S, int32_t max_video_l);
int     mesm_build_video_feat(const void* const* raw, const int32_t* raw_len, const int32_t* dims, int32_t S, int32_t raw_f16,
                              int32_t normalize, int32_t max_video_l, int32_t use_tef, void* out, int32_t out_f16, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- span decode + post-processing + temporal NMS ------------------------------------------------------------- */
/* replaces eval.py:64-66,84-91 (softmax fg score, span_cxw_to_xx * duration, stable sort, 4-decimal rounding),
 * PostProcessorDETR as configured at eval.py:111-115 (utils/post_processing.py:22-47) and, when nms_thd != -1,
 * post_processing_mr_nms (eval.py:476-485 -> utils/temporal_nms.py:25-74). */
typedef struct mesm_decode_params {
    double  clip_len;          /* -1 disables round_multiple; used in fp32 like torch does with the Python scalar */
    double  min_ts_val;
    double  max_ts_val;
    double  nms_thd;           /* -1 disables NMS; compared in fp64 like utils/temporal_nms.py:52 */
    int32_t max_before_nms;
    int32_t max_after_nms;
    int32_t sort_results;
} mesm_decode_params;

int mesm_decode_nms(const float* pred_logits,   /* dev [B,nq,2] */
                    const float* pred_spans,    /* dev [B,nq,2] */
                    const float* duration,      /* dev [B] */
                    int32_t B, int32_t nq, const mesm_decode_params* p,
                    double*  windows,           /* dev [B,nq,3]  ranked [st,ed,score] after post-processing */
                    int32_t* order,             /* dev [B,nq]    query index at each rank */
                    int32_t* keep,              /* dev [B,max_after_nms] query indices kept by NMS, -1 padded (NULL ok) */
                    int32_t* keep_count,        /* dev [B] (NULL ok); a single candidate (min(max_before_nms, nq) == 1) is always
                                                   kept, even with max_after_nms == 0 (utils/temporal_nms.py:38-39) */
                    void* stream);
/* NMS ranks its candidates - the first max_before_nms windows of the output order - by their rounded score (stable), as
 * utils/temporal_nms.py:41 does, whether or not sort_results ranked the windows first.  The foreground score is
 * 1/(1+exp(l1-l0)) evaluated in fp64 and rounded once to fp32. */

/* replaces utils.temporal_nms (utils/temporal_nms.py:25-74) on ragged candidate lists:
 * windows dev [total,3] fp64 rows [st,ed,score]; list i = rows [offsets[i], offsets[i+1]); n_i <= 1024.
 * keep[i, :keep_count[i]] = positions (within list i) of the survivors in output order.  A list longer than 1024 rows is
 * not processed: its keep_count is set to -1 (the offsets live on the device, so the host cannot reject it up front). */
int mesm_temporal_nms(const double* windows, const int64_t* offsets, int32_t n_lists, double nms_thd,
                      int32_t max_after_nms, int32_t* keep, int32_t* keep_count, void* stream);

/* replaces PostProcessorDETR.__call__ (utils/post_processing.py:22-47) with process_func_names ("clip_ts",
 * "round_multiple") [clip_len == -1: ("clip_ts",)] on n already-decoded fp64 rows [st, ed, score]. */
int mesm_post_process(const double* windows, double* out, int64_t n, double clip_len, double min_ts_val,
                      double max_ts_val, void* stream);

/* replaces utils/span_utils.py: temporal_iou (45-72), generalized_temporal_iou (92-121) on fp32 [N,2] x [M,2] -> [N,M];
 * any output may be NULL. */
int mesm_temporal_iou(const float* spans1, int32_t N, const float* spans2, int32_t M, float* iou, float* uni,
                      float* giou, void* stream);
/* span_cxw_to_xx (26-42) when to_xx != 0 else span_xx_to_cxw (5-23), n rows of 2. */
int mesm_span_convert(const float* in, float* out, int64_t n, int to_xx, void* stream);

/* ---- moment-retrieval metrics: the per-query work of eval_moment_retrieval (eval.py:233-263): compute_mr_r1 (eval.py:397-425),
 * compute_average_precision_detection (eval.py:323-394), interpolated_precision_recall (utils/data_utils.py:166-182), for every
 * ground-truth length range of get_data_by_range (eval.py:428-461) at once.  windows dev [B,nq,3] fp64 = the ranked output of
 * mesm_decode_nms (the first max_pred_windows rows are scored, eval.py:266); gt_windows dev [total,2] fp64 with gt_offsets dev
 * [B+1] (<= 32 windows per query); length_ranges dev [n_ranges,2] (min_l, max_l], min_l < 0 = keep everything (the "full" range);
 * iou_thds dev [n_thds <= 16].  Outputs (dev): in_range [n_ranges,B] (the query has a ground-truth window in the range),
 * top1_iou [n_ranges,B] (max IoU of the first window with the in-range ground truth), ap [n_ranges,B,n_thds]. */
int mesm_mr_metrics(const double* windows, int32_t B, int32_t nq, int32_t max_pred_windows, const double* gt_windows,
                    const int64_t* gt_offsets, const double* length_ranges, int32_t n_ranges, const double* iou_thds, int32_t n_thds,
                    uint8_t* in_range, double* top1_iou, double* ap, void* stream);

/* ---- eval-time saliency criterion: replaces Criterion.loss_saliency (model/criterion.py:139-221), which train.py's per-epoch
 * evaluation applies to the forward's outputs (eval.py:101-105).  saliency_scores / neg_saliency_scores dev [B,L] (MESM.forward),
 * video_mask dev [B,L] (1 = valid), label dev [B,L] fp32 = targets["saliency_label"] or targets["clip_mask"].float() (:155-158),
 * rank_coef = the config's value (12 in every shipped config); num_pairs > 0 adds the triplet term of `use_triplet` configs with
 * pos_idx / neg_idx dev int64 [B,num_pairs] and saliency_margin.  out4 (dev) = {loss_saliency, loss_neg_pair,
 * loss_rank_contrastive, loss_triplet}. */
size_t mesm_saliency_loss_workspace_bytes(int32_t B);
int mesm_saliency_loss(const float* saliency_scores, const float* neg_saliency_scores, const uint8_t* video_mask, const float* label,
                       int32_t B, int32_t L, float rank_coef, const int64_t* pos_idx, const int64_t* neg_idx, int32_t num_pairs,
                       float saliency_margin, float* out4, void* workspace, size_t workspace_bytes, void* stream);

/* ---- segment-sentence alignment scores: replaces model/criterion.py:241-266 (up to cos_sim / tau) --------------- */
int mesm_align_scores(const float* projed_video_feat,    /* dev [B,Lv,256] */
                      const uint8_t* clip_mask,          /* dev [B,Lv] */
                      const float* expanded_words_feat,  /* dev [B,Lw,256] */
                      const uint8_t* expanded_words_mask,/* dev [B,Lw] */
                      int32_t B, int32_t Lv, int32_t Lw, float tau,
                      float* scores,                     /* dev [B,B] */
                      void* workspace, size_t workspace_bytes, /* >= mesm_align_workspace_bytes(B) */
                      void* stream);
size_t mesm_align_workspace_bytes(int32_t B);

/* ---- sub-module entry points (drop-in T2VEncoder / Transformer / MultiheadAttention classes) ------------------- */
/* projection-free MHA core of model/attention.py:185-394: q [L,B,E] k [S,B,E] v [S,B,Ev] (seq-first like the
 * reference), out_proj weight [Ev,Ev] + bias; key_padding_mask [B,S] 1 = PAD or NULL. out [L,B,Ev];
 * attn_weights [B,L,S] head-averaged or NULL. */
int mesm_mha_noproj(const float* q, const float* k, const float* v, int32_t L, int32_t S, int32_t B, int32_t E,
                    int32_t Ev, int32_t nheads, const float* out_w, const float* out_b,
                    const uint8_t* key_padding_mask, float* out, float* attn_weights,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t mesm_mha_workspace_bytes(int32_t L, int32_t S, int32_t B, int32_t E, int32_t Ev);

/* T2VEncoder(.forward) (model/transformer.py:83-105) with the weights loaded under `prefix`
 * ("enhance_encoder", "t2v_encoder"): src_txt [B,Lt,256], src_vid [B,Lv,256], *_pad [B,L] 1 = PAD,
 * pos_txt / pos_vid [B,L,256] or NULL.  out [B,Lv,256]. */
int mesm_t2v_encoder(mesm_ctx* ctx, const char* prefix, const float* src_txt, const float* src_vid,
                     const uint8_t* txt_pad, const uint8_t* vid_pad, const float* pos_txt, const float* pos_vid,
                     int32_t B, int32_t Lt, int32_t Lv, float* out, void* workspace, size_t workspace_bytes,
                     void* stream);
size_t mesm_t2v_workspace_bytes(int32_t B, int32_t Lt, int32_t Lv);

/* Transformer(.forward) (model/transformer.py:174-205): src [B,L,256], pad [B,L] 1 = PAD, query_embed [nq,2],
 * pos_embed [B,L,256], global token / pos [256] (the reference repeats them over the batch, model.py:237-238).
 * hs [nl,B,nq,256], references [nl,B,nq,2], memory_local [B,L,256], memory_global [B,256]. */
int mesm_transformer(mesm_ctx* ctx, const float* src, const uint8_t* pad, const float* query_embed,
                     const float* pos_embed, const float* global_token, const float* global_token_pos,
                     int32_t B, int32_t L, float* hs, float* references, float* memory_local, float* memory_global,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t mesm_transformer_workspace_bytes(const mesm_ctx* ctx, int32_t B, int32_t L);

/* Barrier watchdog of the tcgen05 kernels (out128 = {attention[64], linear[64]}; [0] = number of barrier waits that gave up after ~2 s, then {tag, block, thread|barrier|parity} triples). */
int mesm_debug_watchdog(unsigned long long* out128);

/* In-kernel timeline of the pipelined self-attention kernel (library built with -DMESM_ATP_TRACE, tools/attn_trace.py): clock64
 * stamps of CTA 0, [head][event]; returns the number of values written (0 in a normal build). */
int mesm_debug_attn_trace(long long* out128);

/* Test hook (tests/test_gpu_attention.py): repeated self-attention launches + the barrier watchdog record. */
int mesm_debug_attention(const float* qkv, const uint8_t* k_pad, int32_t B, int32_t L, float* out, int32_t use_tc, int32_t iters,
                         unsigned long long* watchdog8, void* stream);

/* Test hook (tests/test_gpu_linear.py): one fused linear out = epilogue(A (+Apos) . W^T) through the fp32 SIMT kernel
 * (use_tc = 0) or the tcgen05 split-bf16 kernel (use_tc = 1); synchronises the stream. */
int mesm_debug_linear(const float* A, const float* Apos, const float* W, const float* bias, const float* residual,
                      const float* ln_g, const float* ln_b, const float* rowstat, const float* prelu, int32_t M, int32_t N,
                      int32_t K, int32_t lda, int32_t act, float out_scale, float* out, float* pre_ln, int32_t use_tc,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MESM_B200_H */
