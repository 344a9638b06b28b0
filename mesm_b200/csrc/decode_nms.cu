// Span decode -> post-processing -> greedy temporal NMS, and the span_utils device functions.
//
// mesm_decode_nms: one warp per (video, query) pair, one lane per span query.  Follows, step by step and in the same
// arithmetic types, eval.py:64-66,84-91 (fp32 softmax / span_cxw_to_xx * duration, stable descending sort on the
// unrounded score, float(f"{x:.4f}") == rint(x*1e4)/1e4 in fp64), utils/post_processing.py:22-47 (back to fp32, clamp,
// round-half-even to multiples of clip_len, score re-rounded) and utils/temporal_nms.py:6-74 (fp64 hull-"IoU",
// strict >, at most max_after_nms survivors).  All fp32 steps use explicit round-to-nearest intrinsics so that nvcc
// cannot contract them into FMAs: the result is bit-identical to the reference's CPU arithmetic.
#include "common.cuh"
#include "../../include/mesm_b200.h"

namespace mesm {

__device__ __forceinline__ double round4(double x) { return rint(x * 1e4) / 1e4; }

__device__ __forceinline__ double hull_iou(double s0, double e0, double s1, double e1) {
    const double inter = fmax(0.0, fmin(e0, e1) - fmax(s0, s1));
    const double uni = fmax(e0, e1) - fmin(s0, s1);
    return uni == 0.0 ? 0.0 : inter / uni;
}

struct DecodeArgs {
    const float* logits; const float* spans; const float* duration;
    int B, nq;
    float clip_len, min_ts, max_ts;
    double nms_thd;
    int max_before, max_after, sort_results, do_nms;
    double* windows; int* order; int* keep; int* keep_count;
};

__global__ void __launch_bounds__(128) decode_nms_kernel(const DecodeArgs a) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= a.B) return;
    const int nq = a.nq;
    const bool act = lane < nq;

    float score = -1.f, st = 0.f, ed = 0.f;
    if (act) {
        const float l0 = a.logits[((long long)b * nq + lane) * 2], l1 = a.logits[((long long)b * nq + lane) * 2 + 1];
        // softmax over two classes = 1 / (1 + exp(l1 - l0)); evaluated in fp64 and rounded once to fp32, i.e. the correctly
        // rounded value of what eval.py:64-66 computes in fp32 (CUDA expf and the host libm differ in the last ulp, which
        // could cross a 4-decimal rounding boundary; the fp64 value cannot be further than that ulp from either)
        score = (float)(1.0 / (1.0 + exp((double)l1 - (double)l0)));
        const float c = a.spans[((long long)b * nq + lane) * 2], w = a.spans[((long long)b * nq + lane) * 2 + 1];
        const float dur = a.duration[b];
        const float hw = __fmul_rn(0.5f, w);
        st = __fmul_rn(__fsub_rn(c, hw), dur);                                // span_cxw_to_xx * duration
        ed = __fmul_rn(__fadd_rn(c, hw), dur);
    }
    // stable descending rank (Python sorted(..., reverse=True) keeps the original order of equal keys)
    int rank = lane;
    if (a.sort_results) {
        rank = 0;
        for (int j = 0; j < nq; ++j) {
            const float sj = __shfl_sync(0xffffffffu, score, j);
            if (act && (sj > score || (sj == score && j < lane))) ++rank;
        }
    }
    // 4-decimal rounding in fp64, then PostProcessorDETR in fp32
    double st4 = round4((double)st), ed4 = round4((double)ed), sc4 = round4((double)score);
    float stf = (float)st4, edf = (float)ed4;
    stf = fminf(fmaxf(stf, a.min_ts), a.max_ts);
    edf = fminf(fmaxf(edf, a.min_ts), a.max_ts);
    if (a.clip_len != -1.f) {
        stf = __fmul_rn(rintf(__fdiv_rn(stf, a.clip_len)), a.clip_len);
        edf = __fmul_rn(rintf(__fdiv_rn(edf, a.clip_len)), a.clip_len);
    }
    const double scf = round4((double)(float)sc4);
    if (act) {
        double* w = a.windows + ((long long)b * nq + rank) * 3;
        w[0] = (double)stf; w[1] = (double)edf; w[2] = scf;
        a.order[(long long)b * nq + rank] = lane;
    }
    if (!a.do_nms || (a.keep == nullptr && a.keep_count == nullptr)) return;      // keep may be NULL when max_after_nms == 0

    // bring the ranked list into lane order: lane r holds the window of rank r
    int src = 0;
    for (int j = 0; j < nq; ++j) {
        const int rj = __shfl_sync(0xffffffffu, rank, j);
        if (rj == lane && j < nq) src = j;
    }
    double rs = __shfl_sync(0xffffffffu, (double)stf, src);
    double re = __shfl_sync(0xffffffffu, (double)edf, src);
    int rq = __shfl_sync(0xffffffffu, lane, src);
    const int nb = min(a.max_before, nq);
    // utils/temporal_nms.py:41 re-sorts its input (the first max_before_nms windows) by the ROUNDED score, stable.  For a
    // ranked list (sort_results) rounding is monotone and the order is unchanged; for an unranked list (sort_results = 0)
    // this is the only sort.  Done unconditionally: lane r takes the window of NMS rank r.
    {
        const double rsc = __shfl_sync(0xffffffffu, scf, src);
        int nrank = lane;
        if (lane < nb) nrank = 0;
        for (int j = 0; j < nb; ++j) {
            const double sj = __shfl_sync(0xffffffffu, rsc, j);
            if (lane < nb && (sj > rsc || (sj == rsc && j < lane))) ++nrank;
        }
        int nsrc = lane;
        for (int j = 0; j < nb; ++j) {
            const int rj = __shfl_sync(0xffffffffu, nrank, j);
            if (rj == lane) nsrc = j;
        }
        rs = __shfl_sync(0xffffffffu, rs, nsrc);
        re = __shfl_sync(0xffffffffu, re, nsrc);
        rq = __shfl_sync(0xffffffffu, rq, nsrc);
    }
    unsigned alive = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    int kept = 0;
    int* kp = a.keep ? a.keep + (long long)b * a.max_after : nullptr;
    if (nb == 1) {                                  // utils/temporal_nms.py:38-39: a single prediction is returned untouched,
        if (lane == 0) {                            // whatever max_after_nms says
            if (kp && a.max_after >= 1) kp[0] = rq;
            for (int i = 1; kp && i < a.max_after; ++i) kp[i] = -1;
            if (a.keep_count) a.keep_count[b] = 1;
        }
        return;
    }
    while (alive && kept < a.max_after) {
        const int head = __ffs(alive) - 1;
        const double hs = __shfl_sync(0xffffffffu, rs, head), he = __shfl_sync(0xffffffffu, re, head);
        const int hq = __shfl_sync(0xffffffffu, rq, head);
        const bool sup = (lane > head) && ((alive >> lane) & 1u) && (hull_iou(hs, he, rs, re) > a.nms_thd);
        const unsigned supmask = __ballot_sync(0xffffffffu, sup);
        alive &= ~supmask;
        alive &= ~(1u << head);
        if (lane == 0 && kp) kp[kept] = hq;
        ++kept;
    }
    if (lane == 0) {
        for (int i = kept; kp && i < a.max_after; ++i) kp[i] = -1;
        if (a.keep_count) a.keep_count[b] = min(kept, a.max_after);
    }
}

// ---- ragged-list temporal NMS (utils/temporal_nms.py:25-74): one CTA per list, n <= 1024 ------------------------------
__global__ void __launch_bounds__(256) temporal_nms_kernel(const double* __restrict__ windows, const int64_t* __restrict__ offsets,
                                                           double nms_thd, int max_after, int* __restrict__ keep,
                                                           int* __restrict__ keep_count) {
    __shared__ double s_st[1024], s_ed[1024];
    __shared__ int s_pos[1024];
    __shared__ unsigned char s_alive[1024];
    const int li = blockIdx.x;
    const long long lo = offsets[li];
    const int n = (int)(offsets[li + 1] - lo);
    int* kp = keep + (long long)li * max_after;
    for (int i = threadIdx.x; i < max_after; i += blockDim.x) kp[i] = -1;
    if (n > 1024 || n < 0) { if (threadIdx.x == 0) keep_count[li] = -1; return; }     // list too long for the shared-memory tables: flagged, not processed
    if (n == 0) { if (threadIdx.x == 0) keep_count[li] = 0; return; }
    if (n == 1) {                                                      // :38-39 returns the list untouched
        if (threadIdx.x == 0) { if (max_after >= 1) kp[0] = 0; keep_count[li] = 1; }
        return;
    }
    // stable descending rank sort
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double si = windows[(lo + i) * 3 + 2];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const double sj = windows[(lo + j) * 3 + 2];
            if (sj > si || (sj == si && j < i)) ++r;
        }
        s_st[r] = windows[(lo + i) * 3];
        s_ed[r] = windows[(lo + i) * 3 + 1];
        s_pos[r] = i;
        s_alive[r] = 1;
    }
    __syncthreads();
    int kept = 0;
    for (int i = 0; i < n && kept < max_after; ++i) {
        if (!s_alive[i]) continue;                                      // uniform: shared value
        const double hs = s_st[i], he = s_ed[i];
        for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x)
            if (s_alive[j] && hull_iou(hs, he, s_st[j], s_ed[j]) > nms_thd) s_alive[j] = 0;
        if (threadIdx.x == 0) kp[kept] = s_pos[i];
        ++kept;
        __syncthreads();
    }
    if (threadIdx.x == 0) keep_count[li] = kept;
}

// ---- span_utils ------------------------------------------------------------------------------------------------------
__global__ void temporal_iou_kernel(const float* __restrict__ s1, int N, const float* __restrict__ s2, int M,
                                    float* __restrict__ iou, float* __restrict__ uni, float* __restrict__ giou) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * M) return;
    const int i = (int)(idx / M), j = (int)(idx % M);
    const float a0 = s1[2 * i], a1 = s1[2 * i + 1], b0 = s2[2 * j], b1 = s2[2 * j + 1];
    const float area1 = __fsub_rn(a1, a0), area2 = __fsub_rn(b1, b0);
    const float inter = fmaxf(__fsub_rn(fminf(a1, b1), fmaxf(a0, b0)), 0.f);
    const float u = __fsub_rn(__fadd_rn(area1, area2), inter);
    const float v = __fdiv_rn(inter, u);
    if (iou) iou[idx] = v;
    if (uni) uni[idx] = u;
    if (giou) {
        const float enc = fmaxf(__fsub_rn(fmaxf(a1, b1), fminf(a0, b0)), 0.f);
        giou[idx] = __fsub_rn(v, __fdiv_rn(__fsub_rn(enc, u), enc));
    }
}

__global__ void span_convert_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int to_xx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = in[2 * i], b = in[2 * i + 1];
    if (to_xx) {                       // (center, width) -> (st, ed), utils/span_utils.py:40-42
        const float hw = __fmul_rn(0.5f, b);
        out[2 * i] = __fsub_rn(a, hw);
        out[2 * i + 1] = __fadd_rn(a, hw);
    } else {                           // (st, ed) -> (center, width), utils/span_utils.py:21-23
        out[2 * i] = __fmul_rn(__fadd_rn(a, b), 0.5f);
        out[2 * i + 1] = __fsub_rn(b, a);
    }
}

// PostProcessorDETR on already-decoded windows (utils/post_processing.py:22-47): fp64 rows -> fp32 -> clamp ->
// round to multiples of clip_len -> score re-rounded to 4 decimals.
__global__ void post_process_kernel(const double* __restrict__ in, double* __restrict__ out, long long n, float clip_len,
                                    float min_ts, float max_ts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float v = fminf(fmaxf((float)in[i * 3 + c], min_ts), max_ts);
        if (clip_len != -1.f) v = __fmul_rn(rintf(__fdiv_rn(v, clip_len)), clip_len);
        out[i * 3 + c] = (double)v;
    }
    out[i * 3 + 2] = round4((double)(float)in[i * 3 + 2]);
}

}  // namespace mesm

using namespace mesm;

extern "C" int mesm_post_process(const double* windows, double* out, int64_t n, double clip_len, double min_ts_val,
                                 double max_ts_val, void* stream) {
    if (!windows || !out) return (int)cudaErrorInvalidValue;
    if (n <= 0) return 0;
    post_process_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(windows, out, n, (float)clip_len,
                                                                                        (float)min_ts_val, (float)max_ts_val);
    g_stats.launches++;
    return (int)cudaGetLastError();
}

extern "C" int mesm_decode_nms(const float* pred_logits, const float* pred_spans, const float* duration, int32_t B,
                               int32_t nq, const mesm_decode_params* p, double* windows, int32_t* order, int32_t* keep,
                               int32_t* keep_count, void* stream) {
    if (!pred_logits || !pred_spans || !duration || !p || !windows || !order) return (int)cudaErrorInvalidValue;
    if (nq < 1 || nq > 32 || B < 0) return (int)cudaErrorInvalidValue;
    if (B == 0) return 0;
    DecodeArgs a;
    a.logits = pred_logits; a.spans = pred_spans; a.duration = duration; a.B = B; a.nq = nq;
    a.clip_len = (float)p->clip_len; a.min_ts = (float)p->min_ts_val; a.max_ts = (float)p->max_ts_val;
    a.nms_thd = p->nms_thd; a.max_before = p->max_before_nms; a.max_after = p->max_after_nms;
    a.sort_results = p->sort_results; a.do_nms = (p->nms_thd != -1.0) ? 1 : 0;
    a.windows = windows; a.order = order; a.keep = keep; a.keep_count = keep_count;
    if (a.do_nms && (a.max_after < 0 || a.max_before < 1 || (!keep && a.max_after > 0 && keep_count))) return (int)cudaErrorInvalidValue;
    decode_nms_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(a);
    g_stats.launches++;
    return (int)cudaGetLastError();
}

extern "C" int mesm_temporal_nms(const double* windows, const int64_t* offsets, int32_t n_lists, double nms_thd,
                                 int32_t max_after_nms, int32_t* keep, int32_t* keep_count, void* stream) {
    if (!windows || !offsets || (!keep && max_after_nms > 0) || !keep_count || max_after_nms < 0) return (int)cudaErrorInvalidValue;
    if (n_lists <= 0) return 0;
    temporal_nms_kernel<<<n_lists, 256, 0, (cudaStream_t)stream>>>(windows, offsets, nms_thd, max_after_nms, keep, keep_count);
    g_stats.launches++;
    return (int)cudaGetLastError();
}

extern "C" int mesm_temporal_iou(const float* spans1, int32_t N, const float* spans2, int32_t M, float* iou, float* uni,
                                 float* giou, void* stream) {
    if (!spans1 || !spans2) return (int)cudaErrorInvalidValue;
    const long long n = (long long)N * M;
    if (n <= 0) return 0;
    temporal_iou_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(spans1, N, spans2, M, iou, uni, giou);
    g_stats.launches++;
    return (int)cudaGetLastError();
}

extern "C" int mesm_span_convert(const float* in, float* out, int64_t n, int to_xx, void* stream) {
    if (!in || !out) return (int)cudaErrorInvalidValue;
    if (n <= 0) return 0;
    span_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, to_xx);
    g_stats.launches++;
    return (int)cudaGetLastError();
}
