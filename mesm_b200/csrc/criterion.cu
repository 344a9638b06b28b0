// Eval-time criterion scores that train.py's per-epoch evaluation computes on the forward's outputs (eval.py:101-105):
// the saliency terms of Criterion.loss_saliency (model/criterion.py:139-221).  Row-wise work: one warp per pair, then one
// block folds the per-pair partials into the batch means the reference takes.
//
//   loss_neg_pair          = mean_b sum_l -log(1 - sigmoid(neg[b,l])) * mask[b,l]                                   (:148-150)
//   loss_rank_contrastive  = (1 / rank_coef) * sum_{r=1..11, any label >= r in the batch} mean_b loss_r[b]               (:166-197)
//       s = cat(sal, neg) masked to -1e3, / tau (0.5);  log_prob = (s - max) - log(sum exp(s - max) + 1e-6)
//       loss_r[b] = -[sum_j (label2[b,j] >= r) * log_prob[b,j] * mask2[b,j]] / (count_r[b] + 1e-6) * (count_r[b] > 0)
//   loss_triplet (optional) = sum clamp(margin + sal[b,neg_idx] - sal[b,pos_idx], 0) / (B * P) * 2                        (:202-213)
#include "common.cuh"
#include "../../include/mesm_b200.h"
#include <math_constants.h>

namespace mesm {

constexpr int kRanks = 11;          // range(1, 12) at model/criterion.py:169

__global__ void __launch_bounds__(256) saliency_rows_kernel(const float* __restrict__ sal, const float* __restrict__ neg, const uint8_t* __restrict__ vmask,
                                                            const float* __restrict__ label, int B, int L, float* __restrict__ part) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* s0 = sal + (long long)b * L;
    const float* s1 = neg + (long long)b * L;
    const uint8_t* m = vmask + (long long)b * L;
    const float* lb = label + (long long)b * L;
    // cur = (mask * score + (1 - mask) * -1e3) / tau over the 2L concatenated entries
    float mx = -CUDART_INF_F, negp = 0.f;
    for (int j = lane; j < 2 * L; j += 32) {
        const int l = j < L ? j : j - L;
        const float mk = m[l] ? 1.f : 0.f;
        const float v = (mk * (j < L ? s0[l] : s1[l]) + (1.f - mk) * -1e3f) / 0.5f;
        mx = fmaxf(mx, v);
        if (j >= L && m[l]) negp += -logf(1.f - 1.f / (1.f + expf(-s1[l])));
    }
    mx = warp_max(mx);
    negp = warp_sum(negp);
    float se = 0.f;
    for (int j = lane; j < 2 * L; j += 32) {
        const int l = j < L ? j : j - L;
        const float mk = m[l] ? 1.f : 0.f;
        const float v = (mk * (j < L ? s0[l] : s1[l]) + (1.f - mk) * -1e3f) / 0.5f;
        se += expf(v - mx);
    }
    se = warp_sum(se);
    const float lse = logf(se + 1e-6f);
    float acc[kRanks], cnt[kRanks];
#pragma unroll
    for (int r = 0; r < kRanks; ++r) { acc[r] = 0.f; cnt[r] = 0.f; }
    for (int l = lane; l < L; l += 32) {            // the appended negative half carries label 0: never positive
        const float lab = lb[l];
        if (lab < 1.f) continue;
        const float mk = m[l] ? 1.f : 0.f;
        const float v = (mk * s0[l] + (1.f - mk) * -1e3f) / 0.5f;
        const float lp = (v - mx) - lse;
#pragma unroll
        for (int r = 0; r < kRanks; ++r)
            if (lab >= (float)(r + 1)) { acc[r] += lp * mk; cnt[r] += 1.f; }
    }
    float* o = part + (long long)b * (2 * kRanks + 1);
#pragma unroll
    for (int r = 0; r < kRanks; ++r) {
        const float a = warp_sum(acc[r]), c = warp_sum(cnt[r]);
        if (lane == 0) { o[2 * r] = c > 0.f ? -(a / (c + 1e-6f)) : 0.f; o[2 * r + 1] = c; }
    }
    if (lane == 0) o[2 * kRanks] = negp;
}

__global__ void __launch_bounds__(256) saliency_finalize_kernel(const float* __restrict__ part, int B, float rank_coef, const float* __restrict__ sal, int L,
                                                                const int64_t* __restrict__ pos_idx, const int64_t* __restrict__ neg_idx, int P, float margin,
                                                                float* __restrict__ out) {
    __shared__ double red[2 * kRanks + 2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[2 * kRanks + 2];
    for (int i = 0; i < 2 * kRanks + 2; ++i) v[i] = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* o = part + (long long)b * (2 * kRanks + 1);
        for (int i = 0; i < 2 * kRanks + 1; ++i) v[i] += (double)o[i];
        if (P > 0)
            for (int c = 0; c < P; ++c) {
                const float ps = sal[(long long)b * L + pos_idx[(long long)b * P + c]], ns = sal[(long long)b * L + neg_idx[(long long)b * P + c]];
                v[2 * kRanks + 1] += (double)fmaxf(margin + ns - ps, 0.f);
            }
    }
    for (int i = 0; i < 2 * kRanks + 2; ++i) {
        double x = v[i];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[i][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[2 * kRanks + 2];
        for (int i = 0; i < 2 * kRanks + 2; ++i) { tot[i] = 0.0; for (int w = 0; w < 8; ++w) tot[i] += red[i][w]; }
        double rank = 0.0;
        for (int r = 0; r < kRanks; ++r)
            if (tot[2 * r + 1] > 0.0) rank += tot[2 * r] / B;                  // `continue` when no label of the batch reaches r (:173-174)
        rank /= rank_coef;
        const double negp = tot[2 * kRanks] / B;
        const double trip = P > 0 ? tot[2 * kRanks + 1] / ((double)B * P) * 2.0 : 0.0;
        out[0] = (float)(rank + negp + trip);       // loss_saliency (:216-221)
        out[1] = (float)negp;
        out[2] = (float)rank;
        out[3] = (float)trip;
    }
}

}  // namespace mesm

using namespace mesm;

extern "C" size_t mesm_saliency_loss_workspace_bytes(int32_t B) { return (size_t)B * (2 * kRanks + 1) * sizeof(float) + 256; }

extern "C" int mesm_saliency_loss(const float* saliency_scores, const float* neg_saliency_scores, const uint8_t* video_mask, const float* label,
                                  int32_t B, int32_t L, float rank_coef, const int64_t* pos_idx, const int64_t* neg_idx, int32_t num_pairs,
                                  float saliency_margin, float* out4, void* workspace, size_t workspace_bytes, void* stream) {
    if (!saliency_scores || !neg_saliency_scores || !video_mask || !label || !out4 || !workspace || B < 1 || L < 1 || rank_coef == 0.f)
        return (int)cudaErrorInvalidValue;
    if (workspace_bytes < mesm_saliency_loss_workspace_bytes(B)) return (int)cudaErrorInvalidValue;
    if (num_pairs > 0 && (!pos_idx || !neg_idx)) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    float* part = (float*)workspace;
    saliency_rows_kernel<<<(B + 7) / 8, 256, 0, s>>>(saliency_scores, neg_saliency_scores, video_mask, label, B, L, part);
    saliency_finalize_kernel<<<1, 256, 0, s>>>(part, B, rank_coef, saliency_scores, L, pos_idx, neg_idx, num_pairs > 0 ? num_pairs : 0, saliency_margin, out4);
    g_stats.launches += 2;
    return (int)cudaGetLastError();
}
