// Moment-retrieval metrics on the device: the per-query part of eval_moment_retrieval (eval.py:233-263) -
// compute_mr_r1 (eval.py:397-425) and compute_average_precision_detection (eval.py:323-394) with
// interpolated_precision_recall (utils/data_utils.py:166-182) and the hull "IoU" of compute_temporal_iou_batch_cross
// (utils/span_utils.py:124-151) - for every query and every ground-truth length range at once.  One thread per
// (range, query); fp64 like the reference's numpy code.  The batch means / percent formatting stay with the caller
// (mesm_b200.utils.eval_moment_retrieval) because they are a handful of reductions over the arrays written here.
#include "common.cuh"
#include "../../include/mesm_b200.h"

namespace mesm {

constexpr int kMaxPred = 16, kMaxGt = 32, kMaxThd = 16;

__device__ __forceinline__ double hull_iou_d(double s0, double e0, double s1, double e1) {
    const double inter = fmax(0.0, fmin(e0, e1) - fmax(s0, s1));
    const double uni = fmax(e0, e1) - fmin(s0, s1);
    return uni != 0.0 ? inter / uni : 0.0;                      // np.divide(..., where=union != 0) on a zero array
}

struct MetricsArgs {
    const double* windows; int nq, npred;                       // [B, nq, 3] ranked [st, ed, score]; the first npred are scored
    const double* gt; const int64_t* gt_off;                    // ragged ground-truth windows [total, 2], offsets [B + 1]
    const double* ranges; int nr;                               // [nr, 2] (min_l, max_l]; a negative min_l = no filter (the "full" range)
    const double* thds; int nt;                                 // AP IoU thresholds
    int B;
    uint8_t* in_range; double* top1_iou; double* ap;            // [nr, B], [nr, B], [nr, B, nt]
};

__global__ void mr_metrics_kernel(const MetricsArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.nr * a.B) return;
    const int r = (int)(idx / a.B), b = (int)(idx - (long long)r * a.B);
    const double lo = a.ranges[2 * r], hi = a.ranges[2 * r + 1];
    double gs[kMaxGt], ge[kMaxGt];
    int ng = 0;
    for (long long j = a.gt_off[b]; j < a.gt_off[b + 1] && ng < kMaxGt; ++j) {
        const double s = a.gt[2 * j], e = a.gt[2 * j + 1];
        if (lo < 0.0 || (lo < e - s && e - s <= hi)) { gs[ng] = s; ge[ng] = e; ++ng; }      // get_data_by_range, eval.py:441-461
    }
    a.in_range[idx] = ng > 0;
    double* ap = a.ap + idx * a.nt;
    for (int t = 0; t < a.nt; ++t) ap[t] = 0.0;
    a.top1_iou[idx] = 0.0;
    if (ng == 0) return;
    const double* w = a.windows + (long long)b * a.nq * 3;
    // R1: IoU of the top-1 window with the ground-truth window it overlaps most (eval.py:405-418)
    {
        double best = 0.0;
        for (int g = 0; g < ng; ++g) best = fmax(best, hull_iou_d(w[0], w[1], gs[g], ge[g]));
        a.top1_iou[idx] = best;
    }
    // AP: predictions in score order (already ranked; equal 4-decimal scores keep their order like the stable list.sort)
    const int np_ = min(a.npred, min(a.nq, kMaxPred));
    for (int t = 0; t < a.nt; ++t) {
        const double thd = a.thds[t];
        int lock[kMaxGt];
        for (int g = 0; g < ng; ++g) lock[g] = -1;
        unsigned tp_mask = 0;
        for (int p = 0; p < np_; ++p) {
            // ground-truth windows by decreasing IoU: the first unlocked one with IoU >= thd is matched (eval.py:366-381)
            unsigned tried = 0;
            for (int it = 0; it < ng; ++it) {
                int bj = -1; double bi = -1.0;
                for (int g = 0; g < ng; ++g) {
                    if (tried >> g & 1u) continue;
                    const double i_ = hull_iou_d(w[3 * p], w[3 * p + 1], gs[g], ge[g]);
                    if (i_ > bi) { bi = i_; bj = g; }
                }
                tried |= 1u << bj;
                if (bi < thd) break;                            // false positive
                if (lock[bj] >= 0) continue;
                lock[bj] = p; tp_mask |= 1u << p;
                break;
            }
        }
        // interpolated precision / recall (VOC 2011): precision envelope from the right, summed where recall changes
        double prec[kMaxPred + 2], rec[kMaxPred + 2];
        int tpc = 0;
        prec[0] = 0.0; rec[0] = 0.0;
        for (int p = 0; p < np_; ++p) {
            tpc += (tp_mask >> p) & 1u;
            prec[p + 1] = (double)tpc / (double)(p + 1);         // tp / (tp + fp)
            rec[p + 1] = (double)tpc / (double)ng;
        }
        prec[np_ + 1] = 0.0; rec[np_ + 1] = 1.0;
        for (int i = np_; i >= 0; --i) prec[i] = fmax(prec[i], prec[i + 1]);
        double s = 0.0;
        for (int i = 1; i <= np_ + 1; ++i)
            if (rec[i] != rec[i - 1]) s += (rec[i] - rec[i - 1]) * prec[i];
        ap[t] = s;
    }
}

}  // namespace mesm

using namespace mesm;

extern "C" int mesm_mr_metrics(const double* windows, int32_t B, int32_t nq, int32_t max_pred_windows, const double* gt_windows,
                               const int64_t* gt_offsets, const double* length_ranges, int32_t n_ranges, const double* iou_thds, int32_t n_thds,
                               uint8_t* in_range, double* top1_iou, double* ap, void* stream) {
    if (!windows || !gt_windows || !gt_offsets || !length_ranges || !iou_thds || !in_range || !top1_iou || !ap) return (int)cudaErrorInvalidValue;
    if (B < 1 || nq < 1 || n_ranges < 1 || n_thds < 1 || n_thds > kMaxThd || max_pred_windows < 1) return (int)cudaErrorInvalidValue;
    MetricsArgs a;
    a.windows = windows; a.nq = nq; a.npred = max_pred_windows; a.gt = gt_windows; a.gt_off = gt_offsets; a.ranges = length_ranges; a.nr = n_ranges;
    a.thds = iou_thds; a.nt = n_thds; a.B = B; a.in_range = in_range; a.top1_iou = top1_iou; a.ap = ap;
    const long long n = (long long)n_ranges * B;
    mr_metrics_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    g_stats.launches++;
    return (int)cudaGetLastError();
}
