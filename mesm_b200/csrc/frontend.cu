// Feature-ingest front-end: what the reference's dataset classes do to a video's raw per-source clip features
// before the model sees them - get_video_feat (dataset/charades.py:108-119: fp32 upcast, per-source L2 normalisation over
// the feature dim, truncation to the shortest source, concatenation), sample_video_feat (dataset/base.py:100-114: mean-pool
// down-sampling to max_video_l clips) and add_tef (dataset/base.py:225-230: the two temporal-endpoint columns) - as two
// row-wise kernels on raw arrays stored in fp32 or fp16.  The result is written as fp32 or as fp16 (the 16-bit storage
// option mesm_forward consumes with video_feat_f16 = 1).
#include "common.cuh"
#include "../../include/mesm_b200.h"
#include <cuda_fp16.h>
#include <cmath>
#include <vector>

namespace mesm {

constexpr int kMaxSources = 4;
struct FrontendArgs {
    const void* raw[kMaxSources]; int dim[kMaxSources]; int off[kMaxSources]; float* inv[kMaxSources];
    int S, raw_f16, normalize, L_raw, L_out, Dtot, use_tef, out_f16, pooled;
    const int* idxs;            // [L_out + 1] window bounds when pooled
    void* out;
};

__device__ __forceinline__ float raw_at(const FrontendArgs& a, int s, long long i) {
    return a.raw_f16 ? __half2float(reinterpret_cast<const __half*>(a.raw[s])[i]) : reinterpret_cast<const float*>(a.raw[s])[i];
}

// inv[s][r] = 1 / max(||raw_s[r]||_2, 1e-12)   (F.normalize(x, dim=1), charades.py:114-115); one warp per (source, row)
__global__ void frontend_norm_kernel(const FrontendArgs a) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= a.S * a.L_raw) return;
    const int s = w / a.L_raw, r = w - s * a.L_raw;
    float ss = 0.f;
    for (int c = lane; c < a.dim[s]; c += 32) { const float v = raw_at(a, s, (long long)r * a.dim[s] + c); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) a.inv[s][r] = a.normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
}

// one CTA per output clip: mean of the normalised rows of its window (or the single row), then the tef columns
__global__ void frontend_rows_kernel(const FrontendArgs a) {
    const int i = blockIdx.x;
    int st = i, ed = i + 1;
    if (a.pooled) { st = a.idxs[i]; ed = a.idxs[i + 1]; if (!(st < ed)) ed = st + 1; }       // base.py:108-112
    const float inv_n = 1.f / (float)(ed - st);
    const int W = a.Dtot + (a.use_tef ? 2 : 0);
    for (int c = threadIdx.x; c < a.Dtot; c += blockDim.x) {
        int s = 0;
        while (s + 1 < a.S && c >= a.off[s + 1]) ++s;
        const int cc = c - a.off[s];
        float v;
        if (ed - st == 1) {
            v = raw_at(a, s, (long long)st * a.dim[s] + cc) * a.inv[s][st];
        } else {
            float acc = 0.f;
            for (int r = st; r < ed; ++r) acc += raw_at(a, s, (long long)r * a.dim[s] + cc) * a.inv[s][r];
            v = acc * inv_n;
        }
        if (a.out_f16) reinterpret_cast<__half*>(a.out)[(long long)i * W + c] = __float2half_rn(v);
        else reinterpret_cast<float*>(a.out)[(long long)i * W + c] = v;
    }
    if (a.use_tef && threadIdx.x == 0) {
        const float t0 = __fdiv_rn((float)i, (float)a.L_out), t1 = __fadd_rn(t0, __fdiv_rn(1.f, (float)a.L_out));
        if (a.out_f16) { __half* o = reinterpret_cast<__half*>(a.out) + (long long)i * W + a.Dtot; o[0] = __float2half_rn(t0); o[1] = __float2half_rn(t1); }
        else { float* o = reinterpret_cast<float*>(a.out) + (long long)i * W + a.Dtot; o[0] = t0; o[1] = t1; }
    }
}

}  // namespace mesm

using namespace mesm;

extern "C" int32_t mesm_video_feat_rows(const int32_t* raw_len, int32_t S, int32_t max_video_l) {
    if (!raw_len || S < 1) return 0;
    int m = raw_len[0];
    for (int s = 1; s < S; ++s) m = std::min(m, raw_len[s]);
    return m > max_video_l ? max_video_l : m;
}

extern "C" size_t mesm_video_feat_workspace_bytes(const int32_t* raw_len, int32_t S, int32_t max_video_l) {
    if (!raw_len || S < 1) return 0;
    int m = raw_len[0];
    for (int s = 1; s < S; ++s) m = std::min(m, raw_len[s]);
    return (size_t)S * m * sizeof(float) + ((size_t)max_video_l + 2) * sizeof(int) + 512;
}

extern "C" int mesm_build_video_feat(const void* const* raw, const int32_t* raw_len, const int32_t* dims, int32_t S, int32_t raw_f16,
                                     int32_t normalize, int32_t max_video_l, int32_t use_tef, void* out, int32_t out_f16, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!raw || !raw_len || !dims || !out || !workspace || S < 1 || S > kMaxSources || max_video_l < 1) return (int)cudaErrorInvalidValue;
    if (workspace_bytes < mesm_video_feat_workspace_bytes(raw_len, S, max_video_l)) return (int)cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    FrontendArgs a;
    int min_len = raw_len[0];
    for (int i = 1; i < S; ++i) min_len = std::min(min_len, (int)raw_len[i]);
    if (min_len < 1) return (int)cudaErrorInvalidValue;
    a.S = S; a.raw_f16 = raw_f16; a.normalize = normalize; a.L_raw = min_len; a.use_tef = use_tef; a.out_f16 = out_f16; a.out = out;
    a.pooled = min_len > max_video_l;
    a.L_out = a.pooled ? max_video_l : min_len;
    char* ws = (char*)workspace;
    int off = 0;
    for (int i = 0; i < S; ++i) {
        a.raw[i] = raw[i]; a.dim[i] = dims[i]; a.off[i] = off; off += dims[i];
        a.inv[i] = (float*)ws + (size_t)i * min_len;
    }
    a.Dtot = off;
    int* d_idx = (int*)(ws + (((size_t)S * min_len * sizeof(float) + 255) & ~(size_t)255));
    a.idxs = d_idx;
    if (a.pooled) {
        // idxs = (arange(0, max_l + 1, 1.0) / max_l * video_length).round().long().clamp(max = video_length - 1)   (base.py:103-104),
        // in the reference's fp32 arithmetic (torch.round = round-half-even = rintf)
        std::vector<int> h(max_video_l + 1);
        for (int i = 0; i <= max_video_l; ++i) {
            const float v = (float)i / (float)max_video_l * (float)min_len;
            long r = lrintf(v);
            h[i] = (int)std::min<long>(r, min_len - 1);
        }
        cudaError_t e = cudaMemcpyAsync(d_idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return (int)e;
        e = cudaStreamSynchronize(s);          // the host vector goes out of scope (front-end call: not on the scoring path)
        if (e != cudaSuccess) return (int)e;
    }
    frontend_norm_kernel<<<(S * min_len + 7) / 8, 256, 0, s>>>(a);
    frontend_rows_kernel<<<a.L_out, 256, 0, s>>>(a);
    g_stats.launches += 2;
    return (int)cudaGetLastError();
}
