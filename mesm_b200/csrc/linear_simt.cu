// Generic fused linear layer on the fp32 SIMT pipe.
//
// Serves every row-wise Linear of the path whose shape does not suit the tcgen05 kernel (odd K / tiny M or N) and is
// the fp32 cross-check of the tensor-core kernel on the GPU (tests compare the two).  One CTA = 64 rows x 256 cols,
// 256 threads, 8x8 register tile per thread, K streamed in slabs of 16 through shared memory with register prefetch.
// The 256-wide tile holds a whole d_model row, so bias / activation / residual / LayerNorm run in the epilogue
// (LayerNorm statistics are a warp-shuffle reduction: one warp owns 8 complete rows).
#include "common.cuh"

namespace mesm {

thread_local LaunchStats g_stats;

namespace {

constexpr int BM = 64, BN = 256, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_PRELU) return v >= 0.f ? v : slope * v;
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

__global__ void __launch_bounds__(NT, 2) linear_simt_kernel(const LinearOp op_in) {
    LinearOp op = op_in;
    if (op.nbatch > 1) {           // per-head batch: shift the operand / output pointers
        const long long z = blockIdx.z;
        op.A += z * op.bsA; op.Wt += z * op.bsW; op.out += z * op.bsOut;
        if (op.bias) op.bias += z * op.bsBias;
    }
    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Ws[BK][BN];

    const int t = threadIdx.x;
    const int tx = t & 31, ty = t >> 5;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // A-load role: row a_r of the tile, 4 consecutive k starting at a_k
    const int a_r = t >> 2, a_k = (t & 3) * 4;
    const int a_m = m0 + a_r;
    const bool a_valid = a_m < op.M;
    const int n_lim = (op.N + 3) & ~3;      // weight columns are readable up to N rounded to the float4 width

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int pass = 0; pass < 2; ++pass) {
        const float* A = pass == 0 ? op.A : op.A2;
        if (A == nullptr) break;
        const int K = pass == 0 ? op.K : op.K2;
        const int lda = pass == 0 ? op.lda : op.lda2;
        const float* Wt = pass == 0 ? op.Wt : op.Wt2;
        const float* Apos = pass == 0 ? op.Apos : nullptr;
        const RowMap amap = pass == 0 ? op.amap : op.a2map;
        const long long a_off = a_valid ? amap(a_m) * (long long)lda : 0;
        const float* a_row = A + a_off;
        const float* p_row = Apos ? Apos + a_off : nullptr;
        const bool vec4 = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                          (!Apos || (reinterpret_cast<uintptr_t>(Apos) & 15) == 0);
        const int ktiles = (K + BK - 1) / BK;

        float a_reg[4];
        float4 w_reg[4];
        auto gload = [&](int kt) {
            const int k = kt * BK + a_k;
#pragma unroll
            for (int j = 0; j < 4; ++j) a_reg[j] = 0.f;
            if (a_valid) {
                if (vec4 && k + 3 < K) {
                    float4 v = *reinterpret_cast<const float4*>(a_row + k);
                    if (p_row) {
                        float4 p = *reinterpret_cast<const float4*>(p_row + k);
                        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
                    }
                    a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k + j < K) a_reg[j] = a_row[k + j] + (p_row ? p_row[k + j] : 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = t + i * NT;
                const int kk = idx >> 6, n = n0 + (idx & 63) * 4;
                w_reg[i] = (n < n_lim) ? *reinterpret_cast<const float4*>(Wt + (long long)(kt * BK + kk) * op.ldw + n)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto sstore = [&]() {
#pragma unroll
            for (int j = 0; j < 4; ++j) As[a_k + j][a_r] = a_reg[j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = t + i * NT;
                *reinterpret_cast<float4*>(&Ws[idx >> 6][(idx & 63) * 4]) = w_reg[i];
            }
        };

        gload(0);
        for (int kt = 0; kt < ktiles; ++kt) {
            __syncthreads();
            sstore();
            __syncthreads();
            if (kt + 1 < ktiles) gload(kt + 1);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
                const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
                const float4 w1 = *reinterpret_cast<const float4*>(&Ws[k][128 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const float slope = (op.act == ACT_PRELU) ? __ldg(op.prelu) : 0.f;
    int ncol[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ncol[j] = n0 + (j < 4 ? tx * 4 + j : 128 + tx * 4 + (j - 4));
    float bias[8], csum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool ok = ncol[j] < op.N;
        bias[j] = (op.bias && ok) ? __ldg(op.bias + ncol[j]) : 0.f;
        csum[j] = (op.colsum && ok) ? __ldg(op.colsum + ncol[j]) : 0.f;
    }
    float lg[8], lb[8];
    if (op.ln_g) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { lg[j] = __ldg(op.ln_g + ncol[j]); lb[j] = __ldg(op.ln_b + ncol[j]); }
    }
    const bool ovec = ((op.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.out) & 15) == 0);
    const bool rvec = op.residual && ((op.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.residual) & 15) == 0);

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        const bool mok = m < op.M;       // warp-uniform (ty is the warp index)
        float v[8];
        float mean = 0.f, rstd = 1.f;
        if (op.rowstat && mok) { mean = __ldg(op.rowstat + 2 * m); rstd = __ldg(op.rowstat + 2 * m + 1); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = acc[i][j];
            if (op.rowstat) x = rstd * (x - mean * csum[j]);
            x = (x + bias[j]) * op.out_scale;
            v[j] = apply_act(x, op.act, slope);
        }
        if (op.residual && mok) {
            const float* r = op.residual + op.rmap(m) * (long long)op.ldr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = ncol[h * 4];
                if (rvec && n + 3 < op.N) {
                    const float4 q = *reinterpret_cast<const float4*>(r + n);
                    v[h * 4] += q.x; v[h * 4 + 1] += q.y; v[h * 4 + 2] += q.z; v[h * 4 + 3] += q.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < op.N) v[h * 4 + j] += r[n + j];
                }
            }
        }
        if (op.pre_ln && mok) {
            float* o = op.pre_ln + (long long)m * op.N;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = ncol[h * 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) if (n + j < op.N) o[n + j] = v[h * 4 + j];
            }
        }
        if (op.ln_g) {  // N == 256: the warp holds the whole row
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
            const float mu = warp_sum(s) * (1.f / 256.f);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[j] - mu; q = fmaf(d, d, q); }
            const float rs = rsqrtf(warp_sum(q) * (1.f / 256.f) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (v[j] - mu) * rs * lg[j] + lb[j];
        }
        if (!mok) continue;
        for (int dst = 0; dst < 2; ++dst) {
            float* base = dst == 0 ? op.out : op.out2;
            if (!base) continue;
            const int ld = dst == 0 ? op.ldo : op.ldo2;
            float* o = base + (dst == 0 ? op.omap(m) : op.o2map(m)) * (long long)ld;
            const bool vec = dst == 0 ? ovec : (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = ncol[h * 4];
                if (vec && n + 3 < op.N) {
                    *reinterpret_cast<float4*>(o + n) = make_float4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < op.N) o[n + j] = v[h * 4 + j];
                }
            }
        }
    }
}

// N <= 8 outputs per row (class / span / anchor heads): one warp per row, K split over the lanes, weights from L1.
__global__ void __launch_bounds__(256) linear_smalln_kernel(const LinearOp op) {
    const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= op.M) return;
    const float* a = op.A + op.amap(m) * (long long)op.lda;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.f;
    for (int k = lane; k < op.K; k += 32) {
        const float x = a[k];
        const float* w = op.Wt + (long long)k * op.ldw;
#pragma unroll
        for (int n = 0; n < 8; ++n) if (n < op.N) acc[n] = fmaf(x, __ldg(w + n), acc[n]);
    }
    const float slope = (op.act == ACT_PRELU) ? __ldg(op.prelu) : 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (n >= op.N) break;
        float v = warp_sum(acc[n]);
        v = (v + (op.bias ? __ldg(op.bias + n) : 0.f)) * op.out_scale;
        v = apply_act(v, op.act, slope);
        if (lane == 0) op.out[op.omap(m) * (long long)op.ldo + n] = v;
    }
}

}  // namespace

cudaError_t launch_linear_simt(const LinearOp& op, cudaStream_t s) {
    if (op.M <= 0 || op.N <= 0) return cudaSuccess;
    if (op.ln_g && op.N != 256) return cudaErrorInvalidValue;
    if (op.fuse_rowstat) return cudaErrorInvalidValue;        // only the tcgen05 kernel computes the statistics in-kernel
    if ((op.ldw & 3) != 0) return cudaErrorInvalidValue;
    if (op.N <= 8 && !op.A2 && !op.Apos && !op.rowstat && !op.residual && !op.ln_g && !op.out2 && !op.pre_ln && op.nbatch <= 1) {
        linear_smalln_kernel<<<(op.M + 7) / 8, 256, 0, s>>>(op);
        g_stats.launches++;
        return cudaGetLastError();
    }
    dim3 grid((op.M + BM - 1) / BM, (op.N + BN - 1) / BN, op.nbatch > 1 ? op.nbatch : 1);
    linear_simt_kernel<<<grid, NT, 0, s>>>(op);
    g_stats.launches++;
    return cudaGetLastError();
}

}  // namespace mesm
