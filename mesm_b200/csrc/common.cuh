// Shared declarations of the mesm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mesm {

constexpr int D = 256;      // hidden_dim of every shipped config (config/*/*.json: "hidden_dim": 256)
constexpr int NH = 8;       // nheads
constexpr int HD = 32;      // head dim
constexpr int FF = 1024;    // dim_feedforward

// Rows of a [groups, rows_per_group, C] tensor embedded in a buffer whose groups are `stride` rows apart and start
// `offset` rows in (e.g. the Lv clip rows inside the [B, Lv+1, 256] encoder buffer that carries the global token).
struct RowMap {
    int group;    // rows per group in the logical row index; 0 = identity
    int stride;   // rows between group starts in the buffer
    int offset;   // first row of the group in the buffer
    const int* table = nullptr;   // device gather table (row r lives at buffer row table[r]); overrides the affine map.
                                  // Used by the packed (variable-length) clip layout, where pairs have different row counts.
    __host__ __device__ inline long long operator()(int r) const {
#ifdef __CUDA_ARCH__
        if (table) return table[r];
#endif
        if (group == 0) return r;
        int g = r / group;
        return (long long)g * stride + offset + (r - g * group);
    }
};
__host__ __device__ static inline RowMap table_map(const int* t) { RowMap m{0, 0, 0}; m.table = t; return m; }
// Rows of pair b in a packed buffer.  cu = prefix sums of the per-pair clip counts, pointing at the first pair of the
// chunk (cu[b + 1] - cu[b] clips for pair b); enc = 1: encoder layout, every pair carries one extra leading row (the
// global token).  cu == nullptr: uniform layout, L rows per pair.
__device__ __forceinline__ void pair_rows(const int* __restrict__ cu, int enc, int b, int L, long long& start, int& len) {
    if (cu) {
        const int c0 = cu[b];
        start = (long long)(c0 - cu[0]) + (enc ? b : 0);
        len = cu[b + 1] - c0 + (enc ? 1 : 0);
    } else {
        start = (long long)b * L;
        len = L;
    }
}
__host__ __device__ static inline RowMap identity_map() { return RowMap{0, 0, 0}; }

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2, ACT_SIGMOID = 3 };

// out = epilogue( (A (+ Apos)) [M,K] . Wt[K,N]  (+ (A2)[M,K2] . Wt2[K2,N]) )
// epilogue order: LN-fold (rowstat/colsum) -> + bias -> * out_scale -> activation -> + residual -> LayerNorm(N)
struct LinearOp {
    int M, N, K;
    const float* A;   int lda;   RowMap amap;     // A[M,K] fp32 row-major (lda floats between rows)
    const float* Apos;                            // optional addend on A (same shape / lda / map), e.g. positional enc.
    const float* Wt;  int ldw;                    // weights pre-transposed: Wt[k*ldw + n], K rows zero-padded to /16
    const void* Wp; const void* Wp2;              // same weights as packed bf16 hi/lo tcgen05 tiles (null: SIMT only)
    int K2; const float* A2; int lda2; RowMap a2map; const float* Wt2;   // optional second input (same ldw)
    const float* bias;                            // [N] or null
    const float* rowstat;                         // [M,2] (mean, rstd) of A rows: LayerNorm folded into the GEMM
    int fuse_rowstat;                             // tcgen05 kernel only: compute (mean, rstd) of the A rows in-kernel instead
    const float* colsum;                          // [N]  sum_k Wt[k,n]  (weights already carry gamma)
    float out_scale;                              // applied after bias (attention q scaling); 1 = none
    int act; const float* prelu;                  // PReLU slope (1 element, device)
    const float* residual; int ldr; RowMap rmap;  // optional fp32 addend [M,N]
    const float* ln_g; const float* ln_b;         // optional LayerNorm over N (requires N == 256)
    float* out; int ldo; RowMap omap;
    float* out2; int ldo2; RowMap o2map;          // optional duplicate store of the final value
    float* pre_ln;                                // optional store of the value before LayerNorm ([M,N], ld = N)
    float* ln_stats;                              // linear_tma only: store (mean, rstd) of the pre-LayerNorm row ([M,2]) INSTEAD of normalising -
                                                  // the consumer (fused FFN) applies LayerNorm itself, so the normalised copy never touches HBM
    int nbatch;                                   // >1: blockIdx.z batches (per-head GEMMs); element strides below
    long long bsA, bsW, bsBias, bsOut;
    // Pre-split 16-bit planes: an activation x is stored as hi = bf16(x) and lo = bf16(x - hi), two row-major
    // [M, ld] planes - the form the tensor pipe consumes, so the TMA-fed GEMM (linear_tma.cu) needs no converter warps.
    const uint16_t* a_hi; const uint16_t* a_lo; int lda_p;    // A given as planes (A is then ignored); a_lo == nullptr: ONE fp16 plane (exact)
    uint16_t* out_hi; uint16_t* out_lo; int ldp;              // optional store of the final value as planes (row m -> row m)
    const uint16_t* res_hi; const uint16_t* res_lo;           // linear_tma only: the residual given as planes [M, N] (row m, pitch N) instead of fp32
    const void* Wtm;                                          // host object: the weights as TMA-addressable planes (TmaWeights, linear_tma.cu)
};

static inline LinearOp make_linear(int M, int N, int K, const float* A, int lda, const float* Wt, int ldw,
                                   const float* bias, float* out, int ldo) {
    LinearOp op;
    op.M = M; op.N = N; op.K = K; op.A = A; op.lda = lda; op.amap = identity_map(); op.Apos = nullptr;
    op.Wt = Wt; op.ldw = ldw; op.Wp = nullptr; op.Wp2 = nullptr; op.K2 = 0; op.A2 = nullptr; op.lda2 = 0; op.a2map = identity_map(); op.Wt2 = nullptr;
    op.bias = bias; op.rowstat = nullptr; op.fuse_rowstat = 0; op.colsum = nullptr; op.out_scale = 1.f; op.act = ACT_NONE; op.prelu = nullptr;
    op.residual = nullptr; op.ldr = 0; op.rmap = identity_map(); op.ln_g = nullptr; op.ln_b = nullptr;
    op.out = out; op.ldo = ldo; op.omap = identity_map(); op.out2 = nullptr; op.ldo2 = 0; op.o2map = identity_map(); op.pre_ln = nullptr; op.ln_stats = nullptr;
    op.nbatch = 1; op.bsA = op.bsW = op.bsBias = op.bsOut = 0;
    op.a_hi = op.a_lo = nullptr; op.lda_p = 0; op.out_hi = op.out_lo = nullptr; op.ldp = 0; op.Wtm = nullptr; op.res_hi = op.res_lo = nullptr;
    return op;
}

// Launch counter + optional per-launch CUDA-event timing of the fused-linear kernels (bench.py's roofline object).
struct LaunchStats {
    long long launches = 0;
    bool profile = false;
    double lin_ms = 0, lin_flops = 0, lin_bytes = 0; long long lin_launches = 0;      // collected totals
    double big_ms = 0, big_flops = 0; long long big_launches = 0;                     // launches with M >= 16384 only
};
extern thread_local LaunchStats g_stats;
const char* profile_report();
void profile_collect();     // synchronises the recorded events and folds them into g_stats
// RAII CUDA-event bracket around one launch (active only between mesm_profile_begin / _end)
struct ProfScope {
    cudaStream_t s; const char* name; double flops; int M; cudaEvent_t a; bool on;
    ProfScope(const char* name_, cudaStream_t s_, double flops_ = 0, int M_ = 0);
    ~ProfScope();
};

cudaError_t launch_linear(const LinearOp& op, cudaStream_t s);        // dispatcher (tcgen05 when eligible)
cudaError_t launch_linear_simt(const LinearOp& op, cudaStream_t s);   // fp32 SIMT kernel
cudaError_t launch_linear_tc(const LinearOp& op, cudaStream_t s);     // tcgen05 split-bf16 kernel
bool linear_tc_eligible(const LinearOp& op);
// TMA-fed persistent tcgen05 kernel (linear_tma.cu): A and W as 16-bit planes, no operand conversion in the kernel
cudaError_t launch_linear_tma(const LinearOp& op, cudaStream_t s);
bool linear_tma_eligible(const LinearOp& op);
// weights as planes + their tensor maps; fp16 = single-plane-A mode (exact fp16 activations x fp16 hi/lo weights, 2 MMAs)
void* tma_pack_weights(const float* W, int row0, int nrows, int K, const float* gamma, bool fp16, cudaStream_t s);   // host object, free with tma_free_weights
void tma_free_weights(void* w);
cudaError_t launch_f32_to_f16(const float* x, long long rows, int cols, int ldx, uint16_t* y, int ldy, cudaStream_t s);
cudaError_t launch_merge_planes(const uint16_t* hi, const uint16_t* lo, long long n, float* out, cudaStream_t s);     // out = hi + lo (bf16 planes)
cudaError_t launch_split_planes(const float* x, long long rows, int cols, int ldx, uint16_t* hi, uint16_t* lo, int ldp, cudaStream_t s);
size_t tc_packed_bytes(int nrows, int K);
cudaError_t launch_pack_tc(const float* W, int row0, int nrows, int K, const float* gamma, void* out, cudaStream_t s);

#define MESM_CHECK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) return _e;                                                        \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace mesm
