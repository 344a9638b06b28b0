// TMA-fed persistent linear layer on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   out[M,N] = epilogue( A[M,K] . W[N,K]^T )        A and W given as 16-bit PLANES, accumulation in fp32 (TMEM)
//
// Round 1's linear_tc_kernel converts its fp32 A operand to bf16 hi/lo inside the kernel: 8 converter warps, ~150 dependent
// instructions per 32-wide K block, a tile that waits ~15 k cycles for its first MMA and a tensor pipe that idles two thirds of
// the time (profiles/r1_linear_tc_ncu.md).  Here the producers of an activation store it pre-split (hi = bf16(x),
// lo = bf16(x - hi): the same 4 bytes per element as fp32) and this kernel moves BOTH operands with the TMA engine:
//   warp 0      : one thread issues cp.async.bulk.tensor loads of the A planes (box 32 x 128, SWIZZLE_64B) and of the weight
//                 planes (box 32 x 256) into a 3-stage ring
//   warp 1      : one thread issues the MMAs (M128 N256 K16; bf16x3: Ahi.Whi + Alo.Whi + Ahi.Wlo; or, for an exact single
//                 fp16 A plane - the 16-bit stored clip features - A.Whi + A.Wlo) and commits stages / accumulators
//   warps 2..9  : epilogue out of TMEM (LayerNorm fold / bias / scale / activation / residual / LayerNorm / stores as fp32
//                 and/or as planes), identical arithmetic to linear_tc_kernel
// The CTA is persistent (one per SM) and the accumulator is double-buffered in TMEM (2 x 256 columns): the epilogue of tile i
// runs under the K loop of tile i+1, and the operand ring runs ahead across tile boundaries.
#include "tc_common.cuh"
#include "tma_host.h"

#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>

namespace mesm {
namespace tma {
using namespace tc;

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int A_TILE = BM * BK * 2;                       // 8 KB: one plane of the A block
constexpr int W_TILE = BN * BK * 2;                       // 16 KB: one plane of the weight block
// operand ring: 3 stages of 48 KB (two A planes + two W planes) for split bf16 operands, 4 stages of 40 KB for a single fp16 A plane
constexpr int RING_BYTES = 4 * (A_TILE + 2 * W_TILE);     // 160 KB (>= 3 * 48 KB)
constexpr int NEPI = 256;                                 // epilogue threads (8 warps)
constexpr int THREADS = 64 + NEPI;
constexpr int OFF_T = RING_BYTES;                         // per-warp 32 x 36 fp32 transpose scratch
constexpr int T_BYTES = 8 * 32 * 36 * 4;
constexpr int OFF_BAR = OFF_T + T_BYTES;                  // full[4] empty[4] tfull[2] tempty[2] tmem_ptr
constexpr int OFF_LNX = OFF_BAR + 128;                    // [2][128]
constexpr int OFF_VEC = OFF_LNX + 1024;                   // [4][256]: bias, colsum, ln_g, ln_b of the N tile
constexpr int OFF_ROWOFF = OFF_VEC + 4096;                // [3][128] long long: out, out2, residual row offsets
constexpr int SMEM_BYTES = OFF_ROWOFF + 3072 + 1024;      // + alignment slack

struct TmaParams {
    LinearOp op;
    CUtensorMap tmA0, tmA1, tmW0, tmW1;
    int nkb, mtiles, ntn, npass;                          // npass 3: split bf16 A; 2: one exact fp16 A plane
    uint32_t idesc;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) linear_tma_kernel(const __grid_constant__ TmaParams P) {
    const LinearOp& op = P.op;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_full = smem_base + OFF_BAR, bar_empty = bar_full + 32, bar_tfull = bar_full + 64, bar_tempty = bar_full + 80;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);
    const int NST = P.npass == 3 ? 3 : 4;                 // ring depth / stage size of this launch's operand format
    const int A_BYTES = P.npass == 3 ? 2 * A_TILE : A_TILE, STG = A_BYTES + 2 * W_TILE;
    float* ln_x = reinterpret_cast<float*>(smem + OFF_LNX);
    float* vec_s = reinterpret_cast<float*>(smem + OFF_VEC);
    long long* rowoff = reinterpret_cast<long long*>(smem + OFF_ROWOFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = P.mtiles * P.ntn;
    const int nmine = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nkb = P.nkb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, NEPI / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&P.tmA0); tma_prefetch_desc(&P.tmW0); tma_prefetch_desc(&P.tmW1);
        if (P.npass == 3) tma_prefetch_desc(&P.tmA1);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // Programmatic dependent launch: the next kernel may start its set-up once every CTA of this grid is resident; this kernel
    // reads what its predecessor wrote only after griddepcontrol.wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================== TMA producer (lane 0) =====================
        int g = 0;
        for (int ti = 0; ti < nmine; ++ti) {
            const int t = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mt = t / P.ntn, nt = t - mt * P.ntn;
            const int m0 = mt * BM, n0 = nt * BN;
            if (lane == 0) {
                // every tile walks the K blocks in its own rotation: the CTAs start together and would otherwise all ask L2
                // for the same weight lines at the same moment
                const int krot = mt % nkb;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    const uint32_t ph = (g / NST) & 1;
                    mbar_wait_spin(bar_empty + 8 * s, ph ^ 1, 1000 + kb);
                    int kk = kb + krot; kk = kk >= nkb ? kk - nkb : kk;
                    const uint32_t dst = smem_base + s * STG, full = bar_full + 8 * s;
                    mbar_arrive_expect_tx(full, STG);
                    tma_load_2d(dst, &P.tmA0, kk * BK, m0, full);
                    if (P.npass == 3) tma_load_2d(dst + A_TILE, &P.tmA1, kk * BK, m0, full);
                    tma_load_2d(dst + A_BYTES, &P.tmW0, kk * BK, n0, full);
                    tma_load_2d(dst + A_BYTES + W_TILE, &P.tmW1, kk * BK, n0, full);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            int g = 0;
            for (int ti = 0; ti < nmine; ++ti) {
                const int acc = ti & 1;
                const uint32_t aph = (ti >> 1) & 1;
                mbar_wait_spin(bar_tempty + 8 * acc, aph ^ 1, 2500 + ti);        // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    const uint32_t ph = (g / NST) & 1;
                    mbar_wait_spin(bar_full + 8 * s, ph, 2000 + kb);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + s * STG, a_lo = a_hi + A_TILE;
                    const uint32_t w_hi = a_hi + A_BYTES, w_lo = w_hi + W_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t koff = k * 32;          // 16 elements = 32 bytes along K inside the 64-byte swizzle row
                        const uint64_t dah = make_desc(a_hi + koff), dwh = make_desc(w_hi + koff), dwl = make_desc(w_lo + koff);
                        umma(tacc, dah, dwh, (kb > 0 || k > 0) ? 1u : 0u, P.idesc);
                        if (P.npass == 3) umma(tacc, make_desc(a_lo + koff), dwh, 1u, P.idesc);
                        umma(tacc, dah, dwl, 1u, P.idesc);
                    }
                    umma_commit(bar_empty + 8 * s);            // stage reusable once these MMAs retire
                }
                umma_commit(bar_tfull + 8 * acc);              // accumulator complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int tc = threadIdx.x - 64;                   // 0..255
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                  // column half handled by this warp
        const int row = q * 32 + lane;
        const float slope_eff = op.act == ACT_PRELU ? __ldg(op.prelu) : (op.act == ACT_RELU ? 0.f : 1.f);
        const bool do_ln = op.ln_g != nullptr || op.ln_stats != nullptr;
        const bool stats_only = op.ln_stats != nullptr;      // (mean, rstd) of the pre-LN row for the consumer; nothing normalised here
        float* T = reinterpret_cast<float*>(smem + OFF_T) + (warp - 2) * (32 * 36);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int trow0 = q * 32;

        for (int ti = 0; ti < nmine; ++ti) {
            const int t = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mt = t / P.ntn, nt = t - mt * P.ntn;
            const int m0 = mt * BM, n0 = nt * BN;
            const int acc = ti & 1;
            const uint32_t aph = (ti >> 1) & 1;
            asm volatile("bar.sync 1, 256;" ::: "memory");           // every warp is done with the previous tile's tables
            {   // per-column epilogue vectors of this N tile and the row offsets of this M tile
                const int n = n0 + tc;
                const bool nok = n < op.N;
                vec_s[tc] = (op.bias && nok) ? __ldg(op.bias + n) : 0.f;
                vec_s[256 + tc] = (op.colsum && nok) ? __ldg(op.colsum + n) : 0.f;
                vec_s[512 + tc] = (op.ln_g && nok) ? __ldg(op.ln_g + n) : 0.f;
                vec_s[768 + tc] = (op.ln_b && nok) ? __ldg(op.ln_b + n) : 0.f;
                if (tc < 128) {
                    const int m = m0 + tc;
                    const bool ok = m < op.M;
                    rowoff[tc] = (ok && op.out) ? op.omap(m) * (long long)op.ldo : -1;
                    rowoff[128 + tc] = (ok && op.out2) ? op.o2map(m) * (long long)op.ldo2 : -1;
                    rowoff[256 + tc] = (ok && op.residual) ? op.rmap(m) * (long long)op.ldr : -1;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int m = m0 + row;
            const bool mok = m < op.M;
            if (op.residual || op.res_hi) {
                // this thread's residual row segment (its TMEM lane x this warp's column half) into L2 now: the loads of
                // rows_pass follow within a few thousand cycles.  (Prefetching from the producer warp, two tiles ahead as
                // linear_tc does, was useless here: ~150 MB stream through L2 in between and the lines were fetched twice -
                // ncu: 920 MB of DRAM reads for 614 MB of operands.)
                if (op.res_hi && mok) {
                    const uint16_t* rh = op.res_hi + (long long)m * op.N + n0 + half * 128;
                    const uint16_t* rl = op.res_lo + (long long)m * op.N + n0 + half * 128;
                    prefetch_l2(rh); prefetch_l2(rh + 64); prefetch_l2(rl); prefetch_l2(rl + 64);
                }
                const long long ro = rowoff[256 + row];
                if (ro >= 0) {
                    const float* r = op.residual + ro + n0 + half * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (n0 + half * 128 + c * 32 < op.N) prefetch_l2(r + c * 32);
                }
            }
            float mean_in = 0.f, rstd_in = 1.f;
            if (op.rowstat && mok) { mean_in = __ldg(op.rowstat + 2 * m); rstd_in = __ldg(op.rowstat + 2 * m + 1); }
            if (lane == 0) mbar_wait(bar_tfull + 8 * acc, aph, 5000 + ti);
            __syncwarp();
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + half * 128;

            // transposed pass over the staged 32 x 32 chunk starting at column n: each instruction moves 4 rows x 128 bytes
            auto rows_pass = [&](int n, bool add_res, bool keep_in_T, bool final_store, float* dstp) {
                const int nn = n + c4;
                const bool nok = nn < op.N;
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(&T[(4 * i + rsub) * 36 + c4]);
                if (add_res) {
                    float4 r[8];
                    if (op.res_hi) {                        // residual stored pre-split: r = hi + lo (16 mantissa bits, the precision it is consumed at anyway)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int mm = m0 + trow0 + 4 * i + rsub;
                            uint2 h = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
                            if (mm < op.M && nok) {
                                h = __ldg(reinterpret_cast<const uint2*>(op.res_hi + (long long)mm * op.N + nn));
                                l = __ldg(reinterpret_cast<const uint2*>(op.res_lo + (long long)mm * op.N + nn));
                            }
                            r[i].x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
                            r[i].y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
                            r[i].z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
                            r[i].w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
                        }
                    } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long long o = rowoff[2 * 128 + trow0 + 4 * i + rsub];
                        r[i] = (o >= 0 && nok) ? __ldg(reinterpret_cast<const float4*>(op.residual + o + nn)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x[i].x += r[i].x; x[i].y += r[i].y; x[i].z += r[i].z; x[i].w += r[i].w; }
                }
                if (keep_in_T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&T[(4 * i + rsub) * 36 + c4]) = x[i];
                }
                if (!nok) return;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int tr = trow0 + 4 * i + rsub;
                    if (final_store) {
                        const long long o0 = rowoff[tr], o1 = rowoff[128 + tr];
                        if (o0 >= 0) *reinterpret_cast<float4*>(op.out + o0 + nn) = x[i];
                        if (o1 >= 0) *reinterpret_cast<float4*>(op.out2 + o1 + nn) = x[i];
                        if (op.out_hi && m0 + tr < op.M) {
                            uint2 h, l;
                            split_bf16x2(x[i].x, x[i].y, h.x, l.x);
                            split_bf16x2(x[i].z, x[i].w, h.y, l.y);
                            const long long o = (long long)(m0 + tr) * op.ldp + nn;
                            *reinterpret_cast<uint2*>(op.out_hi + o) = h;
                            *reinterpret_cast<uint2*>(op.out_lo + o) = l;
                        }
                    }
                    if (dstp && m0 + tr < op.M) *reinterpret_cast<float4*>(dstp + (long long)(m0 + tr) * op.N + nn) = x[i];
                }
            };

            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int n = n0 + half * 128 + c * 32;
                if (n >= op.N) break;                           // warp-uniform
                float v[32];
                tmem_ld32(taddr0 + c * 32, v);
                {   // LN-fold (identity when unused), bias, scale, leaky activation (slope_eff)
                    const int cl0 = n - n0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&vec_s[cl0 + 4 * j]);
                        const float4 c4v = *reinterpret_cast<const float4*>(&vec_s[256 + cl0 + 4 * j]);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, cc[4] = {c4v.x, c4v.y, c4v.z, c4v.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float x = rstd_in * fmaf(-mean_in, cc[u], v[4 * j + u]);
                            x = (x + bb[u]) * op.out_scale;
                            v[4 * j + u] = x >= 0.f ? x : slope_eff * x;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(&T[lane * 36 + 4 * j]) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                if (do_ln) {
                    // residual added and pre-LN value stored in the transposed (coalesced) domain; result kept for the stats
                    rows_pass(n, op.residual != nullptr || op.res_hi != nullptr, true, false, op.pre_ln);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 tt = *reinterpret_cast<const float4*>(&T[lane * 36 + 4 * j]);
                        v[4 * j] = tt.x; v[4 * j + 1] = tt.y; v[4 * j + 2] = tt.z; v[4 * j + 3] = tt.w;
                        sum += (tt.x + tt.y) + (tt.z + tt.w);
                    }
                    tmem_st32(taddr0 + c * 32, v);
                } else {
                    rows_pass(n, op.residual != nullptr || op.res_hi != nullptr, false, true, nullptr);
                }
                __syncwarp();
            }
            if (do_ln) {                                       // N == 256: this thread holds half of row `row`
                ln_x[half * 128 + row] = sum;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float mu = (ln_x[row] + ln_x[128 + row]) * (1.f / 256.f);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                float sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(taddr0 + c * 32, v);
                    if (stats_only && c == 3) {                // last TMEM read of this accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; sq = fmaf(d, d, sq); }
                }
                ln_x[half * 128 + row] = sq;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float rs = rsqrtf((ln_x[row] + ln_x[128 + row]) * (1.f / 256.f) + 1e-5f);
                if (stats_only) {
                    if (half == 0 && mok) { op.ln_stats[2 * (long long)m] = mu; op.ln_stats[2 * (long long)m + 1] = rs; }
                    continue;                                  // next tile (the top-of-loop barrier protects ln_x)
                }
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int n = n0 + half * 128 + c * 32;
                    float v[32];
                    tmem_ld32(taddr0 + c * 32, v);
                    if (c == 3) {                              // last TMEM read of this accumulator: hand it back to the MMA thread
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int cl = n - n0 + j;
                        v[j] = (v[j] - mu) * rs * vec_s[512 + cl] + vec_s[768 + cl];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(&T[lane * 36 + 4 * j]) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    rows_pass(n, false, false, true, nullptr);
                    __syncwarp();
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- fp32 rows -> bf16 hi / lo planes (tests and producers that still emit fp32) -------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, uint16_t* __restrict__ hi,
                                    uint16_t* __restrict__ lo, int ldp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 2 elements
    const int c2 = cols >> 1;
    if (i >= rows * c2) return;
    const long long r = i / c2;
    const int c = (int)(i - r * c2) * 2;
    const float2 v = *reinterpret_cast<const float2*>(x + r * ldx + c);
    uint32_t h, l;
    split_bf16x2(v.x, v.y, h, l);
    *reinterpret_cast<uint32_t*>(hi + r * ldp + c) = h;
    *reinterpret_cast<uint32_t*>(lo + r * ldp + c) = l;
}

__global__ void f32_to_f16_kernel(const float* __restrict__ x, long long rows, int cols, int ldx, uint16_t* __restrict__ y, int ldy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    y[r * ldy + c] = __half_as_ushort(__float2half_rn(x[r * ldx + c]));
}

__global__ void merge_planes_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __uint_as_float((uint32_t)hi[i] << 16) + __uint_as_float((uint32_t)lo[i] << 16);
}

// ---- weights: fp32 W[N,K] rows [row0,row0+nrows) (x gamma[k]) -> hi / lo planes [nrows_p, Kp] (bf16, or fp16) -------
template <bool FP16>
__global__ void pack_planes_kernel(const float* __restrict__ W, int row0, int nrows, int K, const float* __restrict__ gamma,
                                   uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int nrows_p, int Kp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nrows_p * Kp) return;
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    float w = 0.f;
    if (n < nrows && k < K) w = W[(long long)(row0 + n) * K + k] * (gamma ? gamma[k] : 1.f);
    if (FP16) {
        const __half h = __float2half_rn(w);
        const __half l = __float2half_rn(w - __half2float(h));
        hi[i] = __half_as_ushort(h); lo[i] = __half_as_ushort(l);
    } else {
        __nv_bfloat16 h, l;
        split_bf16(w, h, l);
        hi[i] = __bfloat16_as_ushort(h); lo[i] = __bfloat16_as_ushort(l);
    }
}

}  // namespace tma

struct TmaWeights {
    uint16_t* planes = nullptr;      // device: hi [Np, Kp] then lo [Np, Kp]
    int N = 0, Np = 0, K = 0, Kp = 0;
    bool fp16 = false;
    CUtensorMap tm_hi, tm_lo;
};

void* tma_pack_weights(const float* W, int row0, int nrows, int K, const float* gamma, bool fp16, cudaStream_t s) {
    if (!tma_encode_fn()) return nullptr;
    TmaWeights* w = new TmaWeights();
    w->N = nrows; w->Np = (nrows + tma::BN - 1) / tma::BN * tma::BN; w->K = K; w->Kp = (K + tma::BK - 1) / tma::BK * tma::BK; w->fp16 = fp16;
    const size_t plane = (size_t)w->Np * w->Kp;
    if (cudaMalloc((void**)&w->planes, 2 * plane * sizeof(uint16_t)) != cudaSuccess) { delete w; return nullptr; }
    const long long tot = (long long)plane;
    if (fp16) tma::pack_planes_kernel<true><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W, row0, nrows, K, gamma, w->planes, w->planes + plane, w->Np, w->Kp);
    else tma::pack_planes_kernel<false><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(W, row0, nrows, K, gamma, w->planes, w->planes + plane, w->Np, w->Kp);
    g_stats.launches++;
    const bool ok = tma_map_2d_16bit(&w->tm_hi, w->planes, w->Kp, w->Np, (unsigned long long)w->Kp * 2, tma::BK, tma::BN, CU_TENSOR_MAP_SWIZZLE_64B, fp16) &&
                    tma_map_2d_16bit(&w->tm_lo, w->planes + plane, w->Kp, w->Np, (unsigned long long)w->Kp * 2, tma::BK, tma::BN, CU_TENSOR_MAP_SWIZZLE_64B, fp16);
    if (!ok || cudaGetLastError() != cudaSuccess) { cudaFree(w->planes); delete w; return nullptr; }
    return w;
}

void tma_free_weights(void* p) {
    TmaWeights* w = static_cast<TmaWeights*>(p);
    if (!w) return;
    cudaFree(w->planes);
    delete w;
}

cudaError_t launch_split_planes(const float* x, long long rows, int cols, int ldx, uint16_t* hi, uint16_t* lo, int ldp, cudaStream_t s) {
    ProfScope _ps("split_planes", s);
    if (rows <= 0) return cudaSuccess;
    if ((cols & 1) || (ldx & 1) || (ldp & 1)) return cudaErrorInvalidValue;
    const long long tot = rows * (cols >> 1);
    tma::split_planes_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(x, rows, cols, ldx, hi, lo, ldp);
    g_stats.launches++;
    return cudaGetLastError();
}

cudaError_t launch_f32_to_f16(const float* x, long long rows, int cols, int ldx, uint16_t* y, int ldy, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const long long tot = rows * cols;
    tma::f32_to_f16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(x, rows, cols, ldx, y, ldy);
    g_stats.launches++;
    return cudaGetLastError();
}

cudaError_t launch_merge_planes(const uint16_t* hi, const uint16_t* lo, long long n, float* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    tma::merge_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hi, lo, n, out);
    g_stats.launches++;
    return cudaGetLastError();
}

bool linear_tma_eligible(const LinearOp& op) {
    if (!op.Wtm || !op.a_hi || op.nbatch > 1 || op.Apos || op.A2 || op.fuse_rowstat) return false;
    const TmaWeights* w = static_cast<const TmaWeights*>(op.Wtm);
    if (w->K != op.K || w->N != op.N || (op.N % tma::BN) != 0) return false;
    if (w->fp16 != (op.a_lo == nullptr)) return false;                 // single fp16 plane <-> fp16 weight planes
    if (op.M < 1 || op.amap.group != 0 || op.amap.table != nullptr) return false;
    if ((op.ln_g || op.ln_stats) && op.N != 256) return false;
    if (op.ln_stats && (op.ln_g || !op.pre_ln)) return false;          // stats-only: the pre-LN rows are the only output
    if (op.act == ACT_SIGMOID) return false;
    if ((op.lda_p & 7) || (reinterpret_cast<uintptr_t>(op.a_hi) & 15) || (op.a_lo && (reinterpret_cast<uintptr_t>(op.a_lo) & 15))) return false;
    auto al16 = [](const void* p, long long ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0)); };
    if (!al16(op.out, op.ldo) || !al16(op.out2, op.ldo2) || !al16(op.residual, op.ldr) || !al16(op.pre_ln, op.N)) return false;
    if (op.res_hi && (op.residual || !op.res_lo || (op.N & 3) || (reinterpret_cast<uintptr_t>(op.res_hi) & 7) || (reinterpret_cast<uintptr_t>(op.res_lo) & 7))) return false;
    if (op.out_hi && ((op.ldp & 3) || (reinterpret_cast<uintptr_t>(op.out_hi) & 7) || (reinterpret_cast<uintptr_t>(op.out_lo) & 7))) return false;
    if (!op.out && !op.out_hi && !op.ln_stats) return false;
    return true;
}

cudaError_t launch_linear_tma(const LinearOp& op, cudaStream_t s) {
    static int nsm = 0, pdl = -1;
    if (!nsm) {
        int dev = 0;
        MESM_CHECK(cudaGetDevice(&dev));
        MESM_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        MESM_CHECK(cudaFuncSetAttribute(tma::linear_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tma::SMEM_BYTES));
        const char* pe = getenv("MESM_TC_PDL");
        pdl = (pe && pe[0] == '0') ? 0 : 1;
    }
    const TmaWeights* w = static_cast<const TmaWeights*>(op.Wtm);
    tma::TmaParams P;
    P.op = op;
    const bool single = op.a_lo == nullptr;
    if (!tma_map_2d_16bit(&P.tmA0, op.a_hi, (unsigned long long)op.K, (unsigned long long)op.M, (unsigned long long)op.lda_p * 2, tma::BK, tma::BM,
                          CU_TENSOR_MAP_SWIZZLE_64B, single))
        return cudaErrorInvalidValue;
    if (!single) {
        if (!tma_map_2d_16bit(&P.tmA1, op.a_lo, (unsigned long long)op.K, (unsigned long long)op.M, (unsigned long long)op.lda_p * 2, tma::BK, tma::BM,
                              CU_TENSOR_MAP_SWIZZLE_64B, false))
            return cudaErrorInvalidValue;
    } else {
        P.tmA1 = P.tmA0;
    }
    P.tmW0 = w->tm_hi; P.tmW1 = w->tm_lo;
    P.nkb = w->Kp / tma::BK; P.mtiles = (op.M + tma::BM - 1) / tma::BM; P.ntn = op.N / tma::BN; P.npass = single ? 2 : 3;
    // kind::f16 instruction descriptor: D = f32; A = B = bf16 (format 1) or fp16 (format 0); M = 128, N = 256
    P.idesc = single ? ((1u << 4) | ((uint32_t)(tma::BN >> 3) << 17) | ((uint32_t)(tma::BM >> 4) << 24)) : tc::make_idesc(tma::BN);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(P.mtiles * P.ntn, nsm), 1, 1);
    cfg.blockDim = dim3(tma::THREADS, 1, 1);
    cfg.dynamicSmemBytes = tma::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, tma::linear_tma_kernel, P);
    g_stats.launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mesm
