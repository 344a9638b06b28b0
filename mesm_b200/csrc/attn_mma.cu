// Multi-head attention (head_dim 32) for the encoder self-attention (keys = the pair's clips + global token) and the T2V
// cross-attention (keys = <= 33 words, with the reference's attn_mask quirk): model/transformer.py T2V_TransformerEncoderLayer /
// TransformerEncoderLayer -> model/attention.py multi_head_attention_forward (:185-394).
//
// Warp-level tensor-core kernels.  The work per (pair, head) is tiny (<= 224 x 224 x 32) and HBM-bound (q, k, v in, one result row
// out: 4 KB per row); what the tcgen05 kernels in attn_tc.cu lost was LATENCY: one 13-warp CTA per SM walking strictly ordered
// phases (stage operands -> MMA -> TMEM -> softmax -> P to shared memory -> MMA -> TMEM -> store) at 11 % tensor-pipe and 15 % DRAM
// utilisation.  Here every warp owns 16 query rows end to end and never talks to another warp while it computes:
//     S = Q K^T       mma.sync m16n8k16, bf16x3 split operands (Qhi Khi + Qlo Khi + Qhi Klo), fp32 scores in registers
//     online softmax  32 or 64 keys per step (register budget of the variant), exp2 of the scores scaled by log2(e) / sqrt(32)
//     O += P V        the score fragments ARE the A fragments of the next MMA (no shared-memory round trip), bf16x3 again
// K / V of one head sit in shared memory as bf16 hi / lo planes, [key][32 dims] = 64-byte rows with the 16-byte chunks XOR-swizzled
// by (key >> 1) & 3 so that ldmatrix (K) and ldmatrix.trans (V) are conflict-free without padding.
//
//   attn_mma_kernel        encoder self-attention: one CTA per (pair, head) converts K / V of the head from fp32 itself (feeding the
//                          kernel pre-split planes through cp.async was measured SLOWER: 1549 vs 1307 us per 4096 x 147-row launch -
//                          64-byte plane segments instead of 128-byte fp32 lines - and a persistent double-buffered variant slower
//                          still, 2772 us: each (pair, head) item is too short to hide its own Q loads)
//   attn_mma_heads_kernel  T2V cross-attention (<= 64 keys): one CTA per pair, warp = head
//   dec_cross_mma_kernel   decoder cross- / self-attention (<= 16 queries): B fragments straight from global memory
#include "kernels.h"
#include "tc_common.cuh"
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>

namespace mesm {
namespace am {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {      // c = a b
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk c (8 dims) of key row `key` inside one plane
__device__ __forceinline__ int swz(int key, int c) { return key * 64 + ((c ^ ((key >> 1) & 3)) << 4); }

// per-lane constants of the shared-memory operand reads
struct LaneAddr {
    uint32_t k, v;          // ldmatrix row addresses of this lane inside a plane (see tile_attention)
};
__device__ __forceinline__ LaneAddr lane_addr(int lane) {
    LaneAddr la;
    // K (B operand of S, "col" = [key][dim]): matrix i = dims 8i..8i+7 of keys kb..kb+7.
    // V (B operand of O through .trans): matrix i = keys k0 + 8 (i & 1) .. +7, dims 8 (2 dp + (i >> 1)) .. +7; dp = 1 is ^ 32.
    la.k = (uint32_t)((lane & 7) * 64 + (((lane >> 3) ^ ((lane & 7) >> 1)) << 4));
    la.v = (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * 64 + (((lane >> 4) ^ ((lane & 7) >> 1)) << 4));
    return la;
}

struct TileState {
    float o[4][4];
    float m0, m1, l0, l1;   // running row maxima (unscaled scores) and row sums of rows g / g + 8
};

// NT n-tiles (8 keys each, NT even) starting at key c0: S = Q K^T, masks, online-softmax update, O += P V
template <int NT>
__device__ __forceinline__ void attn_chunk(TileState& st, const uint32_t (&qh)[2][4], const uint32_t (&ql)[2][4], uint32_t sK, uint32_t sV,
                                           uint32_t plane_bytes, const uint8_t* flags, bool any_flag, int c0, bool rf0, bool rf1,
                                           float scale, const LaneAddr& la, int t) {
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
        uint32_t kh[2][4], kl[2][4];
        const uint32_t off = sK + (uint32_t)((c0 + 8 * j) * 64) + la.k;
        ldsm4(off, kh[0]); ldsm4(off + 512, kh[1]);
        ldsm4(off + plane_bytes, kl[0]); ldsm4(off + plane_bytes + 512, kl[1]);
        // the two n-tiles alternate so that an MMA never waits for the accumulator of the one issued just before it
        mma16816z(s[j], qh[0], kh[0][0], kh[0][1]);
        mma16816z(s[j + 1], qh[0], kh[1][0], kh[1][1]);
        mma16816(s[j], ql[0], kh[0][0], kh[0][1]);
        mma16816(s[j + 1], ql[0], kh[1][0], kh[1][1]);
        mma16816(s[j], qh[0], kl[0][0], kl[0][1]);
        mma16816(s[j + 1], qh[0], kl[1][0], kl[1][1]);
        mma16816(s[j], qh[1], kh[0][2], kh[0][3]);
        mma16816(s[j + 1], qh[1], kh[1][2], kh[1][3]);
        mma16816(s[j], ql[1], kh[0][2], kh[0][3]);
        mma16816(s[j + 1], ql[1], kh[1][2], kh[1][3]);
        mma16816(s[j], qh[1], kl[0][2], kl[0][3]);
        mma16816(s[j + 1], qh[1], kl[1][2], kl[1][3]);
    }
    if (any_flag) {                                   // warp-uniform: some key of this chunk is masked (padding, tail, quirk)
#pragma unroll
        for (int j = 0; j < NT; ++j) {                // this lane's keys of n-tile j are c0 + 8 j + 2 t, + 1: two flag bytes
            const uint32_t f = *reinterpret_cast<const uint16_t*>(flags + c0 + 8 * j + 2 * t);
            const bool ma = f & 0x001u, mb = f & 0x100u, qa = f & 0x002u, qb = f & 0x200u;
            if (ma || (rf0 && qa)) s[j][0] = -CUDART_INF_F;
            if (mb || (rf0 && qb)) s[j][1] = -CUDART_INF_F;
            if (ma || (rf1 && qa)) s[j][2] = -CUDART_INF_F;
            if (mb || (rf1 && qb)) s[j][3] = -CUDART_INF_F;
        }
    }
    float cm0 = fmaxf(s[0][0], s[0][1]), cm1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int j = 1; j < NT; ++j) { cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1])); cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3])); }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float mn0 = fmaxf(st.m0, cm0), mn1 = fmaxf(st.m1, cm1);
    const float mu0 = mn0 == -CUDART_INF_F ? 0.f : mn0, mu1 = mn1 == -CUDART_INF_F ? 0.f : mn1;   // every key so far masked: p = 0
    const float sc0 = tc::ex2_approx((st.m0 - mu0) * scale), sc1 = tc::ex2_approx((st.m1 - mu1) * scale);
    const float ms0 = mu0 * scale, ms1 = mu1 * scale;
    st.m0 = mn0; st.m1 = mn1;
    st.l0 *= sc0; st.l1 *= sc1;
#pragma unroll
    for (int d = 0; d < 4; ++d) { st.o[d][0] *= sc0; st.o[d][1] *= sc0; st.o[d][2] *= sc1; st.o[d][3] *= sc1; }
#pragma unroll
    for (int k16 = 0; k16 < NT / 2; ++k16) {
        uint32_t ph[4], pl[4];                        // scores of n-tiles 2 k16, 2 k16 + 1 = the A fragment of this 16-key step
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float p0 = tc::ex2_approx(fmaf(s[2 * k16 + u][0], scale, -ms0)), p1 = tc::ex2_approx(fmaf(s[2 * k16 + u][1], scale, -ms0));
            const float p2 = tc::ex2_approx(fmaf(s[2 * k16 + u][2], scale, -ms1)), p3 = tc::ex2_approx(fmaf(s[2 * k16 + u][3], scale, -ms1));
            st.l0 += p0 + p1; st.l1 += p2 + p3;
            tc::split_bf16x2(p0, p1, ph[2 * u], pl[2 * u]);
            tc::split_bf16x2(p2, p3, ph[2 * u + 1], pl[2 * u + 1]);
        }
        const uint32_t off = sV + (uint32_t)((c0 + 16 * k16) * 64) + la.v;
        uint32_t vh[2][4], vl[2][4];                  // [dp]: dim blocks 2 dp, 2 dp + 1
        ldsm4t(off, vh[0]); ldsm4t(off ^ 32u, vh[1]);
        ldsm4t(off + plane_bytes, vl[0]); ldsm4t((off ^ 32u) + plane_bytes, vl[1]);
#pragma unroll
        for (int d = 0; d < 4; ++d) mma16816(st.o[d], ph, vh[d >> 1][2 * (d & 1)], vh[d >> 1][2 * (d & 1) + 1]);
#pragma unroll
        for (int d = 0; d < 4; ++d) mma16816(st.o[d], pl, vh[d >> 1][2 * (d & 1)], vh[d >> 1][2 * (d & 1) + 1]);
#pragma unroll
        for (int d = 0; d < 4; ++d) mma16816(st.o[d], ph, vl[d >> 1][2 * (d & 1)], vl[d >> 1][2 * (d & 1) + 1]);
    }
}

// One 16-row query tile against the Lkp keys of the head in shared memory (planes Khi, Klo, Vhi, Vlo of plane_bytes each from sK).
// qh / ql: A fragments of the two 16-dim steps ({row g k-low, row g+8 k-low, row g k-high, row g+8 k-high}); `scale` multiplies the
// scores (log2 e / sqrt(32) when the fragments hold the unscaled q, 1 when q was scaled before the split).  flags: per-key mask
// bytes; any32: one "some key masked" byte per 32 keys.  Result: st.o = softmax(S) V, normalised.
template <int NTM>
__device__ __forceinline__ void tile_attention(const uint32_t (&qh)[2][4], const uint32_t (&ql)[2][4], uint32_t sK, uint32_t plane_bytes,
                                               const uint8_t* flags, const uint8_t* any32, int Lkp, bool rf0, bool rf1, float scale,
                                               int lane, TileState& st) {
    const int t = lane & 3;
    const LaneAddr la = lane_addr(lane);
    const uint32_t sV = sK + 2 * plane_bytes;
#pragma unroll
    for (int d = 0; d < 4; ++d) { st.o[d][0] = st.o[d][1] = st.o[d][2] = st.o[d][3] = 0.f; }
    st.m0 = st.m1 = -CUDART_INF_F; st.l0 = st.l1 = 0.f;
    int c0 = 0;
    if (NTM == 8) {
        for (; c0 + 64 <= Lkp; c0 += 64)
            attn_chunk<8>(st, qh, ql, sK, sV, plane_bytes, flags, (any32[c0 >> 5] | any32[(c0 >> 5) + 1]) != 0, c0, rf0, rf1, scale, la, t);
    } else {
        for (; c0 + 64 <= Lkp; c0 += 32)              // 32-key steps (register budget of the 3-CTA variants); the last full one below
            attn_chunk<4>(st, qh, ql, sK, sV, plane_bytes, flags, any32[c0 >> 5] != 0, c0, rf0, rf1, scale, la, t);
    }
    if (Lkp - c0 >= 32) { attn_chunk<4>(st, qh, ql, sK, sV, plane_bytes, flags, any32[c0 >> 5] != 0, c0, rf0, rf1, scale, la, t); c0 += 32; }
    if (Lkp - c0 >= 16) attn_chunk<2>(st, qh, ql, sK, sV, plane_bytes, flags, any32[c0 >> 5] != 0, c0, rf0, rf1, scale, la, t);
    float l0 = st.l0, l1 = st.l1;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;          // every key masked: 0 * inf = NaN like the reference's softmax of -inf rows
#pragma unroll
    for (int d = 0; d < 4; ++d) { st.o[d][0] *= i0; st.o[d][1] *= i0; st.o[d][2] *= i1; st.o[d][3] *= i1; }
}

// rows r0 = 16 tile + g and r0 + 8 of the pair, head h: fp32 and / or planes
__device__ __forceinline__ void store_tile(const MhaRowsArgs& a, long long qbase, int Lq, int r0, int h, int t, const float (&o)[4][4]) {
    const int r1 = r0 + 8;
    if (a.out) {
        if (r0 < Lq) {
            float* p = a.out + (qbase + r0) * a.ldo + h * 32 + 2 * t;
#pragma unroll
            for (int d = 0; d < 4; ++d) *reinterpret_cast<float2*>(p + 8 * d) = make_float2(o[d][0], o[d][1]);
        }
        if (r1 < Lq) {
            float* p = a.out + (qbase + r1) * a.ldo + h * 32 + 2 * t;
#pragma unroll
            for (int d = 0; d < 4; ++d) *reinterpret_cast<float2*>(p + 8 * d) = make_float2(o[d][2], o[d][3]);
        }
    }
    if (a.out_hi) {                                    // pre-split planes for the output projection (row pitch 256 elements)
        uint32_t hh[4][2], ll[4][2];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            tc::split_bf16x2(o[d][0], o[d][1], hh[d][0], ll[d][0]);
            tc::split_bf16x2(o[d][2], o[d][3], hh[d][1], ll[d][1]);
        }
        if (r0 < Lq) {
            const long long e = (qbase + r0) * D + h * 32 + 2 * t;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                *reinterpret_cast<uint32_t*>(a.out_hi + e + 8 * d) = hh[d][0];
                *reinterpret_cast<uint32_t*>(a.out_lo + e + 8 * d) = ll[d][0];
            }
        }
        if (r1 < Lq) {
            const long long e = (qbase + r1) * D + h * 32 + 2 * t;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                *reinterpret_cast<uint32_t*>(a.out_hi + e + 8 * d) = hh[d][1];
                *reinterpret_cast<uint32_t*>(a.out_lo + e + 8 * d) = ll[d][1];
            }
        }
    }
}

// flags of the Lkp keys of one (pair, head) + the per-32-key summaries; called by whole warps (nthreads a multiple of 32)
__device__ __forceinline__ void stage_flags(const MhaRowsArgs& a, uint8_t* flags, uint8_t* any32, int Lk, int Lkp, long long kpad_own, int bp,
                                            bool quirk, int tid, int nthreads);

// key flags of one (pair, head): bit 0 = masked for every query, bit 1 = masked for the quirk's flagged queries.  The reference
// builds attn_mask[b * nheads + h] from pair (b * nheads + h) % B (model/transformer.py T2V layer): bp is that pair.
__device__ __forceinline__ uint8_t key_flag(const MhaRowsArgs& a, int key, int Lk, long long kpad_own, int bp, bool quirk) {
    if (key >= Lk) return 1;
    uint8_t f = (a.k_pad && a.k_pad[kpad_own + key]) ? 1 : 0;
    if (quirk && a.k_pad[(long long)bp * Lk + key]) f |= 2;
    return f;
}

__device__ __forceinline__ void stage_flags(const MhaRowsArgs& a, uint8_t* flags, uint8_t* any32, int Lk, int Lkp, long long kpad_own, int bp,
                                            bool quirk, int tid, int nthreads) {
    for (int key = tid; key < ((Lkp + 31) & ~31); key += nthreads) {
        const uint8_t f = key < Lkp ? key_flag(a, key, Lk, kpad_own, bp, quirk) : 0;
        flags[key] = f;
        const unsigned bal = __ballot_sync(0xffffffffu, f != 0);
        if ((tid & 31) == 0) any32[key >> 5] = bal != 0;
    }
}
__host__ __device__ inline int flag_bytes(int Lkp) { return ((Lkp + 31) & ~31) + 32; }      // flags + any32

// raw query rows g / g + 8 of a 16-row tile, head h (loads only: they stay in flight until split_q consumes them)
struct QRaw { float2 x0[4], x1[4]; };
__device__ __forceinline__ void load_q_raw(const MhaRowsArgs& a, long long qbase, int Lq, int tile, int h, int g, int t, QRaw& r) {
    const int r0 = tile * 16 + g, r1 = r0 + 8;
    const float* q0 = a.q + (qbase + r0) * a.ldq + h * 32 + 2 * t;
    const float* q1 = a.q + (qbase + r1) * a.ldq + h * 32 + 2 * t;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        r.x0[c] = r0 < Lq ? __ldg(reinterpret_cast<const float2*>(q0 + 8 * c)) : make_float2(0.f, 0.f);
        r.x1[c] = r1 < Lq ? __ldg(reinterpret_cast<const float2*>(q1 + 8 * c)) : make_float2(0.f, 0.f);
    }
}
// A fragments of the two 16-dim steps ({row g k-low, row g+8 k-low, row g k-high, row g+8 k-high}), scaled by log2 e / sqrt(32)
__device__ __forceinline__ void split_q(const QRaw& r, float qs, uint32_t (&qh)[2][4], uint32_t (&ql)[2][4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {                      // c = 2 ks + half
        tc::split_bf16x2(r.x0[c].x * qs, r.x0[c].y * qs, qh[c >> 1][2 * (c & 1)], ql[c >> 1][2 * (c & 1)]);
        tc::split_bf16x2(r.x1[c].x * qs, r.x1[c].y * qs, qh[c >> 1][2 * (c & 1) + 1], ql[c >> 1][2 * (c & 1) + 1]);
    }
}
// one float4 of K and of V -> the four operand planes of a head (region = Khi | Klo | Vhi | Vlo, plane_bytes each)
__device__ __forceinline__ void store_kv(uint8_t* region, uint32_t plane_bytes, int key, int c4, const float4& kv, const float4& vv) {
    uint2 kh, kl, vh, vl;
    tc::split_bf16x2(kv.x, kv.y, kh.x, kl.x); tc::split_bf16x2(kv.z, kv.w, kh.y, kl.y);
    tc::split_bf16x2(vv.x, vv.y, vh.x, vl.x); tc::split_bf16x2(vv.z, vv.w, vh.y, vl.y);
    const int off = swz(key, c4 >> 1) + (c4 & 1) * 8;
    *reinterpret_cast<uint2*>(region + off) = kh; *reinterpret_cast<uint2*>(region + plane_bytes + off) = kl;
    *reinterpret_cast<uint2*>(region + 2 * plane_bytes + off) = vh; *reinterpret_cast<uint2*>(region + 3 * plane_bytes + off) = vl;
}

// ---------------------------------------------------------------------------------------------------------------------
// One CTA per (pair, head): the encoder self-attention (up to ~800 keys).  NW warps, MINB CTAs per SM (register budget),
// NTM n-tiles per online-softmax step.
// ---------------------------------------------------------------------------------------------------------------------
template <int NW, int MINB, int NTM>
__global__ void __launch_bounds__(NW * 32, MINB) attn_mma_kernel(const MhaRowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NT = NW * 32;
    const int h = blockIdx.x & (NH - 1), b = blockIdx.x / NH;      // the heads of a pair are neighbours: they read the same rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int bg = a.b0 + b;
    const bool quirk = a.q_pad != nullptr;
    long long kbase, qbase; int Lq, Lk;
    pair_rows(a.q_cu, a.q_enc, b, a.Lq, qbase, Lq);
    pair_rows(a.k_cu, a.k_enc, b, a.Lk, kbase, Lk);
    const int Lkp = (Lk + 15) & ~15;
    const uint32_t plane_bytes = (uint32_t)Lkp * 64;
    uint8_t* flags = smem + 4 * plane_bytes;
    uint8_t* any32 = flags + ((Lkp + 31) & ~31);
    const int bp = quirk ? (int)(((long long)bg * NH + h) % a.Btot) : bg;
    const float qs = a.q_scale * kLog2e;

    int tile = warp;
    QRaw qr;
    if (tile * 16 < Lq) load_q_raw(a, qbase, Lq, tile, h, g, t, qr);      // in flight while the CTA stages K and V
    for (int r = NW * 16 + threadIdx.x; r < Lq; r += NT)                    // rows of the later rounds: into L2 meanwhile
        tc::prefetch_l2(a.q + (qbase + r) * a.ldq + h * 32);
    {
        // K, V -> bf16 hi / lo planes (rows beyond Lk are zero: their P is zero, and 0 x garbage could be NaN).  Loads of four
        // passes are issued before the first conversion: one DRAM round trip per CTA instead of one per pass.
        const int total = Lkp * 8;
        for (int base = threadIdx.x; base < total; base += 4 * NT) {
            float4 kv[4], vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT, key = idx >> 3, c4 = idx & 7;
                kv[u] = make_float4(0.f, 0.f, 0.f, 0.f); vv[u] = kv[u];
                if (idx < total && key < Lk) {
                    kv[u] = __ldg(reinterpret_cast<const float4*>(a.k + (kbase + key) * a.ldk + h * 32 + c4 * 4));
                    vv[u] = __ldg(reinterpret_cast<const float4*>(a.v + (kbase + key) * a.ldv + h * 32 + c4 * 4));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NT;
                if (idx < total) store_kv(smem, plane_bytes, idx >> 3, idx & 7, kv[u], vv[u]);
            }
        }
    }
    stage_flags(a, flags, any32, Lk, Lkp, a.k_cu ? kbase : (long long)bg * Lk, bp, quirk, threadIdx.x, NT);
    __syncthreads();

    const uint32_t sK = tc::smem_u32(smem);
    const int q_pad_ld = a.q_pad_ld ? a.q_pad_ld : a.Lq;
    for (; tile * 16 < Lq; tile += NW) {
        if (tile != warp) load_q_raw(a, qbase, Lq, tile, h, g, t, qr);
        uint32_t qh[2][4], ql[2][4];
        split_q(qr, qs, qh, ql);
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        bool rf0 = false, rf1 = false;
        if (quirk) {
            rf0 = a.q_pad[(long long)bp * q_pad_ld + (r0 < Lq ? r0 : 0)] != 0;
            rf1 = a.q_pad[(long long)bp * q_pad_ld + (r1 < Lq ? r1 : 0)] != 0;
        }
        TileState st;
        tile_attention<NTM>(qh, ql, sK, plane_bytes, flags, any32, Lkp, rf0, rf1, 1.f, lane, st);
        store_tile(a, qbase, Lq, r0, h, t, st.o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One CTA per PAIR, warp w = head w: the T2V cross-attention (<= 64 keys; 17 / 33 words + sentence token in the shipped
// configurations).  K / V of all eight heads are staged once per pair by all threads (whole 1 KB rows, coalesced); every warp then
// walks the pair's query tiles with its head - eight equally long, independent instruction streams - and fetches the next tile's
// query rows (and quirk flags) while it computes the current one.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AH_THREADS = NH * 32;

__global__ void __launch_bounds__(AH_THREADS, 2) attn_mma_heads_kernel(const MhaRowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int b = blockIdx.x;
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int bg = a.b0 + b;
    const bool quirk = a.q_pad != nullptr;
    long long kbase, qbase; int Lq, Lk;
    pair_rows(a.q_cu, a.q_enc, b, a.Lq, qbase, Lq);
    pair_rows(a.k_cu, a.k_enc, b, a.Lk, kbase, Lk);
    const int Lkp = (Lk + 15) & ~15;
    const uint32_t plane_bytes = (uint32_t)Lkp * 64;
    const uint32_t region = (4 * plane_bytes + flag_bytes(Lkp) + 127u) & ~127u;   // per head: Khi Klo Vhi Vlo flags any32 (the ^ 32 of the V reads needs 64-byte aligned planes)
    const int bp = quirk ? (int)(((long long)bg * NH + h) % a.Btot) : bg;
    const int q_pad_ld = a.q_pad_ld ? a.q_pad_ld : a.Lq;
    const float qs = a.q_scale * kLog2e;

    QRaw qr;
    bool nf0 = false, nf1 = false;                                          // quirk flags of the rows in qr
    auto fetch = [&](int tile) {
        load_q_raw(a, qbase, Lq, tile, h, g, t, qr);
        if (quirk) {
            const int r0 = tile * 16 + g, r1 = r0 + 8;
            nf0 = a.q_pad[(long long)bp * q_pad_ld + (r0 < Lq ? r0 : 0)] != 0;
            nf1 = a.q_pad[(long long)bp * q_pad_ld + (r1 < Lq ? r1 : 0)] != 0;
        }
    };
    if (Lq > 0) fetch(0);
    {
        const int total = Lk * 64;                      // float4 index = key * 64 + head * 8 + c4: consecutive threads, consecutive addresses
        for (int base = threadIdx.x; base < total; base += 5 * AH_THREADS) {
            float4 kv[5], vv[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int idx = base + u * AH_THREADS, key = idx >> 6, c = idx & 63;
                if (idx < total) {
                    kv[u] = __ldg(reinterpret_cast<const float4*>(a.k + (kbase + key) * a.ldk + c * 4));
                    vv[u] = __ldg(reinterpret_cast<const float4*>(a.v + (kbase + key) * a.ldv + c * 4));
                }
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int idx = base + u * AH_THREADS, key = idx >> 6, c = idx & 63;
                if (idx < total) store_kv(smem + (c >> 3) * region, plane_bytes, key, c & 7, kv[u], vv[u]);
            }
        }
        // zero rows Lk..Lkp of every head's planes (P is zero there; 0 x garbage could be NaN)
        for (int idx = threadIdx.x; idx < (Lkp - Lk) * 64; idx += AH_THREADS) {
            const int key = Lk + (idx >> 6), c = idx & 63;
            store_kv(smem + (c >> 3) * region, plane_bytes, key, c & 7, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f));
        }
        uint8_t* flags = smem + h * region + 4 * plane_bytes;             // this warp's head
        stage_flags(a, flags, flags + ((Lkp + 31) & ~31), Lk, Lkp, a.k_cu ? kbase : (long long)bg * Lk, bp, quirk, lane, 32);
    }
    __syncthreads();

    const uint32_t sK = tc::smem_u32(smem) + h * region;
    const uint8_t* flags = smem + h * region + 4 * plane_bytes;
    const uint8_t* any32 = flags + ((Lkp + 31) & ~31);
    for (int tile = 0; tile * 16 < Lq; ++tile) {
        uint32_t qh[2][4], ql[2][4];
        split_q(qr, qs, qh, ql);
        const bool rf0 = nf0, rf1 = nf1;
        if ((tile + 1) * 16 < Lq) fetch(tile + 1);
        TileState st;
        tile_attention<4>(qh, ql, sK, plane_bytes, flags, any32, Lkp, rf0, rf1, 1.f, lane, st);
        store_tile(a, qbase, Lq, tile * 16 + g, h, t, st.o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Decoder cross-attention (model/transformer.py:367-381 through model/attention.py): <= 16 queries per pair (10 in every shipped
// configuration) against the pair's Lv clips, per-head operands [content ; sine] (head_dim 64) for the scores and 32 value dims.
// One CTA per pair, warp w = head w.  There is a single query tile, so every key / value element is used exactly once per
// (pair, head): the B fragments are loaded from global memory straight into registers (fp32 -> bf16 hi / lo there) - no shared
// memory, no barrier, every load of a 32-key block independent of the others.  The thread-per-key fp32 kernel it replaces
// (dec_cross_kernel, attention.cu) issued ~1700 instructions per 32 keys and head, this one ~600.
// ---------------------------------------------------------------------------------------------------------------------
// NKS = 16-dim steps of the score product: 4 = [content ; sine] operands (cross-attention), 2 = one 32-dim operand (decoder self-attention)
template <int NKS>
__global__ void __launch_bounds__(NH * 32, 2) dec_cross_mma_kernel(const MhaSmallArgs a) {
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int S = a.S;
    long long kfirst = (long long)b * a.k_bs + a.k_off, kpad0 = (long long)b * a.S;
    if (a.k_cu) {
        const int c0 = a.k_cu[b] - a.k_cu[0];
        S = a.k_cu[b + 1] - a.k_cu[b];
        kfirst = (long long)c0 + (a.k_enc ? b : 0) + a.k_off;
        kpad0 = c0;
    }
    const float qs = a.scale * kLog2e;
    const int r0 = g, r1 = g + 8;
    // A fragments of the four 16-dim steps: steps 0, 1 = content query, steps 2, 3 = sine query
    uint32_t qh[NKS][4], ql[NKS][4];
    {
        const long long row0 = (long long)b * a.q_bs + (long long)(r0 < a.L ? r0 : 0) * a.q_is, row1 = (long long)b * a.q_bs + (long long)(r1 < a.L ? r1 : 0) * a.q_is;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            const float* p0 = (ks < 2 ? a.q + row0 * a.ldq : a.q2 + row0 * a.ldq2) + h * 32 + 16 * (ks & 1) + 2 * t;
            const float* p1 = (ks < 2 ? a.q + row1 * a.ldq : a.q2 + row1 * a.ldq2) + h * 32 + 16 * (ks & 1) + 2 * t;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float2 x0 = r0 < a.L ? __ldg(reinterpret_cast<const float2*>(p0 + 8 * half)) : make_float2(0.f, 0.f);
                const float2 x1 = r1 < a.L ? __ldg(reinterpret_cast<const float2*>(p1 + 8 * half)) : make_float2(0.f, 0.f);
                tc::split_bf16x2(x0.x * qs, x0.y * qs, qh[ks][2 * half], ql[ks][2 * half]);
                tc::split_bf16x2(x1.x * qs, x1.y * qs, qh[ks][2 * half + 1], ql[ks][2 * half + 1]);
            }
        }
    }
    float o[4][4];
#pragma unroll
    for (int d = 0; d < 4; ++d) { o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f; }
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;

    for (int k0 = 0; k0 < S; k0 += 32) {
        // validity of key k0 + lane -> one bit per key of the block
        const int mykey = k0 + lane;
        const bool myvalid = mykey < S && !(a.k_pad && a.k_pad[kpad0 + mykey]);
        const unsigned vmask = __ballot_sync(0xffffffffu, myvalid);
        float s[4][4];
#pragma unroll
        for (int jp = 0; jp < 4; jp += 2) {                     // two n-tiles (8 keys each) at a time: 16 independent 8-byte loads
            float2 raw[2][2 * NKS];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int key = min(k0 + 8 * (jp + u) + g, S - 1);      // keys beyond S: any valid row, their scores are masked below
                const long long krow = kfirst + (long long)key * a.k_is;
                const float* kc = a.k + krow * a.ldk + h * 32 + 2 * t;
#pragma unroll
                for (int c = 0; c < 4; ++c) raw[u][c] = __ldg(reinterpret_cast<const float2*>(kc + 8 * c));
                if (NKS == 4) {
                    const float* kp = a.k2 + (a.k2_table ? (long long)__ldg(a.k2_table + krow) : krow) * a.ldk2 + h * 32 + 2 * t;
#pragma unroll
                    for (int c = 0; c < 4; ++c) raw[u][(NKS == 4 ? 4 : 0) + c] = __ldg(reinterpret_cast<const float2*>(kp + 8 * c));
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint32_t kh[2 * NKS], kl[2 * NKS];              // [2 ks + half]: b0 / b1 of the 16-dim steps
#pragma unroll
                for (int c = 0; c < 2 * NKS; ++c) tc::split_bf16x2(raw[u][c].x, raw[u][c].y, kh[c], kl[c]);
                mma16816z(s[jp + u], qh[0], kh[0], kh[1]);
#pragma unroll
                for (int ks = 1; ks < NKS; ++ks) mma16816(s[jp + u], qh[ks], kh[2 * ks], kh[2 * ks + 1]);
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) mma16816(s[jp + u], ql[ks], kh[2 * ks], kh[2 * ks + 1]);
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) mma16816(s[jp + u], qh[ks], kl[2 * ks], kl[2 * ks + 1]);
            }
        }
        if (vmask != 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned bits = vmask >> (8 * j + 2 * t);
                if (!(bits & 1u)) { s[j][0] = -CUDART_INF_F; s[j][2] = -CUDART_INF_F; }
                if (!(bits & 2u)) { s[j][1] = -CUDART_INF_F; s[j][3] = -CUDART_INF_F; }
            }
        }
        float cm0 = fmaxf(s[0][0], s[0][1]), cm1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
        for (int j = 1; j < 4; ++j) { cm0 = fmaxf(cm0, fmaxf(s[j][0], s[j][1])); cm1 = fmaxf(cm1, fmaxf(s[j][2], s[j][3])); }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
        const float mu0 = mn0 == -CUDART_INF_F ? 0.f : mn0, mu1 = mn1 == -CUDART_INF_F ? 0.f : mn1;
        const float sc0 = tc::ex2_approx(m0 - mu0), sc1 = tc::ex2_approx(m1 - mu1);
        m0 = mn0; m1 = mn1;
        l0 *= sc0; l1 *= sc1;
#pragma unroll
        for (int d = 0; d < 4; ++d) { o[d][0] *= sc0; o[d][1] *= sc0; o[d][2] *= sc1; o[d][3] *= sc1; }
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
            if (k0 + 16 * k16 >= S) break;                      // warp-uniform
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float p0 = tc::ex2_approx(s[2 * k16 + u][0] - mu0), p1 = tc::ex2_approx(s[2 * k16 + u][1] - mu0);
                const float p2 = tc::ex2_approx(s[2 * k16 + u][2] - mu1), p3 = tc::ex2_approx(s[2 * k16 + u][3] - mu1);
                l0 += p0 + p1; l1 += p2 + p3;
                tc::split_bf16x2(p0, p1, ph[2 * u], pl[2 * u]);
                tc::split_bf16x2(p2, p3, ph[2 * u + 1], pl[2 * u + 1]);
            }
            // B fragments of V: {key 2t, key 2t + 1} x dim g (b0) and keys + 8 (b1) per 8-dim n-tile; masked / out-of-range keys have p = 0
            float vr[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int key = min(k0 + 16 * k16 + 2 * t + (kk & 1) + 8 * (kk >> 1), S - 1);
                const float* vp = a.v + (kfirst + (long long)key * a.k_is) * a.ldv + h * 32 + g;
#pragma unroll
                for (int d = 0; d < 4; ++d) vr[kk][d] = __ldg(vp + 8 * d);
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                uint32_t vh0, vl0, vh1, vl1;
                tc::split_bf16x2(vr[0][d], vr[1][d], vh0, vl0);
                tc::split_bf16x2(vr[2][d], vr[3][d], vh1, vl1);
                mma16816(o[d], ph, vh0, vh1);
                mma16816(o[d], pl, vh0, vh1);
                mma16816(o[d], ph, vl0, vl1);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;          // l == 0 -> NaN like the reference
    if (r0 < a.L) {
        float* p = a.out + ((long long)b * a.q_bs + (long long)r0 * a.q_is) * a.ldo + h * 32 + 2 * t;
#pragma unroll
        for (int d = 0; d < 4; ++d) *reinterpret_cast<float2*>(p + 8 * d) = make_float2(o[d][0] * i0, o[d][1] * i0);
    }
    if (r1 < a.L) {
        float* p = a.out + ((long long)b * a.q_bs + (long long)r1 * a.q_is) * a.ldo + h * 32 + 2 * t;
#pragma unroll
        for (int d = 0; d < 4; ++d) *reinterpret_cast<float2*>(p + 8 * d) = make_float2(o[d][2] * i1, o[d][3] * i1);
    }
}

}  // namespace am

static size_t attn_mma_smem(int Lk) {
    const int Lkp = (Lk + 15) & ~15;
    return (size_t)Lkp * 4 * 64 + am::flag_bytes(Lkp) + 128;
}

bool attn_mma_eligible(const MhaRowsArgs& a) {
    if (a.Lk < 1 || a.k_count || a.split_stats) return false;
    if (attn_mma_smem(a.Lk) > 220 * 1024) return false;
    auto al = [](const void* p, int ld, int bytes, int mod) { return ((reinterpret_cast<uintptr_t>(p) & (bytes - 1)) == 0) && (ld % mod == 0); };
    if (a.out && !al(a.out, a.ldo, 8, 2)) return false;
    if (!a.out && !a.out_hi) return false;
    return a.q && a.k && a.v && al(a.q, a.ldq, 8, 2) && al(a.k, a.ldk, 16, 4) && al(a.v, a.ldv, 16, 4);
}

template <int NW, int MINB, int NTM>
static cudaError_t launch_variant(const MhaRowsArgs& a, size_t smem, cudaStream_t s) {
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        MESM_CHECK(cudaFuncSetAttribute(am::attn_mma_kernel<NW, MINB, NTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    am::attn_mma_kernel<NW, MINB, NTM><<<(unsigned)a.B * NH, NW * 32, smem, s>>>(a);
    g_stats.launches++;
    return cudaGetLastError();
}

cudaError_t launch_attn_mma(const MhaRowsArgs& a, cudaStream_t s) {
    const size_t smem = attn_mma_smem(a.Lk);
    ProfScope _ps(a.q_pad ? "attn_mma t2v" : "attn_mma self", s);
    static int heads = -1;             // MESM_ATTN_MMA_HEADS=0: (pair, head) CTAs for every key count
    if (heads < 0) { const char* e = getenv("MESM_ATTN_MMA_HEADS"); heads = (e && e[0] == '0') ? 0 : 1; }
    if (heads && a.Lk <= 64) {         // few keys (T2V): one CTA per pair, one warp per head
        const int Lkp = (a.Lk + 15) & ~15;
        const size_t sm = (size_t)NH * (((size_t)Lkp * 256 + am::flag_bytes(Lkp) + 127) & ~(size_t)127) + 128;
        static size_t attr = 0;
        if (sm > 48 * 1024 && sm > attr) { MESM_CHECK(cudaFuncSetAttribute(am::attn_mma_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); attr = sm; }
        am::attn_mma_heads_kernel<<<a.B, am::AH_THREADS, sm, s>>>(a);
        g_stats.launches++;
        return cudaGetLastError();
    }
    static int variant = -1;           // MESM_ATTN_MMA_VARIANT: 0 = 8 warps x 2 CTAs (128 regs, 64-key steps), 1 = 8 x 3 (85 regs, 32-key steps), 2 = 12 x 2 (85 regs)
    if (variant < 0) { const char* e = getenv("MESM_ATTN_MMA_VARIANT"); variant = e ? atoi(e) : 1; }
    int v = variant;
    if (v == 1 && 3 * (smem + 1024) > 227 * 1024) v = 0;
    if (v == 1) return launch_variant<8, 3, 4>(a, smem, s);
    if (v == 2) return launch_variant<12, 2, 4>(a, smem, s);
    return launch_variant<8, 2, 8>(a, smem, s);
}

bool dec_cross_mma_eligible(const MhaSmallArgs& a) {
    auto al8 = [](const float* p, int ld) { return p && ((reinterpret_cast<uintptr_t>(p) & 7) == 0) && (ld % 2 == 0); };
    if ((a.q2 != nullptr) != (a.k2 != nullptr)) return false;
    if (a.q2 && !(al8(a.q2, a.ldq2) && al8(a.k2, a.ldk2))) return false;
    return a.L >= 1 && a.L <= 16 && a.nheads == NH && a.hq == HD && a.hv == HD && !a.attn_w && !a.causal && a.S >= 1 &&
           al8(a.q, a.ldq) && al8(a.k, a.ldk) && al8(a.out, a.ldo) && a.v;
}

cudaError_t launch_dec_cross_mma(const MhaSmallArgs& a, cudaStream_t s) {
    ProfScope _ps(a.q2 ? "dec_cross_mma" : "dec_self_mma", s);
    if (a.q2) am::dec_cross_mma_kernel<4><<<a.B, NH * 32, 0, s>>>(a);
    else am::dec_cross_mma_kernel<2><<<a.B, NH * 32, 0, s>>>(a);
    g_stats.launches++;
    return cudaGetLastError();
}

}  // namespace mesm
