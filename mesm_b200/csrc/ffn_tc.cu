// Fused feed-forward block of the T2V / encoder layers (model/transformer.py:536-539, 647-649) on the tensor cores:
//
//     out = LayerNorm2( R + W2 . PReLU(W1 . X + b1) + b2 )            X, R, out: [M, 256] fp32;  hidden width 1024
//
// One kernel instead of two GEMM launches: the 1024-wide hidden activation never leaves the SM (it used to be written
// to and re-read from HBM/L2 as fp32 - 8 KB per row - and re-split into bf16 hi/lo by the second GEMM).
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 256 rows, 128 per CTA; every weight tile is fetched half by each
// CTA, so the L2 -> SM weight stream is 1 MB per CTA per tile (21 B/clk against 768-cycle MMA groups).
//   shared memory : X as bf16 hi/lo K-major SWIZZLE_64B blocks, resident (128 KB) + a 5-slot ring of 16 KB weight slots
//   tensor memory : Y accumulator [128 x 256] (256 cols) | hidden accumulator chunk [128 x 128] fp32 (128 cols)
//                   | the same chunk after bias + PReLU as packed bf16 hi / lo (64 + 64 cols) = the A operand of GEMM 2,
//                   read by tcgen05.mma straight from TMEM (no shared-memory round trip)
//   per 128-wide hidden chunk j:   G1(j): Hacc  = X . W1[j]^T      (M 256, N 128, K 256, bf16x3)
//                                  E1(j): Hacc -> +b1, PReLU, split hi/lo -> Hbf          (8 warps, registers only)
//                                  G2(j): Y   += Hbf . W2[:, j]^T  (M 256, N 256, K 128, bf16x3, A from TMEM)
//   the MMA thread issues G1(j+1) before G2(j), so the tensor pipe works on the next chunk while E1(j) runs.
//   warp 0 = weight producer, warp 1 = MMA issuer (leader CTA) / ring relay (peer CTA), warps 2..9 = X converters, then
//   E1, then the final epilogue (+b2, +R, LayerNorm, coalesced stores through shared memory).
#include "kernels.h"
#include "tc_common.cuh"
#include "tma_host.h"
#include <cstdlib>
#include <cuda.h>

namespace mesm {
namespace ffn {
using namespace tc;

#ifdef MESM_TC_TIMING
__device__ long long g_ffn_times[160];
#ifndef MESM_TC_TIMING_BLOCK
#define MESM_TC_TIMING_BLOCK 0          // CTA whose timeline is recorded: 0 = first wave (cold HBM burst), e.g. 1480 = a tile of the 10th wave
#endif
#define FSTAMP(i) do { if (blockIdx.x == MESM_TC_TIMING_BLOCK) g_ffn_times[i] = clock64(); } while (0)
#else
#define FSTAMP(i) do {} while (0)
#endif

constexpr int DM = 256, FFD = 1024, BM = 128, HC = 128, NCH = FFD / HC;
constexpr int XBLK = 16384;                      // one K block (32) of X: hi 8 KB | lo 8 KB
constexpr int SLOT = 16384, NSLOT = 5;
constexpr int OFF_RING = 8 * XBLK;               // 131072
constexpr int OFF_BAR = OFF_RING + NSLOT * SLOT; // 212992
constexpr int OFF_B1 = OFF_BAR + 256;            // 1024 floats
constexpr int OFF_VEC = OFF_B1 + 4096;           // b2, ln_g, ln_b: 3 x 256 floats
constexpr int OFF_LNX = OFF_VEC + 3072;          // [2][128]
constexpr int OFF_ROWOFF = OFF_LNX + 1024;       // [2][128] long long: out, residual
constexpr int OFF_LN1 = OFF_ROWOFF + 2048;       // [128][2] (mean, rstd) of the tile's rows (fused LayerNorm-1)
constexpr int SMEM_BYTES = OFF_LN1 + 1024 + 1024;
static_assert(SMEM_BYTES <= 232448, "ffn_pair_kernel: shared memory over the 227 KB limit");
constexpr int THREADS = 320;
constexpr uint32_t IDESC_G1 = make_idesc(HC, 256), IDESC_G2 = make_idesc(DM, 256);
constexpr uint32_t TM_Y = 0, TM_HACC = 256, TM_HHI = 384, TM_HLO = 448;

struct FfnOp {
    const float* X; int ldx;
    const float* R; int ldr;
    float* out; int ldo; RowMap omap;
    int M;
    const uint8_t* W1f; const uint8_t* W2f;
    CUtensorMap tm1, tm2;                    // the packed weight images as [rows of 128 B] tensors: one 16 KB slot = a 128 x 128 B box
    CUtensorMap tmXh, tmXl;                  // X as pre-split bf16 hi / lo planes [M, 256] (box 32 x 128, SWIZZLE_64B); used when x_planes
    int x_planes;                            // 1: X arrives through the TMA engine, no conversion in the kernel
    uint16_t* out_hi; uint16_t* out_lo;      // optional: the result also as planes, same row mapping and pitch as `out` (ldo == 256)
    const float* ln1_g; const float* ln1_b; const float* ln1_stats; int res_ln1;      // LayerNorm-1 applied in the prologue (kernels.h: FfnArgs)
    const float* b1; const float* b2; const float* ln_g; const float* ln_b; const float* prelu;
    int wave_ctas;                           // CTAs resident at once (one per SM): the tile blockIdx.x + wave_ctas starts one tile-time from now
    int dbg;                                 // probe only: bit 0 skips the G1 MMAs, bit 1 the G2 MMAs (timing experiments)
};

__device__ __forceinline__ void umma2_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Tensor-map TMA load of one weight slot.  (Plain cp.async.bulk copies of the same 16 KB took ~5 k cycles each and capped the
// stream at ~14 B/clk per SM with L2 only 14 % busy - see profiles/r1_ffn_fused.md.)
// `bar` is a shared::cluster address: both CTAs of the pair complete their slot on the LEADER's barrier (cta_group::2), so the
// MMA thread waits on one barrier per slot and no relay hop sits in the ring loop.
__device__ __forceinline__ void tma_load_slot(uint32_t dst, const CUtensorMap* tm, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(bar) : "memory");
}
// Remote arrive with the default (CTA-scope) release, as CUTLASS's ClusterBarrier::arrive(cta_id) does.  What the signal orders is data the
// arriving CTA wrote into ITS OWN shared / tensor memory for ITS OWN tensor core (fence.proxy.async / tcgen05.fence before it): nothing
// needs cluster-wide visibility.  The .release.cluster form used in round 1 costs the issuing thread ~1 k cycles per arrive, and the kernel
// does one per K block of X, two per hidden chunk and (before) one per weight slot in every warp (tools/tma_probe.cu).
__device__ __forceinline__ void arrive_remote_light(uint32_t cluster_saddr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_saddr) : "memory");
}
__device__ __forceinline__ void tma_load_x(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(row), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) ffn_pair_kernel(const __grid_constant__ FfnOp op) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + OFF_BAR;
    const uint32_t bar_full = bars, bar_peer = bars + 40, bar_empty = bars + 80, bar_xfull = bars + 128, bar_hacc_full = bars + 192,
                   bar_hacc_free = bars + 200, bar_hbf_full = bars + 208, bar_hbf_free = bars + 216, bar_yfull = bars + 224;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 232);
    float* b1_s = reinterpret_cast<float*>(smem + OFF_B1);
    float* vec_s = reinterpret_cast<float*>(smem + OFF_VEC);
    float* ln_x = reinterpret_cast<float*>(smem + OFF_LNX);
    long long* rowoff = reinterpret_cast<long long*>(smem + OFF_ROWOFF);
    float* ln1_s = reinterpret_cast<float*>(smem + OFF_LN1);

    if (threadIdx.x == 0) FSTAMP(0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = blockIdx.x & 1u;
    const int m0 = blockIdx.x * BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_peer + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }   // full (leader): ONE local arrival + the bytes of both CTAs' loads
        for (int k = 0; k < 8; ++k) mbar_init(bar_xfull + 8 * k, op.x_planes ? 2 : 16);   // the TMA producer / the 8 converter warps of each CTA (used in the leader)
        mbar_init(bar_hacc_full, 1); mbar_init(bar_hbf_free, 1); mbar_init(bar_yfull, 1);
        mbar_init(bar_hacc_free, 16); mbar_init(bar_hbf_full, 16);             // the E1 warps of both CTAs (used in the leader)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) FSTAMP(1);

    // Every pair walks the 8 hidden chunks in a different rotation: the pairs of a wave run in near lock-step and would
    // otherwise all pull the same weight lines from the same L2 slices at the same time.
    const int jrot = (int)((blockIdx.x >> 1) % NCH);
    // order in which the tensor pipe consumes the weight chunks: G1(0), then G1(j+1), G2(j) for j = 0..7
    // segment index e = 0..15 -> (is_g2, chunk)
    auto segment = [](int e, bool& g2, int& j) {
        if (e == 0) { g2 = false; j = 0; }
        else if (e == 2 * NCH - 1) { g2 = true; j = NCH - 1; }
        else { g2 = (e % 2 == 0); j = g2 ? e / 2 - 1 : (e + 1) / 2; }
    };

    if (warp == 0) {
        // ===================== weight producer =====================
        if (lane == 0) {
            int g = 0;
            const uint32_t full_leader = map_to_cta(bar_full, 0);
            if (op.x_planes) {
                // X: 8 K blocks x (hi | lo) = sixteen 8 KB boxes straight into the resident operand image; both CTAs of the pair
                // complete their halves on the LEADER's barrier of the block (rows past M are zero-filled by the TMA engine)
                const uint32_t xfull_leader = map_to_cta(bar_xfull, 0);
                for (int kb = 0; kb < 8; ++kb) {
                    mbar_arrive_expect_tx_remote(xfull_leader + 8 * kb, XBLK);
                    tma_load_x(sbase + kb * XBLK, &op.tmXh, kb * 32, m0, xfull_leader + 8 * kb);
                    tma_load_x(sbase + kb * XBLK + 8192, &op.tmXl, kb * 32, m0, xfull_leader + 8 * kb);
                }
            }
            for (int e = 0; e < 2 * NCH; ++e) {
                bool g2; int j;
                segment(e, g2, j);
                const int jc = (j + jrot) % NCH;               // actual hidden chunk
                for (int t = 0; t < 4; ++t, ++g) {
                    const int s = g % NSLOT;
                    const uint32_t ph = (g / NSLOT) & 1;
                    mbar_wait_spin(bar_empty + 8 * s, ph ^ 1, 1000 + g);
                    if (t == 0) FSTAMP(40 + e);
                    if (g >= 20 && g < 36) FSTAMP(64 + g - 20);
                    if (cta_rank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * SLOT);      // the leader expects both halves; the peer only loads
                    tma_load_slot(sbase + OFF_RING + s * SLOT, g2 ? &op.tm2 : &op.tm1, ((jc * 2 + (int)cta_rank) * 4 + t) * 128, full_leader + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // ===================== MMA issuer (leader CTA) =====================
            int g = 0;
            const uint32_t tY = tmem_base + TM_Y, tH = tmem_base + TM_HACC;
#ifdef MESM_TC_TIMING
            long long w_full = 0, w_peer = 0, w_hacc = 0, w_hbf = 0;
#define WT(acc, stmt) do { const long long t_ = clock64(); stmt; acc += clock64() - t_; } while (0)
#else
#define WT(acc, stmt) do { stmt; } while (0)
#endif
            auto wait_slot = [&](int& s) {
                s = g % NSLOT;
                const uint32_t ph = (g / NSLOT) & 1;
                WT(w_full, mbar_wait_spin(bar_full + 8 * s, ph, 2000 + g, false));      // both CTAs' halves of the slot have landed
                if (g >= 20 && g < 36) FSTAMP(80 + g - 20);
                tc_fence_after();
            };
            auto g1 = [&](int j) {
                for (int t = 0; t < 4; ++t) {
                    int s; wait_slot(s);
                    const uint32_t slot = sbase + OFF_RING + s * SLOT;
#pragma unroll
                    for (int kbi = 0; kbi < 2; ++kbi) {
                        const int kb = t * 2 + kbi;
                        if (j == 0) { mbar_wait_spin(bar_xfull + 8 * kb, 0, 3000 + kb); tc_fence_after(); if (kb == 0) FSTAMP(2); }
                        const uint32_t xa = sbase + kb * XBLK, wb = slot + kbi * 8192;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t koff = k * 32;
                            const uint64_t dah = make_desc(xa + koff), dal = make_desc(xa + 8192 + koff);
                            const uint64_t dwh = make_desc(wb + koff), dwl = make_desc(wb + 4096 + koff);
                            if (op.dbg & 1) continue;
                            umma2(tH, dah, dwh, (t > 0 || kbi > 0 || k > 0) ? 1u : 0u, IDESC_G1);
                            umma2(tH, dal, dwh, 1u, IDESC_G1);
                            umma2(tH, dah, dwl, 1u, IDESC_G1);
                        }
                    }
                    umma_commit2(bar_empty + 8 * s);
                    if (g >= 20 && g < 36) FSTAMP(112 + g - 20);
                    ++g;
                }
                umma_commit2(bar_hacc_full);
                FSTAMP(3 + j);
            };
            auto g2f = [&](int j) {
                for (int t = 0; t < 4; ++t) {
                    int s; wait_slot(s);
                    const uint32_t slot = sbase + OFF_RING + s * SLOT;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t koff = k * 32, kk = t * 2 + k;
                        const uint64_t dwh = make_desc(slot + koff), dwl = make_desc(slot + 8192 + koff);
                        const uint32_t ah = tmem_base + TM_HHI + kk * 8, al = tmem_base + TM_HLO + kk * 8;
                        if (op.dbg & 2) continue;
                        umma2_ts(tY, ah, dwh, (j > 0 || t > 0 || k > 0) ? 1u : 0u, IDESC_G2);
                        umma2_ts(tY, al, dwh, 1u, IDESC_G2);
                        umma2_ts(tY, ah, dwl, 1u, IDESC_G2);
                    }
                    umma_commit2(bar_empty + 8 * s);
                    if (g >= 20 && g < 36) FSTAMP(112 + g - 20);
                    ++g;
                }
                umma_commit2(bar_hbf_free);
                FSTAMP(11 + j);
            };
            g1(0);
            for (int j = 0; j < NCH; ++j) {
                if (j + 1 < NCH) {
                    WT(w_hacc, mbar_wait_spin(bar_hacc_free, j & 1, 4000 + j));       // E1(j) of both CTAs holds Hacc(j) in registers
                    tc_fence_after();
                    g1(j + 1);
                }
                WT(w_hbf, mbar_wait_spin(bar_hbf_full, j & 1, 4500 + j));            // Hbf(j) written in both CTAs
                tc_fence_after();
                g2f(j);
            }
            umma_commit2(bar_yfull);
            FSTAMP(19);
#ifdef MESM_TC_TIMING
            if (blockIdx.x == MESM_TC_TIMING_BLOCK) { g_ffn_times[60] = w_full; g_ffn_times[61] = w_peer; g_ffn_times[62] = w_hacc; g_ffn_times[63] = w_hbf; }
#endif
        }
    } else {
        // ===================== workers: X converters -> E1 per hidden chunk -> final epilogue =====================
        const int tcid = threadIdx.x - 64;                 // 0..255
        const int q = warp & 3, half = (warp - 2) >> 2;    // TMEM lane quadrant / column half of this warp
        const int row = q * 32 + lane;
        {   // per-tile tables and vectors
            for (int i = tcid; i < FFD; i += 256) b1_s[i] = __ldg(op.b1 + i);
            vec_s[tcid] = __ldg(op.b2 + tcid);
            vec_s[256 + tcid] = __ldg(op.ln_g + tcid);
            vec_s[512 + tcid] = __ldg(op.ln_b + tcid);
            if (tcid < 128) {
                const int m = m0 + tcid;
                const bool ok = m < op.M;
                rowoff[tcid] = ok ? op.omap(m) * (long long)op.ldo : -1;
                rowoff[128 + tcid] = ok ? (long long)m * op.ldr : -1;
                if (op.ln1_stats) { ln1_s[2 * tcid] = ok ? __ldg(op.ln1_stats + 2 * (long long)m) : 0.f; ln1_s[2 * tcid + 1] = ok ? __ldg(op.ln1_stats + 2 * (long long)m + 1) : 1.f; }
                if (ok) {                                  // the residual row is read by the final epilogue ~90 k cycles from now: bring it
                    const float* r = op.R + (long long)m * op.ldr;      // into L2 (19 MB chip-wide) so that those loads do not pay HBM latency
#pragma unroll
                    for (int c = 0; c < DM; c += 32) prefetch_l2(r + c);
                }
            }
        }
        // ---- X: fp32 rows -> bf16 hi/lo, K-major SWIZZLE_64B blocks of 32 columns, resident for the whole tile ----
        //      (skipped when the producer of X stored it pre-split: warp 0 then loads the planes with the TMA engine)
        if (!op.x_planes) {
            const int cv = tcid & 7, r0 = tcid >> 3;        // 8 threads per row, 32 rows per pass, 4 passes
            const int sbyte0 = sw64(r0, cv * 4);
            const uint32_t xfull_leader = map_to_cta(bar_xfull, 0);
            const float* xr[4];
            float lmu[4], lrs[4];                            // fused LayerNorm-1: (mean, rstd) of this thread's four rows (identity when unused)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int m = m0 + r0 + 32 * i;
                m = m < op.M ? m : op.M - 1;                 // rows beyond M: any valid row (their outputs are never stored)
                xr[i] = op.X + (long long)m * op.ldx + cv * 4;
                lmu[i] = op.ln1_stats ? __ldg(op.ln1_stats + 2 * (long long)m) : 0.f;
                lrs[i] = op.ln1_stats ? __ldg(op.ln1_stats + 2 * (long long)m + 1) : 1.f;
            }
#pragma unroll
            for (int kb0 = 0; kb0 < 8; kb0 += 4) {
                float4 v[4][4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[kk][i] = __ldg(reinterpret_cast<const float4*>(xr[i] + (kb0 + kk) * 32));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint8_t* xb = smem + (kb0 + kk) * XBLK + sbyte0;
                    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), be4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (op.ln1_stats) {                      // gamma / beta of this thread's 4 columns of the K block
                        g4 = __ldg(reinterpret_cast<const float4*>(op.ln1_g + (kb0 + kk) * 32 + cv * 4));
                        be4 = __ldg(reinterpret_cast<const float4*>(op.ln1_b + (kb0 + kk) * 32 + cv * 4));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 x = v[kk][i];
                        if (op.ln1_stats) {
                            x.x = (x.x - lmu[i]) * lrs[i] * g4.x + be4.x; x.y = (x.y - lmu[i]) * lrs[i] * g4.y + be4.y;
                            x.z = (x.z - lmu[i]) * lrs[i] * g4.z + be4.z; x.w = (x.w - lmu[i]) * lrs[i] * g4.w + be4.w;
                        }
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
                        const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
                        const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - __uint_as_float(u01 << 16), x.y - __uint_as_float(u01 & 0xffff0000u));
                        const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - __uint_as_float(u23 << 16), x.w - __uint_as_float(u23 & 0xffff0000u));
                        *reinterpret_cast<uint2*>(xb + i * 2048) = make_uint2(u01, u23);
                        *reinterpret_cast<uint2*>(xb + 8192 + i * 2048) =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) arrive_remote_light(xfull_leader + 8 * (kb0 + kk));
                }
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");       // tables / vectors visible to all workers
        if (tcid == 0) FSTAMP(20);
        if (!op.x_planes && op.wave_ctas > 0) {
            // L2 prefetch of the X rows of the tile that starts on this SM's successor one wave from now (CTAs are dispatched in index
            // order, one per SM): a wave of tiles starting together asks HBM for 19 MB at once and the first K block reached the MMA
            // thread after ~14 k cycles; prefetched a tile-time ahead, those loads hit L2.  Two threads per row, four 128-byte lines each.
            const long long mn = (long long)(blockIdx.x + op.wave_ctas) * BM + (tcid >> 1);
            if (mn < op.M) {
                const float* xr = op.X + mn * op.ldx + (tcid & 1) * 128;
#pragma unroll
                for (int c = 0; c < 128; c += 32) prefetch_l2(xr + c);
            }
        }

        // ---- E1 per hidden chunk: Hacc -> registers -> +b1, PReLU -> packed bf16 hi / lo -> Hbf (TMEM) ----
        const float slope = __ldg(op.prelu);
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t hacc_free_leader = map_to_cta(bar_hacc_free, 0), hbf_full_leader = map_to_cta(bar_hbf_full, 0);
#pragma unroll 1
        for (int j = 0; j < NCH; ++j) {
            if (lane == 0) mbar_wait(bar_hacc_full, j & 1, 5000 + j);
            __syncwarp();
            if (tcid == 0) FSTAMP(21 + j);
            tc_fence_after();
            float v0[32], v1[32];
            tmem_ld32(lane_base + TM_HACC + half * 64, v0);
            tmem_ld32(lane_base + TM_HACC + half * 64 + 32, v1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_remote_light(hacc_free_leader);       // Hacc may be overwritten by G1(j+1)
            const float* bb = b1_s + ((j + jrot) % NCH) * HC + half * 64;
            float hi[32], lo[32];                                       // packed bf16x2 (bit patterns)
#pragma unroll
            for (int p = 0; p < 32; ++p) {
                const float* src = p < 16 ? v0 : v1;
                const int c = (p & 15) * 2;
                float a0 = src[c] + bb[2 * p], a1 = src[c + 1] + bb[2 * p + 1];
                a0 = a0 >= 0.f ? a0 : slope * a0;
                a1 = a1 >= 0.f ? a1 : slope * a1;
                const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
                const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(a0 - __uint_as_float(hu << 16), a1 - __uint_as_float(hu & 0xffff0000u));
                hi[p] = __uint_as_float(hu);
                lo[p] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&l));
            }
            if (tcid == 0) FSTAMP(56 + j);
            if (lane == 0) mbar_wait(bar_hbf_free, (j & 1) ^ 1, 5500 + j);   // G2(j-1) has consumed the previous Hbf
            __syncwarp();
            if (tcid == 0) FSTAMP(29 + j);
            tc_fence_after();
            tmem_st32(lane_base + TM_HHI + half * 32, hi);
            tmem_st32(lane_base + TM_HLO + half * 32, lo);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_remote_light(hbf_full_leader);
        }

        // ---- final epilogue: Y + b2 + R -> LayerNorm -> out (transposed through shared memory, coalesced) ----
        if (lane == 0) mbar_wait(bar_yfull, 0, 6000);
        __syncwarp();
        if (tcid == 0) FSTAMP(37);
        tc_fence_after();
        const uint32_t taddr0 = lane_base + TM_Y + half * 128;
        float* T = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);      // X blocks are dead once Y is complete
        const int rsub = lane >> 3, c4 = (lane & 7) * 4, trow0 = q * 32;
        auto rows_pass = [&](int n, bool add_res, bool store) {
            const int nn = n + c4;
            float4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(&T[(4 * i + rsub) * 36 + c4]);
            if (add_res) {
                float4 r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long o = rowoff[128 + trow0 + 4 * i + rsub];
                    r[i] = o >= 0 ? __ldg(reinterpret_cast<const float4*>(op.R + o + nn)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (op.res_ln1) {                            // encoder layers: the residual is LayerNorm-1 of the pre-LN rows just read
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(op.ln1_g + nn)), be4 = __ldg(reinterpret_cast<const float4*>(op.ln1_b + nn));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float mu1 = ln1_s[2 * (trow0 + 4 * i + rsub)], rs1 = ln1_s[2 * (trow0 + 4 * i + rsub) + 1];
                        r[i].x = (r[i].x - mu1) * rs1 * g4.x + be4.x; r[i].y = (r[i].y - mu1) * rs1 * g4.y + be4.y;
                        r[i].z = (r[i].z - mu1) * rs1 * g4.z + be4.z; r[i].w = (r[i].w - mu1) * rs1 * g4.w + be4.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { x[i].x += r[i].x; x[i].y += r[i].y; x[i].z += r[i].z; x[i].w += r[i].w; }
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&T[(4 * i + rsub) * 36 + c4]) = x[i];
            }
            if (store) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long o = rowoff[trow0 + 4 * i + rsub];
                    if (o < 0) continue;
                    if (op.out) *reinterpret_cast<float4*>(op.out + o + nn) = x[i];
                    if (op.out_hi) {
                        uint2 hh, ll;
                        split_bf16x2(x[i].x, x[i].y, hh.x, ll.x);
                        split_bf16x2(x[i].z, x[i].w, hh.y, ll.y);
                        *reinterpret_cast<uint2*>(op.out_hi + o + nn) = hh;
                        *reinterpret_cast<uint2*>(op.out_lo + o + nn) = ll;
                    }
                }
            }
        };
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int n = half * 128 + c * 32;
            float v[32];
            tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 b4 = *reinterpret_cast<const float4*>(&vec_s[n + 4 * jj]);
                *reinterpret_cast<float4*>(&T[lane * 36 + 4 * jj]) =
                    make_float4(v[4 * jj] + b4.x, v[4 * jj + 1] + b4.y, v[4 * jj + 2] + b4.z, v[4 * jj + 3] + b4.w);
            }
            __syncwarp();
            rows_pass(n, true, false);
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float4 t = *reinterpret_cast<const float4*>(&T[lane * 36 + 4 * jj]);
                v[4 * jj] = t.x; v[4 * jj + 1] = t.y; v[4 * jj + 2] = t.z; v[4 * jj + 3] = t.w;
                sum += (t.x + t.y) + (t.z + t.w);
            }
            tmem_st32(taddr0 + c * 32, v);
            __syncwarp();
        }
        if (tcid == 0) FSTAMP(150);
        ln_x[half * 128 + row] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mu = (ln_x[row] + ln_x[128 + row]) * (1.f / 256.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) { const float d = v[jj] - mu; sq = fmaf(d, d, sq); }
        }
        if (tcid == 0) FSTAMP(151);
        ln_x[half * 128 + row] = sq;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float rs = rsqrtf((ln_x[row] + ln_x[128 + row]) * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int n = half * 128 + c * 32;
            float v[32];
            tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = (v[jj] - mu) * rs * vec_s[256 + n + jj] + vec_s[512 + n + jj];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<float4*>(&T[lane * 36 + 4 * jj]) = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
            __syncwarp();
            rows_pass(n, false, true);
            __syncwarp();
        }
        if (tcid == 0) FSTAMP(38);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- weight packing: the shared-memory images of the per-CTA weight slots, in consumption order -------------------
//   W1f: [chunk j (8)][cta r (2)][slot t (4)][K block kbi (2)][plane hi|lo][64 rows x 64 B]     W1: [1024, 256]
//   W2f: [chunk j (8)][cta r (2)][slot t (4)][plane hi|lo][128 rows x 64 B]                      W2: [256, 1024]
__global__ void pack_ffn_kernel(const float* __restrict__ W1, const float* __restrict__ W2, uint8_t* __restrict__ W1f,
                                uint8_t* __restrict__ W2f) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // one thread per (matrix, 16-byte chunk of 8 K elements)
    const int per = DM * FFD / 8;
    if (idx >= 2 * per) return;
    const bool second = idx >= per;
    int i = second ? idx - per : idx;
    const int chunk8 = i & 3; i >>= 2;                               // 8-element group inside the 32-wide K block
    uint8_t* dst_hi; uint8_t* dst_lo; const float* src;
    if (!second) {
        const int rl = i & 63; i >>= 6;
        const int kbi = i & 1; i >>= 1;
        const int t = i & 3; i >>= 2;
        const int r = i & 1; i >>= 1;
        const int j = i;
        src = W1 + (long long)(j * HC + r * 64 + rl) * DM + (t * 2 + kbi) * 32 + chunk8 * 8;
        uint8_t* base = W1f + ((size_t)((j * 2 + r) * 4 + t)) * SLOT + kbi * 8192;
        dst_hi = base + sw64(rl, chunk8 * 8); dst_lo = base + 4096 + sw64(rl, chunk8 * 8);
    } else {
        const int rl = i & 127; i >>= 7;
        const int t = i & 3; i >>= 2;
        const int r = i & 1; i >>= 1;
        const int j = i;
        src = W2 + (long long)(r * 128 + rl) * FFD + j * HC + t * 32 + chunk8 * 8;
        uint8_t* base = W2f + ((size_t)((j * 2 + r) * 4 + t)) * SLOT;
        dst_hi = base + sw64(rl, chunk8 * 8); dst_lo = base + 8192 + sw64(rl, chunk8 * 8);
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 h[2], l[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) split_bf16(src[e * 2 + u], h[u], l[u]);
        hi[e] = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        lo[e] = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    }
    *reinterpret_cast<uint4*>(dst_hi) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace ffn

#ifdef MESM_TC_TIMING
void ffn_read_times(long long* out64) { cudaMemcpyFromSymbol(out64, ffn::g_ffn_times, sizeof(long long) * 160); }
#endif

size_t ffn_packed_bytes() { return (size_t)ffn::NCH * 2 * 4 * ffn::SLOT; }      // per matrix (1 MB)

cudaError_t launch_pack_ffn(const float* W1, const float* W2, void* W1f, void* W2f, cudaStream_t s) {
    const int total = 2 * ffn::DM * ffn::FFD / 8;
    ffn::pack_ffn_kernel<<<(total + 255) / 256, 256, 0, s>>>(W1, W2, (uint8_t*)W1f, (uint8_t*)W2f);
    g_stats.launches++;
    return cudaGetLastError();
}

namespace {
bool make_slot_map(CUtensorMap* tm, const void* base, size_t bytes) {
    EncodeTiledFn enc = tma_encode_fn();
    if (!enc) return false;
    const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 128)};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// Host-side tensor maps of one layer's packed FFN weights (returned as an opaque heap object owned by the caller's context).
void* ffn_make_maps(const void* W1f, const void* W2f) {
    CUtensorMap* m = static_cast<CUtensorMap*>(aligned_alloc(64, 2 * sizeof(CUtensorMap)));
    if (!m) return nullptr;
    if (!make_slot_map(&m[0], W1f, ffn_packed_bytes()) || !make_slot_map(&m[1], W2f, ffn_packed_bytes())) { free(m); return nullptr; }
    return m;
}

bool ffn_fused_eligible(const FfnArgs& a) {
    if (!a.maps) return false;
    if (!a.W1f || !a.W2f || a.M <= ffn::BM) return false;
    auto al16 = [](const void* p, long long ld) { return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0); };
    if (a.x_hi ? !a.x_lo : !a.X) return false;
    if (a.ln1_stats && (a.x_hi || !a.ln1_g || !a.ln1_b)) return false;    // fused LayerNorm-1 lives in the fp32 -> bf16 converter
    if (a.res_ln1 && !a.ln1_stats) return false;
    if (a.out_hi && (a.ldo != ffn::DM || !a.out_lo)) return false;        // planes share the row offsets of `out`
    if (!a.out && !a.out_hi) return false;
    return (a.x_hi || al16(a.X, a.ldx)) && al16(a.R, a.ldr) && (!a.out || al16(a.out, a.ldo)) && a.b1 && a.b2 && a.ln_g && a.ln_b && a.prelu;
}

cudaError_t launch_ffn_fused(const FfnArgs& a, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        MESM_CHECK(cudaFuncSetAttribute(ffn::ffn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn::SMEM_BYTES));
        attr_set = true;
    }
    ffn::FfnOp op;
    op.X = a.X; op.ldx = a.ldx; op.R = a.R; op.ldr = a.ldr; op.out = a.out; op.ldo = a.ldo; op.omap = a.omap; op.M = a.M;
    op.W1f = (const uint8_t*)a.W1f; op.W2f = (const uint8_t*)a.W2f;
    op.tm1 = static_cast<const CUtensorMap*>(a.maps)[0]; op.tm2 = static_cast<const CUtensorMap*>(a.maps)[1];
    op.b1 = a.b1; op.b2 = a.b2; op.ln_g = a.ln_g; op.ln_b = a.ln_b; op.prelu = a.prelu;
    op.out_hi = a.out_hi; op.out_lo = a.out_lo;
    op.ln1_g = a.ln1_g; op.ln1_b = a.ln1_b; op.ln1_stats = a.ln1_stats; op.res_ln1 = a.res_ln1;
    op.x_planes = a.x_hi != nullptr;
    if (op.x_planes) {
        if (!tma_map_2d_16bit(&op.tmXh, a.x_hi, ffn::DM, (unsigned long long)a.M, ffn::DM * 2, 32, ffn::BM, CU_TENSOR_MAP_SWIZZLE_64B) ||
            !tma_map_2d_16bit(&op.tmXl, a.x_lo, ffn::DM, (unsigned long long)a.M, ffn::DM * 2, 32, ffn::BM, CU_TENSOR_MAP_SWIZZLE_64B))
            return cudaErrorInvalidValue;
    } else {
        op.tmXh = op.tm1; op.tmXl = op.tm1;
    }
    { static int wave = -1; if (wave < 0) { const char* e = getenv("MESM_FFN_PREFETCH"); int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0); wave = (e && e[0] == '0') ? 0 : (nsm & ~1); } op.wave_ctas = wave; }
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MESM_FFN_DBG"); dbg = e ? atoi(e) : 0; } op.dbg = dbg; }
    const unsigned mt = (unsigned)((a.M + ffn::BM - 1) / ffn::BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mt + 1) & ~1u, 1, 1);
    cfg.blockDim = dim3(ffn::THREADS, 1, 1);
    cfg.dynamicSmemBytes = ffn::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, ffn::ffn_pair_kernel, op);
    g_stats.launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mesm
