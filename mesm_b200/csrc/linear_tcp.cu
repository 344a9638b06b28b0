// Persistent CTA-pair variant of the fused linear layer (same LinearOp contract and numerics as linear_tc.cu).
//
// Why a second kernel: per-launch timelines of linear_tc_kernel (tools/tc_probe.cu) show that a 128 x 256 tile with K = 256
// spends ~15 k cycles waiting for its first operands, ~13 k cycles in the K loop and ~12 k cycles in the epilogue, and
// that the two CTAs sharing an SM run those phases in lock-step, so the tensor pipe idles for two thirds of a tile's
// life.  Here ONE CTA per SM stays resident and streams tiles: the operand ring runs ahead across tile boundaries, the
// accumulator is double-buffered in TMEM (2 x 256 columns) and a dedicated group of warps drains tile i while the
// MMAs of tile i+1 are in flight.
//
//   cluster of 2 CTAs (cta_group::2) = one 256 x 256 output tile per step: each CTA converts its own 128 activation rows
//   and fetches HALF of the weight tile (L2 -> SM weight bytes per output row halve).
//     warp 0        : weight producer (bulk async copies of the packed hi/lo planes) + L2 prefetch of the next tile's rows
//     warp 1        : leader CTA: the single MMA-issuing thread (M = 256, N = 256, K = 16, bf16x3); peer CTA: relays
//                     "my half of stage s has landed" to the leader's barrier
//     warps 2..9    : fp32 -> bf16 hi/lo operand converters (3 K blocks of global loads in flight per thread)
//     warps 10..17  : epilogue out of TMEM (LN fold / bias / scale / activation / residual / LayerNorm, coalesced stores)
//   barriers: full_w / full_a / peer / empty per ring stage, tmem_full / tmem_empty per accumulator buffer.
#include "tc_common.cuh"

namespace mesm {
namespace tcp {
using namespace tc;

#ifdef MESM_TC_TIMING
__device__ long long g_tcp_times[256];
#define PSTAMP(i) do { if (blockIdx.x == 0) g_tcp_times[i] = clock64(); } while (0)
#else
#define PSTAMP(i) do {} while (0)
#endif

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int A_TILE = BM * BK * 2;                       // one bf16 plane of this CTA's A block
constexpr int W_TILE = BN * BK * 2;                       // one bf16 plane of the full 256-row weight block (packed image)
constexpr int W_HALF = W_TILE / 2;                        // this CTA's 128 weight rows
constexpr int STG = 2 * A_TILE + 2 * W_HALF;              // 32768
constexpr int NST = 5;
constexpr int PD = 3;                                     // K blocks of activation loads in flight per converter thread
constexpr int NCONV = 256, NEPI = 256;
constexpr int THREADS = 64 + NCONV + NEPI;
constexpr uint32_t IDESC_PAIR = make_idesc(BN, 256);
constexpr int OFF_T = NST * STG;
constexpr int T_BYTES = 8 * 32 * 36 * 4;
constexpr int OFF_BAR = OFF_T + T_BYTES;                  // full_w[8] full_a[8] empty[8] peer[8] tmem_full[2] tmem_empty[2] ptr
constexpr int OFF_LNSTAT = OFF_BAR + 320;                 // [2 buffers][2][128] fused LayerNorm sums of the A rows
constexpr int OFF_LNX = OFF_LNSTAT + 2048;                // [2][128] epilogue LayerNorm exchange
constexpr int OFF_VEC = OFF_LNX + 1024;                   // [4][256] bias, colsum, ln_g, ln_b of the current N tile
constexpr int OFF_ROWOFF = OFF_VEC + 4096;                // [3][128] out, out2, residual row offsets
constexpr int OFF_ROWOFFA = OFF_ROWOFF + 3072;            // [2 tiles][2][128] A / A2 row offsets
constexpr int SMEM_BYTES = OFF_ROWOFFA + 4096 + 1024;

template <int VEC>
__global__ void __launch_bounds__(THREADS, 1) linear_tcp_kernel(const LinearOp op, const int nkb1, const int nkb2, const int npairs,
                                                                const int ntn, const int flags) {
    const int nkbp = op.Apos ? nkb1 : 0;
    const int nkb = nkb1 + nkbp + nkb2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + OFF_BAR;
    const uint32_t bar_full_w = bar_base, bar_full_a = bar_base + 64, bar_empty = bar_base + 128, bar_peer = bar_base + 192,
                   bar_tfull = bar_base + 256, bar_tempty = bar_base + 272;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 288);
    float* ln_stat = reinterpret_cast<float*>(smem + OFF_LNSTAT);
    float* ln_x = reinterpret_cast<float*>(smem + OFF_LNX);
    float* vec_s = reinterpret_cast<float*>(smem + OFF_VEC);
    long long* rowoff = reinterpret_cast<long long*>(smem + OFF_ROWOFF);
    long long* rowoffA = reinterpret_cast<long long*>(smem + OFF_ROWOFFA);

    if (threadIdx.x == 0) PSTAMP(0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = blockIdx.x & 1u;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int nwork = npairs * ntn;
    const int ntiles_mine = cluster_id < nwork ? (nwork - cluster_id + nclusters - 1) / nclusters : 0;
    auto tile_of = [&](int ti, int& m0, int& nt) {          // ti-th work item of this cluster (N tile fastest)
        const int w = cluster_id + ti * nclusters;
        const int p = w / ntn;
        nt = w - p * ntn;
        m0 = (2 * p + (int)cta_rank) * BM;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full_w + 8 * s, 1);
            mbar_init(bar_full_a + 8 * s, NCONV / 32);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_peer + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 2 * (NEPI / 32));   // the epilogue warps of BOTH CTAs release a buffer
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 512; i += THREADS) ln_stat[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) PSTAMP(1);

    if (warp == 0) {
        // ===================== weight producer + L2 prefetch of the NEXT tile's activation / residual rows =====================
        int g = 0;
        for (int ti = 0; ti < ntiles_mine; ++ti) {
            int m0, nt;
            tile_of(ti, m0, nt);
            if ((flags & 1) && ti + 1 < ntiles_mine && op.K <= 1024) {
                int m0n, ntn_;
                tile_of(ti + 1, m0n, ntn_);
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {
                    const int m = m0n + lane + 32 * i;
                    if (m >= op.M) continue;
                    const long long ro = op.amap(m) * (long long)op.lda;
                    for (int c = 0; c < op.K; c += 32) prefetch_l2(op.A + ro + c);
                    if (op.Apos) for (int c = 0; c < op.K; c += 32) prefetch_l2(op.Apos + ro + c);
                    if (op.A2 && op.K2 <= 1024) {
                        const float* r2 = op.A2 + op.a2map(m) * (long long)op.lda2;
                        for (int c = 0; c < op.K2; c += 32) prefetch_l2(r2 + c);
                    }
                    if (op.residual) {
                        const int n0n = ntn_ * BN;
                        const float* r = op.residual + op.rmap(m) * (long long)op.ldr + n0n;
                        const int nn = min(BN, op.N - n0n);
                        for (int c = 0; c < nn; c += 32) prefetch_l2(r + c);
                    }
                }
            }
            if (lane == 0) {
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    const uint32_t ph = (g / NST) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1, 1000 + kb);
                    if (g < 16) PSTAMP(32 + g);
                    const int kk = kb < nkb1 ? kb : (kb < nkb1 + nkbp ? kb - nkb1 : kb - nkb1 - nkbp);
                    const uint8_t* src = kb < nkb1 + nkbp
                                             ? reinterpret_cast<const uint8_t*>(op.Wp) + ((size_t)nt * nkb1 + kk) * (2 * W_TILE)
                                             : reinterpret_cast<const uint8_t*>(op.Wp2) + ((size_t)nt * nkb2 + kk) * (2 * W_TILE);
                    mbar_arrive_expect_tx(bar_full_w + 8 * s, 2 * W_HALF);
                    const uint32_t dst = smem_base + s * STG + 2 * A_TILE;
                    bulk_copy_g2s(dst, src + cta_rank * W_HALF, W_HALF, bar_full_w + 8 * s);
                    bulk_copy_g2s(dst + W_HALF, src + W_TILE + cta_rank * W_HALF, W_HALF, bar_full_w + 8 * s);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // ===================== MMA issuer (leader CTA, for the pair) =====================
            int g = 0;
            for (int ti = 0; ti < ntiles_mine; ++ti) {
                const int b = ti & 1;
                mbar_wait_cluster(bar_tempty + 8 * b, ((ti >> 1) & 1) ^ 1, 6000 + ti);      // both epilogues drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + b * BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    const uint32_t ph = (g / NST) & 1;
                    mbar_wait(bar_full_w + 8 * s, ph, 2000 + kb);
                    mbar_wait(bar_full_a + 8 * s, ph, 3000 + kb);
                    if (g < 16) PSTAMP(64 + g);
                    mbar_wait_cluster(bar_peer + 8 * s, ph, 3500 + kb);
                    if (g < 16) PSTAMP(8 + g);
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + s * STG, a_lo = a_hi + A_TILE;
                    const uint32_t w_hi = a_hi + 2 * A_TILE, w_lo = w_hi + W_HALF;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t koff = k * 32;
                        const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                        const uint64_t dwh = make_desc(w_hi + koff), dwl = make_desc(w_lo + koff);
                        umma2(tacc, dah, dwh, (kb > 0 || k > 0) ? 1u : 0u, IDESC_PAIR);
                        umma2(tacc, dal, dwh, 1u, IDESC_PAIR);
                        umma2(tacc, dah, dwl, 1u, IDESC_PAIR);
                    }
                    umma_commit2(bar_empty + 8 * s);
                }
                umma_commit2(bar_tfull + 8 * b);
                if (ti < 8) PSTAMP(56 + ti);
            }
        } else if (lane == 0) {
            // ===================== peer CTA: relay stage readiness to the leader =====================
            const uint32_t remote = map_to_cta(bar_peer, 0);
            const int total = ntiles_mine * nkb;
            for (int g = 0; g < total; ++g) {
                const int s = g % NST;
                const uint32_t ph = (g / NST) & 1;
                mbar_wait(bar_full_w + 8 * s, ph, 2000);
                mbar_wait(bar_full_a + 8 * s, ph, 3000);
                mbar_arrive_remote(remote + 8 * s);
            }
        }
    } else if (warp < 2 + NCONV / 32) {
        // ===================== activation converters: global fp32 -> bf16 hi/lo swizzled K-major smem =====================
        // Lean by construction (this role is latency-bound per warp): incremental tile / K-block cursors, row offsets held
        // in registers per sweep, packed bf16x2 conversions, constant shared-memory store offsets, one polling lane per warp.
        const int tcid = threadIdx.x - 64;                 // 0..255
        constexpr int PER_ROW = BK / VEC;                  // threads per tile row: 8 (float4) or 16 (float2)
        constexpr int NV = (BM * BK / VEC) / NCONV;        // vector loads per thread per K block: 4 or 8
        constexpr int ROW_STEP = NCONV / PER_ROW;          // 32 or 16 (a multiple of 8: the swizzle term is the same for all i)
        const int cv = tcid % PER_ROW, r0 = tcid / PER_ROW;
        const int total = ntiles_mine * nkb;
        const int sbyte0 = sw64(r0, cv * VEC);

        auto fill_table = [&](int ti) {
            if (tcid < 128 && ti < ntiles_mine) {
                int m0, nt;
                tile_of(ti, m0, nt);
                const int m = m0 + tcid;
                const bool ok = m < op.M;
                long long* t = rowoffA + (ti & 1) * 256;
                t[tcid] = ok ? op.amap(m) * (long long)op.lda : -1;
                t[128 + tcid] = (ok && op.A2) ? op.a2map(m) * (long long)op.lda2 : -1;
            }
        };
        fill_table(0);
        fill_table(1);
        asm volatile("bar.sync 1, 256;" ::: "memory");

        // ---- load cursor (runs PD blocks ahead of the convert cursor) ----
        int l_ti = 0, l_kb = 0, l_kb0 = 0, l_K = op.K;
        const float* l_base = op.A;
        long long l_ro[NV];
        auto l_sweep = [&]() {                             // (re)load the row offsets when a new tile / K sweep starts
            int tab = (l_ti & 1) * 256;
            if (l_kb < nkb1) { l_base = op.A; l_K = op.K; l_kb0 = 0; }
            else if (l_kb < nkb1 + nkbp) { l_base = op.Apos; l_K = op.K; l_kb0 = nkb1; }
            else { l_base = op.A2; l_K = op.K2; l_kb0 = nkb1 + nkbp; tab += 128; }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const long long ro = rowoffA[tab + r0 + i * ROW_STEP];
                l_ro[i] = (ro < 0 ? 0 : ro) + cv * VEC;    // rows beyond M read row 0: their accumulator rows are never stored
            }
        };
        auto load_next = [&](float (&dst)[NV * VEC]) {
            if (l_kb == 0 || l_kb == nkb1 || l_kb == nkb1 + nkbp) l_sweep();
            int kcol = (l_kb - l_kb0) * BK;
            if (kcol + cv * VEC + VEC > l_K) kcol = -cv * VEC;   // K tail: clamped here, re-read element-wise at conversion time
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float* p = l_base + l_ro[i] + kcol;
                if (VEC == 4) {
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(dst[i * 4]), "=f"(dst[i * 4 + 1]), "=f"(dst[i * 4 + 2]), "=f"(dst[i * 4 + 3]) : "l"(p));
                } else {
                    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(dst[i * 2]), "=f"(dst[i * 2 + 1]) : "l"(p));
                }
            }
            if (++l_kb == nkb) { l_kb = 0; ++l_ti; }
        };

        // ---- convert cursor ----
        int c_ti = 0, c_kb = 0, c_s = 0;
        uint32_t c_ph = 0;
        float st_sum[NV], st_sq[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) { st_sum[i] = 0.f; st_sq[i] = 0.f; }

        auto convert_block = [&](int g, float (&src)[NV * VEC]) {
            if (tcid == 0 && g < 16) PSTAMP(128 + g);
            if (c_kb == 0 && c_ti > 0) {
                // every thread has finished tile c_ti-1: its table slot can take tile c_ti+1 (the load cursor gets there next)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                fill_table(c_ti + 1);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            // K sweep of this block and whether it holds the sweep's K tail
            int K, kb0, tab = (c_ti & 1) * 256;
            const float* base;
            if (c_kb < nkb1) { base = op.A; K = op.K; kb0 = 0; }
            else if (c_kb < nkb1 + nkbp) { base = op.Apos; K = op.K; kb0 = nkb1; }
            else { base = op.A2; K = op.K2; kb0 = nkb1 + nkbp; tab += 128; }
            const int kcol = (c_kb - kb0) * BK;
            if (kcol + BK > K) {                               // tail block (at most one per sweep): element-wise, zero padded
                const int k = kcol + cv * VEC;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const long long ro = rowoffA[tab + r0 + i * ROW_STEP];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) src[i * VEC + j] = (k + j < K && ro >= 0) ? base[ro + k + j] : 0.f;
                }
            }
            uint32_t hi[NV * VEC / 2], lo[NV * VEC / 2];
#pragma unroll
            for (int e = 0; e < NV * VEC / 2; ++e) {
                const float x0 = src[2 * e], x1 = src[2 * e + 1];
                const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hu << 16), x1 - __uint_as_float(hu & 0xffff0000u));
                hi[e] = hu;
                lo[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            if (op.fuse_rowstat) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { const float v = src[i * VEC + j]; st_sum[i] += v; st_sq[i] = fmaf(v, v, st_sq[i]); }
            }
            if (tcid == 0 && g < 16) PSTAMP(144 + g);
            if (g + PD < total) load_next(src);                // refill this register buffer: PD blocks stay in flight
            if (tcid == 0 && g < 16) PSTAMP(80 + g);
            if (lane == 0) mbar_wait(bar_empty + 8 * c_s, c_ph ^ 1, 4000 + c_kb);
            __syncwarp();
            if (tcid == 0 && g < 16) PSTAMP(96 + g);
            uint8_t* a_hi = smem + c_s * STG + sbyte0;
            uint8_t* a_lo = a_hi + A_TILE;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (VEC == 4) {
                    *reinterpret_cast<uint2*>(a_hi + i * (ROW_STEP * 64)) = make_uint2(hi[2 * i], hi[2 * i + 1]);
                    *reinterpret_cast<uint2*>(a_lo + i * (ROW_STEP * 64)) = make_uint2(lo[2 * i], lo[2 * i + 1]);
                } else {
                    *reinterpret_cast<uint32_t*>(a_hi + i * (ROW_STEP * 64)) = hi[i];
                    *reinterpret_cast<uint32_t*>(a_lo + i * (ROW_STEP * 64)) = lo[i];
                }
            }
            if (op.fuse_rowstat && c_kb == nkb - 1) {          // the tile's row sums, visible before the last stage is signalled
                float* st = ln_stat + (c_ti & 1) * 256;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    atomicAdd(&st[r0 + i * ROW_STEP], st_sum[i]);
                    atomicAdd(&st[128 + r0 + i * ROW_STEP], st_sq[i]);
                    st_sum[i] = 0.f; st_sq[i] = 0.f;
                }
                __threadfence_block();
            }
            if (tcid == 0 && g < 16) PSTAMP(160 + g);
            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor-core (async) proxy
            if (tcid == 0 && g < 16) PSTAMP(176 + g);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full_a + 8 * c_s);
            if (tcid == 0 && g < 16) PSTAMP(112 + g);
            if (++c_s == NST) { c_s = 0; c_ph ^= 1; }
            if (++c_kb == nkb) { c_kb = 0; ++c_ti; }
        };

        float buf[PD][NV * VEC];
#pragma unroll
        for (int d = 0; d < PD; ++d)
            if (d < total) load_next(buf[d]);
        for (int g = 0; g < total; g += PD) {
#pragma unroll
            for (int d = 0; d < PD; ++d)
                if (g + d < total) convert_block(g + d, buf[d]);
        }
    } else {
        // ===================== epilogue warps: TMEM -> registers -> (transposed through smem) -> global =====================
        const int te = threadIdx.x - 64 - NCONV;           // 0..255
        const int ew = warp - (2 + NCONV / 32);            // 0..7
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                          // column half handled by this warp
        const int row = q * 32 + lane;
        float* T = reinterpret_cast<float*>(smem + OFF_T) + ew * (32 * 36);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int trow0 = q * 32;
        const float slope_eff = op.act == ACT_PRELU ? __ldg(op.prelu) : (op.act == ACT_RELU ? 0.f : 1.f);
        const bool do_ln = op.ln_g != nullptr;
        const uint32_t tempty_leader = map_to_cta(bar_tempty, 0);

        for (int ti = 0; ti < ntiles_mine; ++ti) {
            int m0, nt;
            tile_of(ti, m0, nt);
            const int n0 = nt * BN, b = ti & 1;
            asm volatile("bar.sync 2, 256;" ::: "memory");         // the previous tile's readers of the tables are done
            {
                const int n = n0 + te;
                const bool nok = n < op.N;
                vec_s[te] = (op.bias && nok) ? __ldg(op.bias + n) : 0.f;
                vec_s[256 + te] = (op.colsum && nok) ? __ldg(op.colsum + n) : 0.f;
                vec_s[512 + te] = (op.ln_g && nok) ? __ldg(op.ln_g + n) : 0.f;
                vec_s[768 + te] = (op.ln_b && nok) ? __ldg(op.ln_b + n) : 0.f;
                if (te < 128) {
                    const int m = m0 + te;
                    const bool ok = m < op.M;
                    rowoff[te] = ok ? op.omap(m) * (long long)op.ldo : -1;
                    rowoff[128 + te] = (ok && op.out2) ? op.o2map(m) * (long long)op.ldo2 : -1;
                    rowoff[256 + te] = (ok && op.residual) ? op.rmap(m) * (long long)op.ldr : -1;
                }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const int m = m0 + row;
            const bool mok = m < op.M;
            float mean_in = 0.f, rstd_in = 1.f;
            if (op.rowstat && mok) { mean_in = __ldg(op.rowstat + 2 * m); rstd_in = __ldg(op.rowstat + 2 * m + 1); }

            if (te == 0 && ti < 4) PSTAMP(40 + 4 * ti);
            if (lane == 0) mbar_wait(bar_tfull + 8 * b, (ti >> 1) & 1, 5000 + ti);
            __syncwarp();
            if (te == 0 && ti < 4) PSTAMP(41 + 4 * ti);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + b * BN + half * 128;
            if (op.fuse_rowstat) {
                float* st = ln_stat + b * 256;
                const float invK = 1.f / (float)op.K;
                mean_in = st[row] * invK;
                rstd_in = rsqrtf(fmaxf(st[128 + row] * invK - mean_in * mean_in, 0.f) + 1e-5f);
                asm volatile("bar.sync 2, 256;" ::: "memory");     // both column halves have read the sums
                if (half == 0) { st[row] = 0.f; st[128 + row] = 0.f; }
            }

            // transposed pass over the staged 32x32 chunk starting at column n: each instruction moves 4 rows x 128 bytes
            auto rows_pass = [&](int n, bool add_res, bool keep_in_T, float* dst0, const long long* off0, float* dst1,
                                 const long long* off1, float* dstp) {
                const int nn = n + c4;
                const bool nok = nn < op.N;                 // N % 4 == 0 is an eligibility condition
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(&T[(4 * i + rsub) * 36 + c4]);
                if (add_res) {
                    float4 r[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const long long o = rowoff[2 * 128 + trow0 + 4 * i + rsub];
                        r[i] = (o >= 0 && nok) ? __ldg(reinterpret_cast<const float4*>(op.residual + o + nn)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x[i].x += r[i].x; x[i].y += r[i].y; x[i].z += r[i].z; x[i].w += r[i].w; }
                }
                if (keep_in_T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&T[(4 * i + rsub) * 36 + c4]) = x[i];
                }
                if (nok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int tr = trow0 + 4 * i + rsub;
                        if (dst0) { const long long o = off0[tr]; if (o >= 0) *reinterpret_cast<float4*>(dst0 + o + nn) = x[i]; }
                        if (dst1) { const long long o = off1[tr]; if (o >= 0) *reinterpret_cast<float4*>(dst1 + o + nn) = x[i]; }
                        if (dstp && m0 + tr < op.M) *reinterpret_cast<float4*>(dstp + (long long)(m0 + tr) * op.N + nn) = x[i];
                    }
                }
            };

            float sum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int n = n0 + half * 128 + c * 32;
                if (n >= op.N) break;                           // warp-uniform
                float v[32];
                tmem_ld32(taddr0 + c * 32, v);
                {
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
                    rows_pass(n, op.residual != nullptr, true, nullptr, nullptr, nullptr, nullptr, op.pre_ln);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = *reinterpret_cast<const float4*>(&T[lane * 36 + 4 * j]);
                        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                        sum += (t.x + t.y) + (t.z + t.w);
                    }
                    tmem_st32(taddr0 + c * 32, v);
                } else {
                    rows_pass(n, op.residual != nullptr, false, op.out, rowoff, op.out2, rowoff + 128, nullptr);
                }
                __syncwarp();
            }
            if (do_ln) {                                       // N == 256: this thread holds half of row `row`
                ln_x[half * 128 + q * 32 + lane] = sum;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const float mu = (ln_x[q * 32 + lane] + ln_x[128 + q * 32 + lane]) * (1.f / 256.f);
                asm volatile("bar.sync 2, 256;" ::: "memory");
                float sq = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; sq = fmaf(d, d, sq); }
                }
                ln_x[half * 128 + q * 32 + lane] = sq;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const float rs = rsqrtf((ln_x[q * 32 + lane] + ln_x[128 + q * 32 + lane]) * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int n = n0 + half * 128 + c * 32;
                    float v[32];
                    tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int cl = n - n0 + j;
                        v[j] = (v[j] - mu) * rs * vec_s[512 + cl] + vec_s[768 + cl];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(&T[lane * 36 + 4 * j]) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    rows_pass(n, false, false, op.out, rowoff, op.out2, rowoff + 128, nullptr);
                    __syncwarp();
                }
            }
            // this warp no longer touches accumulator buffer b: release it to the MMA issuer of the pair
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_leader + 8 * b);
            if (te == 0 && ti < 4) PSTAMP(42 + 4 * ti);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace tcp

#ifdef MESM_TC_TIMING
void tcp_read_times(long long* out128) { cudaMemcpyFromSymbol(out128, tcp::g_tcp_times, sizeof(long long) * 256); }
#endif

bool linear_tcp_eligible(const LinearOp& op) {
    if (!linear_tc_eligible(op) || op.M <= tcp::BM) return false;
    const int nkb1 = (op.K + tcp::BK - 1) / tcp::BK, nkb2 = op.A2 ? (op.K2 + tcp::BK - 1) / tcp::BK : 0;
    const int nkb = nkb1 * (op.Apos ? 2 : 1) + nkb2;
    if (nkb < tcp::PD) return false;
    if (op.fuse_rowstat && nkb <= tcp::NST) return false;     // the row sums of tile t+2 must not overtake the epilogue of tile t
    return true;
}

static int g_tcp_flags = 0;       // bit 0: L2 prefetch of the next tile's rows by the producer warp
void tcp_set_flags(int f) { g_tcp_flags = f; }

template <int VEC>
static cudaError_t launch_tcp_variant(const LinearOp& op, int nkb1, int nkb2, cudaStream_t s) {
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        MESM_CHECK(cudaFuncSetAttribute(tcp::linear_tcp_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::SMEM_BYTES));
        int dev = 0;
        MESM_CHECK(cudaGetDevice(&dev));
        MESM_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    const int mt = (op.M + tcp::BM - 1) / tcp::BM, npairs = (mt + 1) / 2, ntn = (op.N + tcp::BN - 1) / tcp::BN;
    const int nclusters = std::min(npairs * ntn, std::max(1, num_sms / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters, 1, 1);
    cfg.blockDim = dim3(tcp::THREADS, 1, 1);
    cfg.dynamicSmemBytes = tcp::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tcp::linear_tcp_kernel<VEC>, op, nkb1, nkb2, npairs, ntn, g_tcp_flags);
}

cudaError_t launch_linear_tcp(const LinearOp& op, cudaStream_t s) {
    const int nkb1 = (op.K + tcp::BK - 1) / tcp::BK, nkb2 = op.A2 ? (op.K2 + tcp::BK - 1) / tcp::BK : 0;
    auto v4 = [](const float* p, int ld, int K) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0) && (K % 4 == 0)); };
    const bool vec4 = v4(op.A, op.lda, op.K) && v4(op.Apos, op.lda, op.K) && v4(op.A2, op.lda2, op.K2);
    const cudaError_t e = vec4 ? launch_tcp_variant<4>(op, nkb1, nkb2, s) : launch_tcp_variant<2>(op, nkb1, nkb2, s);
    g_stats.launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mesm
