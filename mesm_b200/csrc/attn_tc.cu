// Multi-head attention (head_dim 32, up to 224 keys) on the tensor cores: tcgen05 + TMEM, split-bf16 operands.
// Serves the encoder self-attention (keys = Lv+1 clips) and the T2V cross-attention (keys = <= 33 words, with the
// reference's attn_mask quirk) of the MESM path.
//
// One CTA = one (pair, head), one CTA per SM.  K and V^T of the head are converted once to bf16 hi/lo operand tiles in
// shared memory; then TWO 128-row query tiles are in flight at a time (TMEM: 2 x [S 224 cols | O 32 cols] = 512 columns):
//     S = Q K^T                (UMMA M=128, N = valid key blocks x 32, K = 32; fp32 scores in TMEM)
//     softmax                  (one thread per query row straight out of TMEM: row max, exp2, row sum - no shuffles)
//     O = P V                  (UMMA M=128, N=32 per 32-key block; the softmax threads write P as the A operand into a
//                               2-slot ring that the MMA thread drains, so exp and MMA overlap)
//     out = O / rowsum         (transposed through shared memory, coalesced stores)
// bf16x3 everywhere (Qhi Khi + Qlo Khi + Qhi Klo, same for P V) keeps ~1e-5 relative accuracy (see linear_tc.cu).
// Warp 0 = MMA issuer (one thread); warps 1..4 own query tile 0, warps 5..8 query tile 1 of the current tile pair.
//
// Why one CTA per SM: an earlier 1-tile / 2-CTAs-per-SM version of this kernel dead-locked about once per 10^6 CTAs
// inside a full forward (MMA thread waiting for P, softmax warps waiting for a tcgen05.commit arrival that never came)
// and never with a single resident CTA; the barrier watchdog in tc_common.cuh caught it.  Two tiles per CTA recover the
// overlap without co-resident CTAs.
#include "kernels.h"
#include "tc_common.cuh"
#include <math_constants.h>

namespace mesm {
namespace tc {

constexpr int AT_THREADS = 32 + 256;
constexpr int AT_LKP_MAX = 224;
constexpr int AT_OCOL = 224;                      // O accumulator column inside a tile's 256-column TMEM half
// shared-memory map (bytes from the 1024-aligned base)
constexpr int AT_Q = 0;                                                           // [2 tiles][hi 8 KB | lo 8 KB]
constexpr int AT_KHI = 32768, AT_KLO = AT_KHI + AT_LKP_MAX * 64;                  // 14336 each
constexpr int AT_VHI = AT_KLO + AT_LKP_MAX * 64, AT_VLO = AT_VHI + 7 * 2048;      // V^T: 7 key blocks x (32 dims x 64 B)
constexpr int AT_P = AT_VLO + 7 * 2048;                                           // [2 tiles][2 slots][hi 8 KB | lo 8 KB]
constexpr int AT_MASK = AT_P + 4 * 16384;                                         // key masks: own[8], partner[8], nkb_eff
constexpr int AT_BAR = AT_MASK + 128;                                             // 12 mbarriers + tmem pointer
constexpr int AT_SMEM = AT_BAR + 128 + 1024;

__global__ void __launch_bounds__(AT_THREADS, 1) attn_tc_kernel(const MhaRowsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    // barriers (8 bytes each): S[t] @0,8 ; O[t] @16,24 ; Pfull[t][slot] @32+16t+8slot ; Pempty[t][slot] @64+16t+8slot
    const uint32_t bars = sbase + AT_BAR;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + AT_BAR + 96);
    uint32_t* kmask_own = reinterpret_cast<uint32_t*>(smem + AT_MASK);
    uint32_t* kmask_oth = kmask_own + 8;
    int* nkb_eff_s = reinterpret_cast<int*>(kmask_own + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int bg = a.b0 + b;                                            // global pair index (masks are batch-global)
    const bool quirk = a.q_pad != nullptr;
    const int bp = quirk ? (int)(((long long)bg * NH + h) % a.Btot) : bg;   // T2V attn_mask quirk partner (see attention.cu)
    long long kbase, qbase; int Lq, Lk;                                 // rows of this pair (packed layouts: per-pair counts)
    pair_rows(a.q_cu, a.q_enc, b, a.Lq, qbase, Lq);
    pair_rows(a.k_cu, a.k_enc, b, a.Lk, kbase, Lk);
    long long kpad_own = a.k_cu ? kbase : (long long)bg * Lk;
    if (a.k_count > 0) {                                                // key-split pass: this launch sees one chunk of the keys
        const int n = min(a.k_count, Lk - a.k_begin);
        kbase += a.k_begin; kpad_own += a.k_begin;
        Lk = n;
        if (n <= 0) {                                                   // this pair has no key in the chunk: weight 0 in the merge
            for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
                a.split_stats[((qbase + i) * NH + h) * 2] = -CUDART_INF_F;
                a.split_stats[((qbase + i) * NH + h) * 2 + 1] = 0.f;
            }
            return;
        }
    }
    const int q_pad_ld = a.q_pad_ld ? a.q_pad_ld : a.Lq;
    const int Lkp = (Lk + 31) & ~31;
    const int ntiles = (Lq + 127) >> 7;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);                     // S[0..1], O[0..1]
        for (int i = 0; i < 4; ++i) mbar_init(bars + 32 + 8 * i, 4);                // Pfull: 4 softmax warps per tile
        for (int i = 0; i < 4; ++i) mbar_init(bars + 64 + 8 * i, 1);                // Pempty: tcgen05.commit
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // ---- stage K (B operand of S = Q K^T) and V^T (B operand of O = P V) once per (pair, head) ----
    if (warp >= 1) {
        const int t = threadIdx.x - 32;                                    // 0..255
        const int total = Lkp * 8;
        for (int base = t; base < total; base += 256 * 4) {
            float4 kv[4], vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {                                  // issue the batch of loads first
                const int idx = base + u * 256, key = idx >> 3, c4 = idx & 7;
                kv[u] = make_float4(0.f, 0.f, 0.f, 0.f); vv[u] = kv[u];
                if (idx < total && key < Lk) {
                    const long long row = kbase + key;
                    kv[u] = __ldg(reinterpret_cast<const float4*>(a.k + row * a.ldk + h * 32 + c4 * 4));
                    vv[u] = __ldg(reinterpret_cast<const float4*>(a.v + row * a.ldv + h * 32 + c4 * 4));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 256, key = idx >> 3, c4 = idx & 7;
                if (idx >= total) break;
                const float vvv[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
                uint2 kh2, kl2;
                split_bf16x2(kv[u].x, kv[u].y, kh2.x, kl2.x);
                split_bf16x2(kv[u].z, kv[u].w, kh2.y, kl2.y);
                const int kb = sw64(key, c4 * 4);
                *reinterpret_cast<uint2*>(smem + AT_KHI + kb) = kh2;
                *reinterpret_cast<uint2*>(smem + AT_KLO + kb) = kl2;
                const int jb = (key >> 5) * 2048, col = key & 31;
uint32_t vh2[2], vl2[2];
                split_bf16x2(vvv[0], vvv[1], vh2[0], vl2[0]);
                split_bf16x2(vvv[2], vvv[3], vh2[1], vl2[1]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int vb = jb + sw64(c4 * 4 + e, col);
                    *reinterpret_cast<uint16_t*>(smem + AT_VHI + vb) = (uint16_t)(vh2[e >> 1] >> ((e & 1) * 16));
                    *reinterpret_cast<uint16_t*>(smem + AT_VLO + vb) = (uint16_t)(vl2[e >> 1] >> ((e & 1) * 16));
                }
            }
        }
        if (warp == 1) {                                                   // one warp builds the per-block key masks
            int last = 0;
            for (int c = 0; c < (Lkp >> 5); ++c) {
                const int k = c * 32 + lane;
                const int kc = k < Lk ? k : 0;
                const bool m_own = (k >= Lk) || a.k_pad[kpad_own + kc];
                const bool m_oth = quirk && ((k >= Lk) || a.k_pad[(long long)bp * Lk + kc]);
                const unsigned mo = __ballot_sync(0xffffffffu, m_own), mt = __ballot_sync(0xffffffffu, m_oth);
                if (lane == 0) { kmask_own[c] = mo; kmask_oth[c] = mt; }
                if (mo != 0xffffffffu) last = c + 1;
            }
            if (lane == 0) *nkb_eff_s = last > 0 ? last : 1;                // key blocks after the last valid key are skipped
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int nkb_eff = *nkb_eff_s;
    const uint32_t idesc_s = make_idesc(nkb_eff * 32), idesc_o = make_idesc(32);

    const int wt = warp >= 5 ? 1 : 0;                   // which tile of the pair this worker warp serves
    int p_use = 0;                                      // P blocks produced so far per tile (slot = p_use & 1)
    int pair_i = 0;                                     // tile pairs done (parity of the S / O barriers)
    for (int tp = 0; tp < ntiles; tp += 2, ++pair_i) {
        const uint32_t tph = pair_i & 1;
        const int nact = (tp + 1 < ntiles) ? 2 : 1;     // tiles active in this pair
        const int tile = tp + wt;
        const bool gact = warp >= 1 && wt < nact;       // this worker group has a tile
        if (gact) {
            // ---- stage the Q tile (A operand), scaled by head_dim^-0.5 ----
            const int t = (threadIdx.x - 32) & 127;
            uint8_t* qs = smem + AT_Q + wt * 16384;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = t + 128 * i, row = idx >> 3, c4 = idx & 7;
                const int qi = tile * 128 + row;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qi < Lq) q = __ldg(reinterpret_cast<const float4*>(a.q + (qbase + qi) * a.ldq + h * 32 + c4 * 4));
                uint2 qh2, ql2;
                split_bf16x2(q.x * a.q_scale, q.y * a.q_scale, qh2.x, ql2.x);
                split_bf16x2(q.z * a.q_scale, q.w * a.q_scale, qh2.y, ql2.y);
                const int qb = sw64(row, c4 * 4);
                *reinterpret_cast<uint2*>(qs + qb) = qh2;
                *reinterpret_cast<uint2*>(qs + 8192 + qb) = ql2;
            }
        }
        fence_proxy_async();
        __syncthreads();                                                   // [A] operands staged

        if (warp == 0) {
            if (lane == 0) {
                tc_fence_after();
                for (int t2 = 0; t2 < nact; ++t2) {                        // S_t = Q_t K^T, K = 32 -> two K=16 steps
                    const uint32_t qb = sbase + AT_Q + t2 * 16384;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint64_t qh = make_desc(qb + k * 32), ql = make_desc(qb + 8192 + k * 32);
                        const uint64_t kh = make_desc(sbase + AT_KHI + k * 32), kl = make_desc(sbase + AT_KLO + k * 32);
                        umma(tmem_base + t2 * 256, qh, kh, k > 0 ? 1u : 0u, idesc_s);
                        umma(tmem_base + t2 * 256, ql, kh, 1u, idesc_s);
                        umma(tmem_base + t2 * 256, qh, kl, 1u, idesc_s);
                    }
                    umma_commit(bars + 8 * t2);
                }
                int pu = p_use;
                for (int j = 0; j < nkb_eff; ++j, ++pu) {                  // O_t += P_t,j V_j
                    const int slot = pu & 1;
                    for (int t2 = 0; t2 < nact; ++t2) {
                        mbar_wait(bars + 32 + 16 * t2 + 8 * slot, (pu >> 1) & 1, 100 + t2);
                        tc_fence_after();
                        const uint32_t ph_ = sbase + AT_P + (t2 * 2 + slot) * 16384, pl_ = ph_ + 8192;
                        const uint32_t vh_ = sbase + AT_VHI + j * 2048, vl_ = sbase + AT_VLO + j * 2048;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint64_t dph = make_desc(ph_ + k * 32), dpl = make_desc(pl_ + k * 32);
                            const uint64_t dvh = make_desc(vh_ + k * 32), dvl = make_desc(vl_ + k * 32);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dph, dvh, (j > 0 || k > 0) ? 1u : 0u, idesc_o);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dpl, dvh, 1u, idesc_o);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dph, dvl, 1u, idesc_o);
                        }
                        umma_commit(bars + 64 + 16 * t2 + 8 * slot);
                        if (j == nkb_eff - 1) umma_commit(bars + 16 + 8 * t2);
                    }
                }
            }
        } else if (gact) {
            const int q4 = warp & 3;                                       // TMEM lane quadrant of this warp
            const int row = q4 * 32 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + wt * 256;
            const uint32_t bar_Pfull = bars + 32 + 16 * wt, bar_Pempty = bars + 64 + 16 * wt;
            const bool wact = tile * 128 + q4 * 32 < Lq;                   // warp has at least one real query row
            const int qi = tile * 128 + row;
            const bool rflag = quirk && a.q_pad[(long long)bp * q_pad_ld + (qi < Lq ? qi : 0)];
            mbar_wait(bars + 8 * wt, tph, 200 + wt);
            tc_fence_after();
            // pass 1: row maximum over the valid keys
            float mx = -CUDART_INF_F;
            if (wact) {
                for (int c = 0; c < nkb_eff; ++c) {
                    float v[32];
                    tmem_ld32(trow + c * 32, v);
                    const uint32_t mk = kmask_own[c] | (rflag ? kmask_oth[c] : 0u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (!((mk >> j) & 1u)) mx = fmaxf(mx, v[j]);
                }
            }
            const float mxl = mx * 1.4426950408889634f;
            // pass 2: p = exp(s - max) -> bf16 hi/lo A-operand blocks of 32 keys, double-buffered against the P V MMAs
            float sum = 0.f;
            int pu = p_use;
            for (int c = 0; c < nkb_eff; ++c, ++pu) {
                const int slot = pu & 1;
                uint32_t hi[16], lo[16];
                if (wact) {
                    float v[32];
                    tmem_ld32(trow + c * 32, v);
                    const uint32_t mk = kmask_own[c] | (rflag ? kmask_oth[c] : 0u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float p0 = ((mk >> (2 * j)) & 1u) ? 0.f : exp2f(fmaf(v[2 * j], 1.4426950408889634f, -mxl));
                        const float p1 = ((mk >> (2 * j + 1)) & 1u) ? 0.f : exp2f(fmaf(v[2 * j + 1], 1.4426950408889634f, -mxl));
                        sum += p0 + p1;
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
                        const __nv_bfloat162 l2 = __floats2bfloat162_rn(p0 - __uint_as_float(hb << 16), p1 - __uint_as_float(hb & 0xffff0000u));
                        hi[j] = hb;
                        lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { hi[j] = 0u; lo[j] = 0u; }
                }
                mbar_wait(bar_Pempty + 8 * slot, ((pu >> 1) & 1) ^ 1, 300 + wt);    // slot free (first two uses pass at once)
                uint8_t* ph_ = smem + AT_P + (wt * 2 + slot) * 16384;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int off = sw64(row, cc * 8);
                    *reinterpret_cast<uint4*>(ph_ + off) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
                    *reinterpret_cast<uint4*>(ph_ + 8192 + off) = make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_Pfull + 8 * slot);
            }
            // epilogue: O / rowsum, transposed through shared memory (this tile's P ring is idle once its O barrier fires)
            mbar_wait(bars + 16 + 8 * wt, tph, 400 + wt);
            tc_fence_after();
            if (wact) {
                float o[32];
                tmem_ld32(trow + AT_OCOL, o);
                float inv = 1.f / sum;
                if (a.split_stats) {
                    if (sum == 0.f) inv = 0.f;                             // every key of the chunk masked for this row: weight 0 in the merge
                    if (qi < Lq) {
                        a.split_stats[((qbase + qi) * NH + h) * 2] = mx;
                        a.split_stats[((qbase + qi) * NH + h) * 2 + 1] = sum;
                    }
                }
                float* T = reinterpret_cast<float*>(smem + AT_P + wt * 32768) + q4 * (32 * 36);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(&T[lane * 36 + 4 * j]) = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
                __syncwarp();
                const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + rsub;
                    const int qo = tile * 128 + q4 * 32 + r;
                    if (qo < Lq) {
                        const float4 t4 = *reinterpret_cast<const float4*>(&T[r * 36 + c4]);
                        if (a.out) *reinterpret_cast<float4*>(a.out + (qbase + qo) * a.ldo + h * 32 + c4) = t4;
                        if (a.out_hi) {
                            uint2 hh, ll;
                            split_bf16x2(t4.x, t4.y, hh.x, ll.x);
                            split_bf16x2(t4.z, t4.w, hh.y, ll.y);
                            *reinterpret_cast<uint2*>(a.out_hi + (qbase + qo) * D + h * 32 + c4) = hh;
                            *reinterpret_cast<uint2*>(a.out_lo + (qbase + qo) * D + h * 32 + c4) = ll;
                        }
                    }
                }
            }
        }
        p_use += nkb_eff;
        tc_fence_before();
        __syncthreads();                                                   // [B] TMEM / smem reads of this tile pair are done
        tc_fence_after();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}


// =====================================================================================================================
// Pipelined variant: one CTA per PAIR walks its 8 heads; four loader warps stage the next head's operands while the current
// head is in its softmax, and the S MMAs of the next head run under the current head's epilogue.
//
//   loader warps (9..12)   head i+1:  wait "S MMAs of head i retired"  -> K (hi/lo), Q tiles, key masks -> KQ barrier
//                                     wait "P V MMAs of head i retired" -> V^T (hi/lo)                   -> V barrier
//   MMA thread (warp 0)    head i:    wait KQ -> S = Q K^T (both tiles) ; per key block: wait P (and V once) -> O += P V
//                          (S of head i+1 is issued right after the last P block of head i arrived: every softmax thread has
//                           finished reading S by then; its O columns are free once the same warps' epilogue of head i ran,
//                           which precedes their first P of head i+1 in program order)
//   softmax warps (1..8)   head i:    wait S -> max, exp, P blocks -> wait O -> O / rowsum -> global (private transpose scratch)
// TMEM is allocated once per CTA, barriers flip once per head (parity = head & 1).  The one-CTA-per-(pair, head) kernel above
// spent its ~14 us per CTA in strictly serial phases (tensor pipe 7 % active, ncu profiles/r1_final_kernels_ncu.md).
// =====================================================================================================================
// In-kernel timeline of CTA 0 (compile with -DMESM_ATP_TRACE; read back with mesm_debug_trace): slot [head][event] = clock64.
#ifdef MESM_ATP_TRACE
__device__ long long g_atp_trace[8 * 16];
#define ATRACE(h, ev) do { if (blockIdx.x == 0) g_atp_trace[(h) * 16 + (ev)] = clock64(); } while (0)
#else
#define ATRACE(h, ev) do { } while (0)
#endif
constexpr int ATP_THREADS = 32 + 256 + 128;
constexpr int ATP_MASK = AT_P + 4 * 16384;                  // 2 x {own[8], partner[8], nkb_eff, pad} u32 (128 B per buffer)
constexpr int ATP_T = ATP_MASK + 256;                       // epilogue transpose scratch: 8 warps x 32 rows x 36 floats
constexpr int ATP_BAR = ATP_T + 8 * 32 * 36 * 4;            // S[2] 0,8 | O[2] 16,24 | Pfull[2][2] 32.. | Pempty[2][2] 64.. | KQ 96 | V 104 | tmem 112
constexpr int ATP_SMEM = ATP_BAR + 128 + 1024;

__global__ void __launch_bounds__(ATP_THREADS, 1) attn_tcp_kernel(const MhaRowsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + ATP_BAR;
    const uint32_t bar_kq = bars + 96, bar_v = bars + 104;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + ATP_BAR + 112);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + ATP_MASK);       // buffer (head & 1) at masks + 32 * (head & 1)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int bg = a.b0 + b;
    const bool quirk = a.q_pad != nullptr;
    long long kbase, qbase; int Lq, Lk;
    pair_rows(a.q_cu, a.q_enc, b, a.Lq, qbase, Lq);
    pair_rows(a.k_cu, a.k_enc, b, a.Lk, kbase, Lk);
    const long long kpad_own = a.k_cu ? kbase : (long long)bg * Lk;
    const int q_pad_ld = a.q_pad_ld ? a.q_pad_ld : a.Lq;
    const int Lkp = (Lk + 31) & ~31;
    const int nact = Lq > 128 ? 2 : 1;                      // query tiles of this pair (Lq <= 256)

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);                     // S[0..1], O[0..1]: tcgen05.commit
        for (int i = 0; i < 4; ++i) mbar_init(bars + 32 + 8 * i, 4);                // Pfull: 4 softmax warps per tile
        for (int i = 0; i < 4; ++i) mbar_init(bars + 64 + 8 * i, 1);                // Pempty: tcgen05.commit
        mbar_init(bar_kq, 4);                                                       // 4 loader warps
        mbar_init(bar_v, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 9) {
        // ================================ loaders ================================
        const int t = threadIdx.x - 288;                                            // 0..127
        for (int h = 0; h < NH; ++h) {
            const int bp = quirk ? (int)(((long long)bg * NH + h) % a.Btot) : bg;
            if (h > 0) mbar_wait(bars + 8 * (nact - 1), (h - 1) & 1, 600 + h);      // S MMAs of the previous head retired: K / Q free
            if (t == 0) ATRACE(h, 0);
            {   // K -> bf16 hi/lo, K-major SWIZZLE_64B rows of 32
                const int total = Lkp * 8;
                for (int base = t; base < total; base += 128 * 8) {
                    float4 kv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = base + u * 128, key = idx >> 3, c4 = idx & 7;
                        kv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (idx < total && key < Lk) kv[u] = __ldg(reinterpret_cast<const float4*>(a.k + (kbase + key) * a.ldk + h * 32 + c4 * 4));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = base + u * 128, key = idx >> 3, c4 = idx & 7;
                        if (idx >= total) break;
                        uint2 kh2, kl2;
                        split_bf16x2(kv[u].x, kv[u].y, kh2.x, kl2.x);
                        split_bf16x2(kv[u].z, kv[u].w, kh2.y, kl2.y);
                        const int kb = sw64(key, c4 * 4);
                        *reinterpret_cast<uint2*>(smem + AT_KHI + kb) = kh2;
                        *reinterpret_cast<uint2*>(smem + AT_KLO + kb) = kl2;
                    }
                }
            }
            for (int tile = 0; tile < nact; ++tile) {       // Q tiles (A operand), scaled by head_dim^-0.5
                uint8_t* qs = smem + AT_Q + tile * 16384;
                float4 qv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = t + 128 * i, row = idx >> 3, c4 = idx & 7;
                    const int qi = tile * 128 + row;
                    qv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (qi < Lq) qv[i] = __ldg(reinterpret_cast<const float4*>(a.q + (qbase + qi) * a.ldq + h * 32 + c4 * 4));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = t + 128 * i, row = idx >> 3, c4 = idx & 7;
                    uint2 qh2, ql2;
                    split_bf16x2(qv[i].x * a.q_scale, qv[i].y * a.q_scale, qh2.x, ql2.x);
                    split_bf16x2(qv[i].z * a.q_scale, qv[i].w * a.q_scale, qh2.y, ql2.y);
                    const int qb = sw64(row, c4 * 4);
                    *reinterpret_cast<uint2*>(qs + qb) = qh2;
                    *reinterpret_cast<uint2*>(qs + 8192 + qb) = ql2;
                }
            }
            if (warp == 9) {                                                   // per-block key masks of this head
                uint32_t* mk = masks + 32 * (h & 1);
                int last = 0;
                for (int c = 0; c < (Lkp >> 5); ++c) {
                    const int k = c * 32 + lane;
                    const int kc = k < Lk ? k : 0;
                    const bool m_own = (k >= Lk) || a.k_pad[kpad_own + kc];
                    const bool m_oth = quirk && ((k >= Lk) || a.k_pad[(long long)bp * Lk + kc]);
                    const unsigned mo = __ballot_sync(0xffffffffu, m_own), mt = __ballot_sync(0xffffffffu, m_oth);
                    if (lane == 0) { mk[c] = mo; mk[8 + c] = mt; }
                    if (mo != 0xffffffffu) last = c + 1;
                }
                if (lane == 0) mk[16] = (uint32_t)(last > 0 ? last : 1);        // key blocks after the last valid key are skipped
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_kq);
            if (t == 0) ATRACE(h, 1);
            if (h > 0) mbar_wait(bars + 16 + 8 * (nact - 1), (h - 1) & 1, 620 + h);  // P V MMAs of the previous head retired: V free
            if (t == 0) ATRACE(h, 2);
            {   // V^T -> bf16 hi/lo: per 32-key block a [32 dims][32 keys] K-major tile
                const int total = Lkp * 8;
                for (int base = t; base < total; base += 128 * 8) {
                    float4 vv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = base + u * 128, key = idx >> 3, c4 = idx & 7;
                        vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (idx < total && key < Lk) vv[u] = __ldg(reinterpret_cast<const float4*>(a.v + (kbase + key) * a.ldv + h * 32 + c4 * 4));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        // lanes l and l ^ 8 hold the same four dims of keys k (even) and k + 1: swap halves so that the even lane
                        // owns dims 0,1 and the odd lane dims 2,3 of BOTH keys - adjacent keys are adjacent bf16 of a V^T row, so
                        // each lane writes two packed 32-bit words per plane instead of four 16-bit ones (idx < total is
                        // uniform per lane pair: total is a multiple of 256)
                        const int idx = base + u * 128, key = idx >> 3, c4 = idx & 7;
                        const bool odd = key & 1;
                        const float s0 = __shfl_xor_sync(0xffffffffu, odd ? vv[u].x : vv[u].z, 8);
                        const float s1 = __shfl_xor_sync(0xffffffffu, odd ? vv[u].y : vv[u].w, 8);
                        if (idx >= total) continue;
                        const float a0 = odd ? s0 : vv[u].x, b0 = odd ? vv[u].z : s0;      // (key even, key odd) of this lane's first dim
                        const float a1 = odd ? s1 : vv[u].y, b1 = odd ? vv[u].w : s1;      // ... second dim
                        const int d0 = c4 * 4 + (odd ? 2 : 0);
                        const int jb = (key >> 5) * 2048, col = key & 30;
                        uint32_t h0, l0, h1, l1;
                        split_bf16x2(a0, b0, h0, l0);
                        split_bf16x2(a1, b1, h1, l1);
                        const int vb0 = jb + sw64(d0, col), vb1 = jb + sw64(d0 + 1, col);
                        *reinterpret_cast<uint32_t*>(smem + AT_VHI + vb0) = h0;
                        *reinterpret_cast<uint32_t*>(smem + AT_VLO + vb0) = l0;
                        *reinterpret_cast<uint32_t*>(smem + AT_VHI + vb1) = h1;
                        *reinterpret_cast<uint32_t*>(smem + AT_VLO + vb1) = l1;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_v);
            if (t == 0) ATRACE(h, 3);
        }
    } else if (warp == 0) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc_o = make_idesc(32);
            int pu = 0;
            for (int h = 0; h < NH; ++h) {
                mbar_wait(bar_kq, h & 1, 700 + h);
                ATRACE(h, 4);
                tc_fence_after();
                const int nkb_eff = (int)masks[32 * (h & 1) + 16];
                const uint32_t idesc_s = make_idesc(nkb_eff * 32);
                for (int t2 = 0; t2 < nact; ++t2) {                        // S_t = Q_t K^T, K = 32 -> two K=16 steps
                    const uint32_t qb = sbase + AT_Q + t2 * 16384;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint64_t qh = make_desc(qb + k * 32), ql = make_desc(qb + 8192 + k * 32);
                        const uint64_t kh = make_desc(sbase + AT_KHI + k * 32), kl = make_desc(sbase + AT_KLO + k * 32);
                        umma(tmem_base + t2 * 256, qh, kh, k > 0 ? 1u : 0u, idesc_s);
                        umma(tmem_base + t2 * 256, ql, kh, 1u, idesc_s);
                        umma(tmem_base + t2 * 256, qh, kl, 1u, idesc_s);
                    }
                    umma_commit(bars + 8 * t2);
                }
                ATRACE(h, 5);
                for (int j = 0; j < nkb_eff; ++j, ++pu) {                  // O_t += P_t,j V_j
                    const int slot = pu & 1;
                    for (int t2 = 0; t2 < nact; ++t2) {
                        mbar_wait(bars + 32 + 16 * t2 + 8 * slot, (pu >> 1) & 1, 100 + t2);
                        if (j == 0 && t2 == 0) { ATRACE(h, 6); mbar_wait(bar_v, h & 1, 720 + h); ATRACE(h, 7); }
                        tc_fence_after();
                        const uint32_t ph_ = sbase + AT_P + (t2 * 2 + slot) * 16384, pl_ = ph_ + 8192;
                        const uint32_t vh_ = sbase + AT_VHI + j * 2048, vl_ = sbase + AT_VLO + j * 2048;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint64_t dph = make_desc(ph_ + k * 32), dpl = make_desc(pl_ + k * 32);
                            const uint64_t dvh = make_desc(vh_ + k * 32), dvl = make_desc(vl_ + k * 32);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dph, dvh, (j > 0 || k > 0) ? 1u : 0u, idesc_o);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dpl, dvh, 1u, idesc_o);
                            umma(tmem_base + t2 * 256 + AT_OCOL, dph, dvl, 1u, idesc_o);
                        }
                        umma_commit(bars + 64 + 16 * t2 + 8 * slot);
                        if (j == nkb_eff - 1) umma_commit(bars + 16 + 8 * t2);
                    }
                }
                ATRACE(h, 8);
            }
        }
    } else {
        // ================================ softmax / epilogue warps ================================
        const int wt = warp >= 5 ? 1 : 0;
        if (wt < nact) {
            const int q4 = warp & 3;                                       // TMEM lane quadrant of this warp
            const int row = q4 * 32 + lane;
            const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + wt * 256;
            const uint32_t bar_Pfull = bars + 32 + 16 * wt, bar_Pempty = bars + 64 + 16 * wt;
            const bool wact = wt * 128 + q4 * 32 < Lq;                     // warp has at least one real query row
            const int qi = wt * 128 + row;
            float* T = reinterpret_cast<float*>(smem + ATP_T) + (warp - 1) * (32 * 36);
            int pu = 0;
            for (int h = 0; h < NH; ++h) {
                const uint32_t hp = h & 1;
                const int bp = quirk ? (int)(((long long)bg * NH + h) % a.Btot) : bg;
                const bool rflag = quirk && a.q_pad[(long long)bp * q_pad_ld + (qi < Lq ? qi : 0)];
                mbar_wait(bars + 8 * wt, hp, 200 + wt);
                if (threadIdx.x == 32) ATRACE(h, 9);
                tc_fence_after();
                const uint32_t* kmask_own = masks + 32 * hp;
                const uint32_t* kmask_oth = kmask_own + 8;
                const int nkb_eff = (int)kmask_own[16];
                // pass 1: row maximum over the valid keys
                float mx = -CUDART_INF_F;
                if (wact) {
                    for (int c = 0; c < nkb_eff; ++c) {
                        float v[32];
                        tmem_ld32(trow + c * 32, v);
                        const uint32_t mk = kmask_own[c] | (rflag ? kmask_oth[c] : 0u);
                        if (mk == 0u) {                                      // block without masked keys (warp-uniform unless the quirk flag differs)
#pragma unroll
                            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (!((mk >> j) & 1u)) mx = fmaxf(mx, v[j]);
                        }
                    }
                }
                const float mxl = mx * 1.4426950408889634f;
                if (threadIdx.x == 32) ATRACE(h, 10);
                // pass 2: p = exp(s - max) -> bf16 hi/lo A-operand blocks of 32 keys, double-buffered against the P V MMAs
                float sum = 0.f;
                for (int c = 0; c < nkb_eff; ++c, ++pu) {
                    const int slot = pu & 1;
                    uint32_t hi[16], lo[16];
                    if (wact) {
                        float v[32];
                        tmem_ld32(trow + c * 32, v);
                        const uint32_t mk = kmask_own[c] | (rflag ? kmask_oth[c] : 0u);
                        // exponents are <= 0 here: one MUFU.EX2 (ex2.approx.ftz, 2^-22 relative error) instead of exp2f's range fix-ups
                        if (mk == 0u) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float p0 = ex2_approx(fmaf(v[2 * j], 1.4426950408889634f, -mxl));
                                const float p1 = ex2_approx(fmaf(v[2 * j + 1], 1.4426950408889634f, -mxl));
                                sum += p0 + p1;
                                split_bf16x2(p0, p1, hi[j], lo[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float p0 = ((mk >> (2 * j)) & 1u) ? 0.f : ex2_approx(fmaf(v[2 * j], 1.4426950408889634f, -mxl));
                                const float p1 = ((mk >> (2 * j + 1)) & 1u) ? 0.f : ex2_approx(fmaf(v[2 * j + 1], 1.4426950408889634f, -mxl));
                                sum += p0 + p1;
                                split_bf16x2(p0, p1, hi[j], lo[j]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) { hi[j] = 0u; lo[j] = 0u; }
                    }
                    mbar_wait(bar_Pempty + 8 * slot, ((pu >> 1) & 1) ^ 1, 300 + wt);    // slot free (first two uses pass at once)
                    uint8_t* ph_ = smem + AT_P + (wt * 2 + slot) * 16384;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int off = sw64(row, cc * 8);
                        *reinterpret_cast<uint4*>(ph_ + off) = make_uint4(hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
                        *reinterpret_cast<uint4*>(ph_ + 8192 + off) = make_uint4(lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
                    }
                    fence_proxy_async();
                    tc_fence_before();               // this warp's TMEM reads (S, and O of the previous head) precede the arrival
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_Pfull + 8 * slot);
                }
                // epilogue: O / rowsum, transposed through this warp's private scratch
                if (threadIdx.x == 32) ATRACE(h, 11);
                mbar_wait(bars + 16 + 8 * wt, hp, 400 + wt);
                if (threadIdx.x == 32) ATRACE(h, 12);
                tc_fence_after();
                if (wact) {
                    float o[32];
                    tmem_ld32(trow + AT_OCOL, o);
                    const float inv = 1.f / sum;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(&T[lane * 36 + 4 * j]) = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
                    __syncwarp();
                    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + rsub;
                        const int qo = wt * 128 + q4 * 32 + r;
                        if (qo < Lq) {
                            const float4 t4 = *reinterpret_cast<const float4*>(&T[r * 36 + c4]);
                            if (a.out) *reinterpret_cast<float4*>(a.out + (qbase + qo) * a.ldo + h * 32 + c4) = t4;
                            if (a.out_hi) {
                                uint2 hh, ll;
                                split_bf16x2(t4.x, t4.y, hh.x, ll.x);
                                split_bf16x2(t4.z, t4.w, hh.y, ll.y);
                                *reinterpret_cast<uint2*>(a.out_hi + (qbase + qo) * D + h * 32 + c4) = hh;
                                *reinterpret_cast<uint2*>(a.out_lo + (qbase + qo) * D + h * 32 + c4) = ll;
                            }
                        }
                    }
                    __syncwarp();                    // scratch is rewritten by the next head's epilogue
                }
                if (threadIdx.x == 32) ATRACE(h, 13);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace tc

int tc_read_attn_trace(long long* out128) {
#ifdef MESM_ATP_TRACE
    return cudaMemcpyFromSymbol(out128, tc::g_atp_trace, sizeof(long long) * 128) == cudaSuccess ? 128 : 0;
#else
    (void)out128;
    return 0;
#endif
}

void tc_read_watchdog(unsigned long long* out64) { cudaMemcpyFromSymbol(out64, tc::g_tc_watchdog, 512); unsigned long long z[64] = {0}; cudaMemcpyToSymbol(tc::g_tc_watchdog, z, 512); }

// merge of the key-split passes: out[row, head, :] = sum_c w_c O_c / sum_c w_c,  w_c = rowsum_c * exp(rowmax_c - max_c rowmax_c)
__global__ void attn_combine_kernel(const float* __restrict__ parts, const float* __restrict__ stats, int nchunks, long long rows,
                                    long long part_stride, long long stat_stride, float* __restrict__ out, int ldo,
                                    uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;         // (row, head, 4 dims)
    if (idx >= rows * NH * 8) return;
    const int c4 = (int)(idx & 7) * 4, h = (int)((idx >> 3) & 7);
    const long long row = idx >> 6;
    float m = -CUDART_INF_F;
    for (int c = 0; c < nchunks; ++c) m = fmaxf(m, stats[c * stat_stride + (row * NH + h) * 2]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
    for (int c = 0; c < nchunks; ++c) {
        const float mc = stats[c * stat_stride + (row * NH + h) * 2], lc = stats[c * stat_stride + (row * NH + h) * 2 + 1];
        if (lc == 0.f) continue;
        const float w = lc * __expf(mc - m);
        const float4 o = *reinterpret_cast<const float4*>(parts + c * part_stride + row * D + h * 32 + c4);
        acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
        wsum += w;
    }
    const float inv = 1.f / wsum;                                                   // all chunks empty -> NaN like the reference
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (out) *reinterpret_cast<float4*>(out + row * ldo + h * 32 + c4) = acc;
    if (out_hi) {
        uint2 hh, ll;
        tc::split_bf16x2(acc.x, acc.y, hh.x, ll.x);
        tc::split_bf16x2(acc.z, acc.w, hh.y, ll.y);
        *reinterpret_cast<uint2*>(out_hi + row * D + h * 32 + c4) = hh;
        *reinterpret_cast<uint2*>(out_lo + row * D + h * 32 + c4) = ll;
    }
}

size_t attn_split_floats(long long rows, int Lk) {
    if (Lk <= tc::AT_LKP_MAX) return 0;
    const int nch = (Lk + tc::AT_LKP_MAX - 1) / tc::AT_LKP_MAX;
    return (size_t)nch * ((size_t)rows * D + (size_t)rows * NH * 2);
}

// Self-attention over more keys than one tile holds: one attn_tc pass per 224-key chunk, then the exact merge above.
cudaError_t launch_attn_tc_split(const MhaRowsArgs& a, long long rows, cudaStream_t s) {
    const int nch = (a.Lk + tc::AT_LKP_MAX - 1) / tc::AT_LKP_MAX;
    float* parts = a.split_ws;
    float* stats = parts + (size_t)nch * rows * D;
    static bool attr_set = false;
    if (!attr_set) {
        MESM_CHECK(cudaFuncSetAttribute(tc::attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::AT_SMEM));
        attr_set = true;
    }
    {
        ProfScope _ps("attn_tc self (key-split)", s);
        for (int c = 0; c < nch; ++c) {
            MhaRowsArgs p = a;
            p.out = parts + (size_t)c * rows * D; p.ldo = D; p.out_hi = nullptr; p.out_lo = nullptr;
            p.k_begin = c * tc::AT_LKP_MAX; p.k_count = tc::AT_LKP_MAX;
            p.split_stats = stats + (size_t)c * rows * NH * 2;
            dim3 grid(NH, a.B);
            tc::attn_tc_kernel<<<grid, tc::AT_THREADS, tc::AT_SMEM, s>>>(p);
            g_stats.launches++;
        }
    }
    ProfScope _ps2("attn_combine", s);
    const long long n = rows * NH * 8;
    attn_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(parts, stats, nch, rows, rows * D, rows * NH * 2, a.out, a.ldo, a.out_hi, a.out_lo);
    g_stats.launches++;
    return cudaGetLastError();
}

bool attn_tc_split_eligible(const MhaRowsArgs& a) {
    if (a.Lk <= tc::AT_LKP_MAX || a.q_pad || !a.split_ws || a.k_count) return false;       // self-attention only (no quirk partner mask)
    auto al = [](const float* p, int ld) { return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0); };
    return al(a.q, a.ldq) && al(a.k, a.ldk) && al(a.v, a.ldv) && (a.out ? al(a.out, a.ldo) : a.out_hi != nullptr);
}

bool attn_tc_eligible(const MhaRowsArgs& a) {
    if (a.Lk > tc::AT_LKP_MAX || a.Lk < 1) return false;
    // with <= 64 keys (the T2V cross-attention: 17 / 33 words) the per-CTA staging latency outweighs the tensor-core
    // gain; measured 13.9 ms vs 9.5 ms per bench step for the fp32 thread-per-row kernel
    if (a.Lk <= 64) return false;
    auto al = [](const float* p, int ld) { return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0); };
    return al(a.q, a.ldq) && al(a.k, a.ldk) && al(a.v, a.ldv) && (a.out ? al(a.out, a.ldo) : a.out_hi != nullptr);
}

cudaError_t launch_attn_tc(const MhaRowsArgs& a, cudaStream_t s) {
    ProfScope _ps(a.q_pad ? "attn_tc t2v" : "attn_tc self", s);
    static int pipe = -1;
    if (pipe < 0) { const char* e = getenv("MESM_ATTN_PIPE"); pipe = (e && e[0] == '0') ? 0 : 1; }
    if (pipe && a.Lq <= 256) {                    // one CTA per pair, heads pipelined (attn_tcp_kernel)
        static bool attr_set_p = false;
        if (!attr_set_p) {
            MESM_CHECK(cudaFuncSetAttribute(tc::attn_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::ATP_SMEM));
            attr_set_p = true;
        }
        tc::attn_tcp_kernel<<<a.B, tc::ATP_THREADS, tc::ATP_SMEM, s>>>(a);
        g_stats.launches++;
        return cudaGetLastError();
    }
    static bool attr_set = false;
    if (!attr_set) {
        MESM_CHECK(cudaFuncSetAttribute(tc::attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::AT_SMEM));
        attr_set = true;
    }
    dim3 grid(NH, a.B);
    tc::attn_tc_kernel<<<grid, tc::AT_THREADS, tc::AT_SMEM, s>>>(a);
    g_stats.launches++;
    return cudaGetLastError();
}

}  // namespace mesm
