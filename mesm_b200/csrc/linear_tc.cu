// Fused linear layer on the 5th-generation tensor cores (tcgen05 + TMEM), split-bf16 ("bf16x3") operands.
//
//   out[M,N] = epilogue( (A (+Apos))[M,K] . W[N,K]^T  (+ A2[M,K2] . W2[N,K2]^T) )       fp32 in, fp32 out
//
// Why bf16x3: the parity bar is <= 1e-3 on the network outputs after ~50 chained GEMMs; the precision study shows
// plain bf16 operands give 3e-3..1e-2 and tf32 1e-3..4e-3, while x = hi + lo with hi = bf16(x), lo = bf16(x - hi) and
// D += Ahi.Whi + Alo.Whi + Ahi.Wlo (fp32 accumulation in TMEM) gives ~1e-5.
//
// One CTA = one 128 x 256 output tile (UMMA M=128, N=256, K=16 per instruction, 256 fp32 TMEM columns).
//   warp 0      : bulk-async-copy (TMA engine, cp.async.bulk) producer of the weight tiles.  Weights are packed once at
//                 load time into the exact shared-memory image the MMA wants (K-major, 128B-swizzled, hi and lo planes),
//                 so a 64-wide K block of the 256-row tile is ONE contiguous 64 KB copy - no tensor map needed.
//   warp 1      : TMEM allocation + the single MMA-issuing thread (3 x 4 tcgen05.mma per K block) + tcgen05.commit.
//   warps 2..9  : stream the fp32 activation tile from global memory (vectorised, coalesced, register double-buffered),
//                 add the positional term, split into bf16 hi/lo and store the swizzled K-major operand tiles; afterwards
//                 the same warps run the epilogue straight out of TMEM: LayerNorm-fold / bias / scale / activation /
//                 residual / LayerNorm over the 256-wide row / stores.
// Two 96 KB stages (A hi/lo 2 x 16 KB + W hi/lo 2 x 32 KB) ride an mbarrier full/empty ring.
#include "tc_common.cuh"

namespace mesm {
namespace tc {

#ifdef MESM_TC_TIMING
__device__ long long g_tc_times[64];
#define TSTAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) g_tc_times[i] = clock64(); } while (0)
#else
#define TSTAMP(i) do {} while (0)
#endif

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int A_TILE = BM * BK * 2;                       // bytes of one bf16 plane of the A tile
constexpr int W_TILE = BN * BK * 2;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * W_TILE;      // 49152: two CTAs (2 x ~105 KB) share one SM
constexpr int NCONV = 256;                                // converter / epilogue threads (8 warps)
constexpr int THREADS = 64 + NCONV;
constexpr uint32_t IDESC = make_idesc(BN);
constexpr int BAR_BYTES = 192;                            // full_w[4] full_a[4] empty[4] (unused[4]) tmem_full tmem_ptr
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + 1024 + 4096 + 3072 + 2048 /*barriers, LN exchange, epilogue vectors, row offsets*/;

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_PRELU) return v >= 0.f ? v : slope * v;
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// VEC = floats per global load of the A operand (4: 16-byte aligned rows; 2: 8-byte aligned rows such as Dv = 2818)
template <int VEC>
__global__ void __launch_bounds__(THREADS, 2) linear_tc_kernel(const LinearOp op, const int nkb1, const int nkb2) {
    // K sweeps accumulated into one tile: A.W^T, then (if present) Apos.W^T with the same weights, then A2.W2^T.
    constexpr int NST = STAGES;
    constexpr int STG = STAGE_BYTES;
    constexpr int WT = W_TILE;                             // bytes of one bf16 plane of the weight tile
    const int nkbp = op.Apos ? nkb1 : 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + NST * STG;
    const uint32_t bar_full_w = bar_base, bar_full_a = bar_base + 32, bar_empty = bar_base + 64, bar_tmem = bar_base + 128;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + NST * STG + 136);
    float* ln_x = reinterpret_cast<float*>(smem + NST * STG + BAR_BYTES);          // [2][128] partial sums
    float* vec_s = ln_x + 256;                                                      // [4][256]: bias, colsum, ln_g, ln_b of this N tile
    long long* rowoff = reinterpret_cast<long long*>(vec_s + 1024);                 // [3][128]: out, out2, residual row offsets (-1: no row)
    long long* rowoffA = rowoff + 3 * 128;                                          // [2][128]: A / A2 operand row offsets

    if (threadIdx.x == 0) TSTAMP(0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // single-CTA tiles: the column tiles of a row tile are adjacent in launch order, so a row tile's operand rows are read from
    // HBM once and from L2 by its other column tiles (N = 512 / 768 / 1024 launches used to stream A once per column tile)
    const int ntn = (op.N + BN - 1) / BN;
    const int mtile = (int)(blockIdx.x / (unsigned)ntn);
    const int nt = (int)(blockIdx.x % (unsigned)ntn);
    const int m0 = mtile * BM, n0 = nt * BN;
    const int nkb = nkb1 + nkbp + nkb2;
    // Every tile walks the K blocks in a different rotation: the CTAs of a wave start together, and without this they
    // all ask L2 for the same weight lines at the same moment (the first weight block took ~10 k cycles to arrive).
    const int krot = (int)((unsigned)mtile % (unsigned)nkb);
    auto rotk = [&](int it) { const int j = it + krot; return j >= nkb ? j - nkb : j; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar_full_w + 8 * s, 1);
            mbar_init(bar_full_a + 8 * s, NCONV / 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) TSTAMP(1);
    // Programmatic dependent launch: the next kernel of the stream may start its own set-up (barriers, TMEM, first weight
    // tiles - none of which depend on this kernel) as soon as every CTA of this grid is resident; it still waits for this
    // grid to complete before it reads activations (griddepcontrol.wait below, on the same path in that kernel).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===================== weight producer + L2 prefetcher of the activation rows =====================
        // The converters keep only two K blocks per thread in flight (registers), which at HBM latency caps the operand
        // stream far below the tensor pipe; the 32 lanes of this warp therefore pull the tile's activation rows (and the
        // residual rows of the epilogue) from HBM into L2 PF_DIST K blocks ahead - prefetches hold no registers.
        constexpr int PF_DIST = 8;
        long long poff[4], poff2[4];                       // row offsets of this lane's 4 tile rows (-1: no such row)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + lane + 32 * i;
            const bool ok = m < op.M;
            poff[i] = ok ? op.amap(m) * (long long)op.lda : -1;
            poff2[i] = (ok && op.A2) ? op.a2map(m) * (long long)op.lda2 : -1;
            if (ok && op.residual && n0 < op.N) {
                const float* r = op.residual + op.rmap(m) * (long long)op.ldr + n0;
                const int nn = min(BN, op.N - n0);
                for (int c = 0; c < nn; c += 32) prefetch_l2(r + c);
            }
        }
        auto prefetch_block = [&](int it) {
            if (it >= nkb) return;
            const int kb = rotk(it);
            const float* base; int K, k0; bool second = false;
            if (kb < nkb1) { base = op.A; K = op.K; k0 = kb * BK; }
            else if (kb < nkb1 + nkbp) { base = op.Apos; K = op.K; k0 = (kb - nkb1) * BK; }
            else { base = op.A2; K = op.K2; k0 = (kb - nkb1 - nkbp) * BK; second = true; }
            const int k1 = min(k0 + BK, K) - 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long ro = second ? poff2[i] : poff[i];
                if (ro < 0) continue;
                const float* r = base + ro;
                prefetch_l2(r + k0);
                if (((reinterpret_cast<uintptr_t>(r + k0)) & 127) != 0) prefetch_l2(r + k1);   // block straddles two lines
            }
        };
        for (int kb = 0; kb < PF_DIST; ++kb) prefetch_block(kb);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NST;
            const uint32_t ph = (kb / NST) & 1;
            prefetch_block(kb + PF_DIST);
            if (lane == 0) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1, 1000 + kb);
                const int jb = rotk(kb);
                const int kk = jb < nkb1 ? jb : (jb < nkb1 + nkbp ? jb - nkb1 : jb - nkb1 - nkbp);
                const uint8_t* src = jb < nkb1 + nkbp
                                         ? reinterpret_cast<const uint8_t*>(op.Wp) + ((size_t)nt * nkb1 + kk) * (2 * W_TILE)
                                         : reinterpret_cast<const uint8_t*>(op.Wp2) + ((size_t)nt * nkb2 + kk) * (2 * W_TILE);
                mbar_arrive_expect_tx(bar_full_w + 8 * s, 2 * WT);
                const uint32_t dst = smem_base + s * STG + 2 * A_TILE;
                bulk_copy_g2s(dst, src, 2 * WT, bar_full_w + 8 * s);      // hi plane, lo plane: contiguous in the packed image
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NST;
                const uint32_t ph = (kb / NST) & 1;
                mbar_wait(bar_full_w + 8 * s, ph, 2000 + kb);
                if (kb < 8) TSTAMP(8 + kb);
                mbar_wait(bar_full_a + 8 * s, ph, 3000 + kb);
                if (kb < 8) TSTAMP(16 + kb);
                tc_fence_after();
                const uint32_t a_hi = smem_base + s * STG, a_lo = a_hi + A_TILE;
                const uint32_t w_hi = a_hi + 2 * A_TILE, w_lo = w_hi + WT;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint32_t koff = k * 32;          // 16 bf16 = 32 bytes along K inside the 64B swizzle row
                    const uint64_t dah = make_desc(a_hi + koff), dal = make_desc(a_lo + koff);
                    const uint64_t dwh = make_desc(w_hi + koff), dwl = make_desc(w_lo + koff);
                    umma(tmem_base, dah, dwh, (kb > 0 || k > 0) ? 1u : 0u, IDESC);
                    umma(tmem_base, dal, dwh, 1u, IDESC);
                    umma(tmem_base, dah, dwl, 1u, IDESC);
                }
                umma_commit(bar_empty + 8 * s);            // stage reusable once these MMAs retire
            }
            umma_commit(bar_tmem);                         // accumulator complete
            TSTAMP(2);
        }
    } else {
        // ===================== A converters, then epilogue =====================
        const int tc = threadIdx.x - 64;                   // 0..255
        constexpr int PER_ROW = BK / VEC;                  // vector loads per tile row (16 or 32)
        constexpr int NV = (BM * BK / VEC) / NCONV;        // vector loads per thread per K block (8 or 16)
        constexpr int ROW_STEP = NCONV / PER_ROW;          // 16 or 8
        const int cv = tc % PER_ROW, r0 = tc / PER_ROW;

        // Row offsets of the tile's 128 rows for the A / A2 operands, once (RowMap needs an integer division).
        if (tc < 128) {
            const int m = m0 + tc;
            const bool ok = m < op.M;
            rowoffA[tc] = ok ? op.amap(m) * (long long)op.lda : -1;
            rowoffA[128 + tc] = (ok && op.A2) ? op.a2map(m) * (long long)op.lda2 : -1;
        }
        {   // stage the per-column epilogue vectors of this N tile once (read by every row of the tile)
            const int n = n0 + tc;
            const bool nok = n < op.N;
            vec_s[tc] = (op.bias && nok) ? __ldg(op.bias + n) : 0.f;
            vec_s[256 + tc] = (op.colsum && nok) ? __ldg(op.colsum + n) : 0.f;
            vec_s[512 + tc] = (op.ln_g && nok) ? __ldg(op.ln_g + n) : 0.f;
            vec_s[768 + tc] = (op.ln_b && nok) ? __ldg(op.ln_b + n) : 0.f;
            if (tc < 128) {
                const int m = m0 + tc;
                const bool ok = m < op.M;
                rowoff[tc] = ok ? op.omap(m) * (long long)op.ldo : -1;
                rowoff[128 + tc] = (ok && op.out2) ? op.o2map(m) * (long long)op.ldo2 : -1;
                rowoff[256 + tc] = (ok && op.residual) ? op.rmap(m) * (long long)op.ldr : -1;
            }
        }
        ln_x[tc] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");   // everything read from here on may come from the preceding kernel

        // Global loads are issued branch-free (clamped addresses, validity applied at conversion time) and TWO K blocks
        // ahead of the block being converted, alternating between two register buffers.
        float buf0[NV * VEC], buf1[NV * VEC];
        float st_sum[NV], st_sq[NV];                    // fused LayerNorm statistics of this thread's A elements
#pragma unroll
        for (int i = 0; i < NV; ++i) { st_sum[i] = 0.f; st_sq[i] = 0.f; }
        struct Src { const float* base; int K; int k0; int tab; };
        auto source = [&](int kb) {
            Src r;
            if (kb < nkb1) { r.base = op.A; r.K = op.K; r.k0 = kb * BK; r.tab = 0; }
            else if (kb < nkb1 + nkbp) { r.base = op.Apos; r.K = op.K; r.k0 = (kb - nkb1) * BK; r.tab = 0; }
            else { r.base = op.A2; r.K = op.K2; r.k0 = (kb - nkb1 - nkbp) * BK; r.tab = 128; }
            return r;
        };
        auto load_block = [&](int kb, float (&dst)[NV * VEC]) {
            const Src sc = source(rotk(kb));
            const int k = sc.k0 + cv * VEC;
            const int kc = (k + VEC <= sc.K) ? k : 0;              // clamped; the K tail is re-read at conversion time
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const long long ro = rowoffA[sc.tab + r0 + i * ROW_STEP];
                const long long off = (ro < 0 ? 0 : ro) + kc;
                if (VEC == 4) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(sc.base + off));
                    dst[i * 4] = t.x; dst[i * 4 + 1] = t.y; dst[i * 4 + 2] = t.z; dst[i * 4 + 3] = t.w;
                } else {
                    const float2 t = __ldg(reinterpret_cast<const float2*>(sc.base + off));
                    dst[i * 2] = t.x; dst[i * 2 + 1] = t.y;
                }
            }
        };
        auto convert_block = [&](int kb, float (&src)[NV * VEC]) {
            const int s = kb % NST;
            const uint32_t ph = (kb / NST) & 1;
            const Src sc = source(rotk(kb));
            const int k = sc.k0 + cv * VEC;
            const int ks = (k + VEC <= sc.K) ? 2 : (k < sc.K ? 1 : 0);
            float cur[NV * VEC];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const long long ro = rowoffA[sc.tab + r0 + i * ROW_STEP];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float v = src[i * VEC + j];
                    if (ks != 2) v = (ks == 1 && k + j < sc.K && ro >= 0) ? sc.base[ro + k + j] : 0.f;   // K tail (last block only)
                    cur[i * VEC + j] = ro >= 0 ? v : 0.f;
                }
            }
            if (op.fuse_rowstat) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { const float v = cur[i * VEC + j]; st_sum[i] += v; st_sq[i] = fmaf(v, v, st_sq[i]); }
            }
            if (kb + 2 < nkb) load_block(kb + 2, src);             // refill this buffer: two blocks stay in flight
            if (tc == 0 && kb < 8) TSTAMP(24 + kb);
            mbar_wait(bar_empty + 8 * s, ph ^ 1, 4000 + kb);
            if (tc == 0 && kb < 8) TSTAMP(32 + kb);
            uint8_t* a_hi = smem + s * STG;
            uint8_t* a_lo = a_hi + A_TILE;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int row = r0 + i * ROW_STEP;
                const int byte = sw64(row, cv * VEC);
                if (VEC == 4) {
                    uint2 ph2, pl2;
                    split_bf16x2(cur[i * VEC], cur[i * VEC + 1], ph2.x, pl2.x);
                    split_bf16x2(cur[i * VEC + 2], cur[i * VEC + 3], ph2.y, pl2.y);
                    *reinterpret_cast<uint2*>(a_hi + byte) = ph2;
                    *reinterpret_cast<uint2*>(a_lo + byte) = pl2;
                } else {
                    uint32_t ph, pl;
                    split_bf16x2(cur[i * VEC], cur[i * VEC + 1], ph, pl);
                    *reinterpret_cast<uint32_t*>(a_hi + byte) = ph;
                    *reinterpret_cast<uint32_t*>(a_lo + byte) = pl;
                }
            }
            fence_proxy_async();                     // make the generic-proxy stores visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full_a + 8 * s);
            if (tc == 0 && kb < 8) TSTAMP(40 + kb);
        };

        load_block(0, buf0);
        if (nkb > 1) load_block(1, buf1);
        for (int kb = 0; kb < nkb; kb += 2) {
            convert_block(kb, buf0);
            if (kb + 1 < nkb) convert_block(kb + 1, buf1);
        }

        // ---------------- epilogue: TMEM -> registers -> global ----------------
        if (op.fuse_rowstat) {                             // per-row sums: PER_ROW threads share a row
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                atomicAdd(&ln_x[r0 + i * ROW_STEP], st_sum[i]);
                atomicAdd(&ln_x[128 + r0 + i * ROW_STEP], st_sq[i]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        mbar_wait(bar_tmem, 0, 5000);
        tc_fence_after();
        if (tc == 0) TSTAMP(3);
        const int q = warp & 3;                            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                  // column half handled by this warp
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool mok = m < op.M;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
        const float slope_eff = op.act == ACT_PRELU ? __ldg(op.prelu) : (op.act == ACT_RELU ? 0.f : 1.f);
        float mean_in = 0.f, rstd_in = 1.f;
        if (op.rowstat && mok) { mean_in = __ldg(op.rowstat + 2 * m); rstd_in = __ldg(op.rowstat + 2 * m + 1); }
        if (op.fuse_rowstat) {
            const float invK = 1.f / (float)op.K;
            mean_in = ln_x[row] * invK;
            rstd_in = rsqrtf(fmaxf(ln_x[128 + row] * invK - mean_in * mean_in, 0.f) + 1e-5f);
            asm volatile("bar.sync 1, 256;" ::: "memory");       // ln_x is reused by the output LayerNorm below
        }
        const float* res = (op.residual && mok) ? op.residual + op.rmap(m) * (long long)op.ldr : nullptr;
        float* out = mok ? op.out + op.omap(m) * (long long)op.ldo : nullptr;
        float* out2 = (op.out2 && mok) ? op.out2 + op.o2map(m) * (long long)op.ldo2 : nullptr;
        float* pre = (op.pre_ln && mok) ? op.pre_ln + (long long)m * op.N : nullptr;
        const bool do_ln = op.ln_g != nullptr;
        const bool vec_ok = ((op.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.out) & 15) == 0) &&
                            (!op.out2 || (((op.ldo2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.out2) & 15) == 0))) &&
                            (!op.residual || (((op.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(op.residual) & 15) == 0)));

        // Per-warp 32 x 32 transpose buffer (the operand stages are idle once the accumulator is complete): TMEM hands each
        // thread one ROW; global memory wants a warp instruction to cover whole 128-byte row segments.  Row stride 36
        // floats keeps both the thread-per-row float4 writes and the 4-rows-per-instruction float4 reads conflict-free.
        float* T = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;
        const int trow0 = q * 32;                           // first tile row of this warp's quadrant
        (void)vec_ok; (void)out; (void)out2; (void)pre; (void)res; (void)row;

        // transposed pass over the staged 32x32 chunk starting at column n: each instruction moves 4 rows x 128 bytes
        auto rows_pass = [&](int n, bool add_res, bool keep_in_T, float* dst0, const long long* off0, float* dst1,
                             const long long* off1, float* dstp) {
            const int nn = n + c4;
            const bool nok = nn < op.N;                     // N % 4 == 0 is an eligibility condition
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
            if (tc == 0 && c < 2) TSTAMP(48 + 4 * c);
            tmem_ld32(taddr0 + c * 32, v);
            if (tc == 0 && c < 2) TSTAMP(49 + 4 * c);
            {   // lean per-element math: LN-fold (identity when unused), bias, scale, leaky activation (slope_eff)
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
            if (tc == 0 && c < 2) TSTAMP(50 + 4 * c);
            if (do_ln) {
                // residual added and pre-LN value stored in the transposed (coalesced) domain; result kept for the stats
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
            if (tc == 0 && c < 2) TSTAMP(51 + 4 * c);
        }
        if (tc == 0) TSTAMP(4);
        if (do_ln) {                                       // N == 256: this thread holds half of row `row`
            ln_x[half * 128 + q * 32 + lane] = sum;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mu = (ln_x[q * 32 + lane] + ln_x[128 + q * 32 + lane]) * (1.f / 256.f);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float sq = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(taddr0 + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; sq = fmaf(d, d, sq); }
            }
            ln_x[half * 128 + q * 32 + lane] = sq;
            asm volatile("bar.sync 1, 256;" ::: "memory");
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
    }
    if (threadIdx.x == 64) TSTAMP(5);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TSTAMP(6);
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN));
    }
}

// ---- weight packing: fp32 W[N,K] rows [row0,row0+nrows) (x gamma[k]) -> bf16 hi/lo tiles in the swizzled smem image ----
__global__ void pack_tc_kernel(const float* __restrict__ W, int row0, int nrows, int K, const float* __restrict__ gamma,
                               uint8_t* __restrict__ out, int ntiles, int nkb) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (tile row, 8-element chunk)
    const long long total = (long long)ntiles * nkb * BN * 4;
    if (idx >= total) return;
    const int chunk = (int)(idx & 3);
    const int nl = (int)((idx >> 2) % BN);
    const long long tkb = (idx >> 2) / BN;
    const int kb = (int)(tkb % nkb), t = (int)(tkb / nkb);
    const int n = t * BN + nl;
    uint8_t* tile = out + ((size_t)t * nkb + kb) * (2 * W_TILE);
    const int byte = sw64(nl, chunk * 8);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 h[2], l[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = kb * BK + chunk * 8 + e * 2 + u;
            float w = 0.f;
            if (n < nrows && k < K) w = W[(long long)(row0 + n) * K + k] * (gamma ? gamma[k] : 1.f);
            split_bf16(w, h[u], l[u]);
        }
        hi[e] = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        lo[e] = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    }
    *reinterpret_cast<uint4*>(tile + byte) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile + W_TILE + byte) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace tc

size_t tc_packed_bytes(int nrows, int K) {
    const int ntiles = (nrows + tc::BN - 1) / tc::BN, nkb = (K + tc::BK - 1) / tc::BK;
    return (size_t)ntiles * nkb * 2 * tc::W_TILE;
}

cudaError_t launch_pack_tc(const float* W, int row0, int nrows, int K, const float* gamma, void* out, cudaStream_t s) {
    const int ntiles = (nrows + tc::BN - 1) / tc::BN, nkb = (K + tc::BK - 1) / tc::BK;
    const long long total = (long long)ntiles * nkb * tc::BN * 4;
    tc::pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, row0, nrows, K, gamma, (uint8_t*)out, ntiles, nkb);
    g_stats.launches++;
    return cudaGetLastError();
}

#ifdef MESM_TC_TIMING
void tc_read_times(long long* out64) { cudaMemcpyFromSymbol(out64, tc::g_tc_times, sizeof(long long) * 64); }
#endif

void tc_read_watchdog_linear(unsigned long long* out64) { cudaMemcpyFromSymbol(out64, tc::g_tc_watchdog, 512); unsigned long long z[64] = {0}; cudaMemcpyToSymbol(tc::g_tc_watchdog, z, 512); }

bool linear_tc_eligible(const LinearOp& op) {
    if (!op.Wp || (op.A2 && !op.Wp2) || op.nbatch > 1) return false;
    if (op.M < 128 || op.N < 64) return false;
    if (op.ln_g && op.N != 256) return false;
    if (op.act == ACT_SIGMOID) return false;
    if (op.fuse_rowstat && (op.Apos || op.A2 || !op.colsum)) return false;
    auto al16 = [](const float* p, long long ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0)); };
    if ((op.N & 3) || !al16(op.out, op.ldo) || !al16(op.out2, op.ldo2) || !al16(op.residual, op.ldr) || !al16(op.pre_ln, op.N)) return false;
    auto ok = [](const float* p, int ld, int K) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 7) == 0) && (ld % 2 == 0) && (K % 2 == 0)); };
    if (!ok(op.A, op.lda, op.K) || !ok(op.Apos, op.lda, op.K) || !ok(op.A2, op.lda2, op.K2)) return false;
    return true;
}

static int g_tc_pdl = -1;        // programmatic dependent launch of the tcgen05 linear kernels (MESM_TC_PDL=0 disables)

template <int VEC>
static cudaError_t launch_tc_variant(const LinearOp& op, int nkb1, int nkb2, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        MESM_CHECK(cudaFuncSetAttribute(tc::linear_tc_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        attr_set = true;
    }
    const unsigned mt = (unsigned)((op.M + tc::BM - 1) / tc::BM);
    cudaLaunchConfig_t cfg = {};
    const unsigned ntn = (unsigned)((op.N + tc::BN - 1) / tc::BN);
    cfg.gridDim = dim3(mt * ntn, 1, 1);
    cfg.blockDim = dim3(tc::THREADS, 1, 1);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_tc_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, tc::linear_tc_kernel<VEC>, op, nkb1, nkb2);
}

cudaError_t launch_linear_tc(const LinearOp& op, cudaStream_t s) {
    if (g_tc_pdl < 0) { const char* pe = getenv("MESM_TC_PDL"); g_tc_pdl = (pe && pe[0] == '0') ? 0 : 1; }
    const int nkb1 = (op.K + tc::BK - 1) / tc::BK, nkb2 = op.A2 ? (op.K2 + tc::BK - 1) / tc::BK : 0;
    auto v4 = [](const float* p, int ld, int K) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0) && (K % 4 == 0)); };
    const bool vec4 = v4(op.A, op.lda, op.K) && v4(op.Apos, op.lda, op.K) && v4(op.A2, op.lda2, op.K2);
    const cudaError_t e = vec4 ? launch_tc_variant<4>(op, nkb1, nkb2, s) : launch_tc_variant<2>(op, nkb1, nkb2, s);
    g_stats.launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mesm
