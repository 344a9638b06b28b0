// Attention cores of the MESM path (fp32 SIMT; QK^T / PV are ~2-5 % of the path's FLOPs).
//
//   mha_rows_kernel   one thread per query row, K_h / V_h of one (pair, head) resident in shared memory,
//                     blocked online softmax.  Used for the T2V cross-attention (<= 33 text keys, with the
//                     reference's attn_mask quirk) and the encoder self-attention (<= Lv+1 keys).
//   dec_self_attn     10 x 10 decoder self-attention, one CTA per pair, one warp per head.
//   dec_cross_attn    10 queries x Lv keys, per-head [content ; sine] operands (hd 64), value hd 32.
//   recon_pool        SS-MESM single-query attention evaluated on the *unprojected* clip rows
//                     (scores = (Wk_h^T q_h) . x_k, output = Wv_h (sum_k p_k x_k) + b): no K/V projection of the clips.
#include "kernels.h"
#include "tc_common.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace mesm {

// ---------------------------------------------------------------------------------------------------------------
// mha_rows_kernel
// ---------------------------------------------------------------------------------------------------------------

// 128 query rows per CTA (blockIdx.z tiles a pair's rows): with ~96 registers per thread that is 5 CTAs = 20 warps per SM,
// against 2 CTAs = 14 warps for one 224-thread CTA per (pair, head) (ncu: registers were the occupancy limiter, 16 % warps
// active), and short pairs of a packed batch launch no idle warps beyond their last tile.
constexpr int MR_ROWS = 128;
template <bool QUIRK>
__global__ void __launch_bounds__(MR_ROWS, 5) mha_rows_kernel(const MhaRowsArgs a) {
    extern __shared__ float smem[];
    const int h = blockIdx.x, b = blockIdx.y;
    long long kbase, qbase; int Lk, Lq;
    pair_rows(a.k_cu, a.k_enc, b, a.Lk, kbase, Lk);       // key rows of this pair (packed layouts: variable count)
    pair_rows(a.q_cu, a.q_enc, b, a.Lq, qbase, Lq);
    if ((int)blockIdx.z * MR_ROWS >= Lq) return;            // packed batch: this pair has no rows in this tile
    const int Lkmax = a.Lk;
    // every thread of a warp reads the SAME key row (broadcast): rows stay 16-byte aligned for LDS.128, no padding needed
    float* Ks = smem;                       // [Lk][32]
    float* Vs = Ks + Lkmax * 32;            // [Lk][32]
    uint8_t* pad_own = reinterpret_cast<uint8_t*>(Vs + Lkmax * 32);  // [Lk]
    uint8_t* pad_oth = pad_own + Lkmax;                               // [Lk]   k_pad of pair b'
    const int bg = a.b0 + b;                                           // global pair index
    const int bp = QUIRK ? (int)(((long long)bg * NH + h) % a.Btot) : bg;

    for (int idx = threadIdx.x; idx < Lk * 32; idx += blockDim.x) {
        const int kk = idx >> 5, j = idx & 31;
        const long long row = kbase + kk;
        Ks[kk * 32 + j] = a.k[row * a.ldk + h * 32 + j];
        Vs[kk * 32 + j] = a.v[row * a.ldv + h * 32 + j];
    }
    for (int kk = threadIdx.x; kk < Lk; kk += blockDim.x) {
        pad_own[kk] = a.k_pad[(a.k_cu ? kbase : (long long)bg * Lk) + kk];
        pad_oth[kk] = QUIRK ? a.k_pad[(long long)bp * Lk + kk] : 0;      // the quirk only exists with uniform (text) keys
    }
    __syncthreads();

    const int i = blockIdx.z * MR_ROWS + threadIdx.x;
    if (i >= Lq) return;
    const long long qrow = qbase + i;
    float q[32], o[32];
    {
        const float4* qp = reinterpret_cast<const float4*>(a.q + qrow * a.ldq + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = qp[j];
            q[4 * j] = t.x * a.q_scale; q[4 * j + 1] = t.y * a.q_scale; q[4 * j + 2] = t.z * a.q_scale; q[4 * j + 3] = t.w * a.q_scale;
        }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
    const bool qpad_oth = QUIRK ? (a.q_pad[(long long)bp * (a.q_pad_ld ? a.q_pad_ld : a.Lq) + i] != 0) : false;
    float m = -CUDART_INF_F, l = 0.f;

    for (int k0 = 0; k0 < Lk; k0 += 4) {
        float s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int kk = k0 + u;
            float acc = -CUDART_INF_F;
            if (kk < Lk) {
                const bool masked = pad_own[kk] || (qpad_oth && pad_oth[kk]);
                if (!masked) {
                    const float4* kr = reinterpret_cast<const float4*>(Ks + kk * 32);
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = kr[j];
                        a0 = fmaf(q[4 * j], t.x, a0); a1 = fmaf(q[4 * j + 1], t.y, a1);
                        a2 = fmaf(q[4 * j + 2], t.z, a2); a3 = fmaf(q[4 * j + 3], t.w, a3);
                    }
                    acc = (a0 + a1) + (a2 + a3);
                }
            }
            s[u] = acc;
        }
        const float bm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
        if (bm == -CUDART_INF_F) continue;
        const float mn = fmaxf(m, bm);
        const float sc = __expf(m - mn);      // m = -inf on first block -> 0
        l *= sc;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] *= sc;
        m = mn;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (s[u] == -CUDART_INF_F) continue;
            const float p = __expf(s[u] - mn);
            l += p;
            const float4* vr = reinterpret_cast<const float4*>(Vs + (k0 + u) * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = vr[j];
                o[4 * j] = fmaf(p, t.x, o[4 * j]); o[4 * j + 1] = fmaf(p, t.y, o[4 * j + 1]);
                o[4 * j + 2] = fmaf(p, t.z, o[4 * j + 2]); o[4 * j + 3] = fmaf(p, t.w, o[4 * j + 3]);
            }
        }
    }
    const float inv = 1.f / l;                 // l == 0 (all keys masked) -> inf*0 = NaN like the reference
    if (a.out) {
        float4* op = reinterpret_cast<float4*>(a.out + qrow * a.ldo + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(o[4 * j] * inv, o[4 * j + 1] * inv, o[4 * j + 2] * inv, o[4 * j + 3] * inv);
    }
    if (a.out_hi) {                            // pre-split planes for the output projection (row pitch 256 elements)
        uint4* ph = reinterpret_cast<uint4*>(a.out_hi + qrow * D + h * 32);
        uint4* pl = reinterpret_cast<uint4*>(a.out_lo + qrow * D + h * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 hh, ll;
            tc::split_bf16x2(o[8 * j] * inv, o[8 * j + 1] * inv, hh.x, ll.x);
            tc::split_bf16x2(o[8 * j + 2] * inv, o[8 * j + 3] * inv, hh.y, ll.y);
            tc::split_bf16x2(o[8 * j + 4] * inv, o[8 * j + 5] * inv, hh.z, ll.z);
            tc::split_bf16x2(o[8 * j + 6] * inv, o[8 * j + 7] * inv, hh.w, ll.w);
            ph[j] = hh; pl[j] = ll;
        }
    }
}

cudaError_t launch_mha_rows(const MhaRowsArgs& a, cudaStream_t s, bool force_simt) {
    if (a.B <= 0 || a.Lq <= 0) return cudaSuccess;
    {
        static int force = -1;
        if (force < 0) { const char* e = getenv("MESM_FORCE_SIMT"); const char* e2 = getenv("MESM_FORCE_SIMT_ATTN"); force = ((e && e[0] == '1') || (e2 && e2[0] == '1')) ? 1 : 0; }
        static int mma = -1;
        if (mma < 0) { const char* e = getenv("MESM_ATTN_MMA"); mma = (e && e[0] == '0') ? 0 : 1; }
        if (!force && !force_simt && mma && attn_mma_eligible(a)) return launch_attn_mma(a, s);
        if (!force && !force_simt && attn_tc_eligible(a)) return launch_attn_tc(a, s);
        if (!force && !force_simt && attn_tc_split_eligible(a)) {
            // rows of the query side = rows of the output: uniform B * Lq, packed: the caller passes it through split_rows
            return launch_attn_tc_split(a, a.split_rows, s);
        }
    }
    ProfScope _ps(a.q_pad ? "mha_rows t2v" : "mha_rows self", s);
    const int threads = MR_ROWS;
    const size_t smem = (size_t)a.Lk * 32 * 2 * sizeof(float) + 2 * (size_t)a.Lk;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    dim3 grid(NH, a.B, (a.Lq + MR_ROWS - 1) / MR_ROWS);
    if (a.q_pad) {
        if (smem > 48 * 1024) MESM_CHECK(cudaFuncSetAttribute(mha_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mha_rows_kernel<true><<<grid, threads, smem, s>>>(a);
    } else {
        if (smem > 48 * 1024) MESM_CHECK(cudaFuncSetAttribute(mha_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mha_rows_kernel<false><<<grid, threads, smem, s>>>(a);
    }
    g_stats.launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Generic projection-free MHA (model/attention.py:185-394) for the drop-in MultiheadAttention class and the decoder:
// q rows (b*L + i) with E = nheads*hq columns, k rows (b*S + j), v rows with nheads*hv columns.  One CTA per
// (head, pair); thread-per-key scores into shared memory, warp-per-query softmax, thread-per-(query,dim) PV.
// ---------------------------------------------------------------------------------------------------------------

__global__ void mha_small_kernel(const MhaSmallArgs a) {
    extern __shared__ float smem[];
    const int h = blockIdx.x, b = blockIdx.y;
    const int L = a.L, hq = a.hq, hv = a.hv;
    int S = a.S;
    long long kfirst = (long long)b * a.k_bs + a.k_off, kpad0 = (long long)b * a.S;      // first key row / first k_pad entry
    if (a.k_cu) {                                                                        // packed keys (kernels.h)
        const int c0 = a.k_cu[b] - a.k_cu[0];
        S = a.k_cu[b + 1] - a.k_cu[b];
        kfirst = (long long)c0 + (a.k_enc ? b : 0) + a.k_off;
        kpad0 = c0;
    }
    const int nparts = a.q2 ? 2 : 1;
    const int E = hq * nparts;
    float* Qs = smem;                 // [L][E]
    float* Sc = Qs + L * E;           // [L][S]  (row pitch = this pair's S)
    for (int idx = threadIdx.x; idx < L * E; idx += blockDim.x) {
        const int i = idx / E, c = idx % E;
        const long long row = (long long)b * a.q_bs + (long long)i * a.q_is;
        const float val = c < hq ? a.q[row * a.ldq + h * hq + c] : a.q2[row * a.ldq2 + h * hq + (c - hq)];
        Qs[idx] = val * a.scale;
    }
    __syncthreads();
    // scores: thread per key
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const bool masked = a.k_pad && a.k_pad[kpad0 + j];
        const long long krow = kfirst + (long long)j * a.k_is;
        for (int i0 = 0; i0 < L; i0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (!masked) {
                for (int part = 0; part < nparts; ++part) {
                    const float* kp = part == 0 ? a.k + krow * a.ldk + h * hq
                                                : a.k2 + (a.k2_table ? (long long)a.k2_table[krow] : krow) * a.ldk2 + h * hq;
                    for (int c = 0; c < hq; c += 4) {
                        const float4 kv = *reinterpret_cast<const float4*>(kp + c);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (i0 + u < L) {
                                const float* qr = Qs + (i0 + u) * E + part * hq + c;
                                acc[u] = fmaf(qr[0], kv.x, fmaf(qr[1], kv.y, fmaf(qr[2], kv.z, fmaf(qr[3], kv.w, acc[u]))));
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u < L) Sc[(i0 + u) * S + j] = (masked || (a.causal && j > i0 + u)) ? -CUDART_INF_F : acc[u];
        }
    }
    __syncthreads();
    // softmax: warp per query row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = warp; i < L; i += nwarps) {
        float* row = Sc + i * S;
        float m = -CUDART_INF_F;
        for (int j = lane; j < S; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < S; j += 32) { const float p = __expf(row[j] - m); row[j] = p; sum += p; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int j = lane; j < S; j += 32) {
            const float p = row[j] * inv;
            row[j] = p;
            if (a.attn_w) atomicAdd(a.attn_w + ((long long)b * L + i) * a.S + j, p / a.nheads);
        }
    }
    __syncthreads();
    // PV: thread per (query, dim)
    for (int idx = threadIdx.x; idx < L * hv; idx += blockDim.x) {
        const int i = idx / hv, c = idx % hv;
        const float* prow = Sc + i * S;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) {
            const long long vrow = kfirst + (long long)j * a.k_is;
            acc = fmaf(prow[j], a.v[vrow * a.ldv + h * hv + c], acc);
        }
        a.out[((long long)b * a.q_bs + (long long)i * a.q_is) * a.ldo + h * hv + c] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Decoder cross-attention (model/transformer.py:770-790): LQ queries x S memory keys, per-head operands
// [content ; sine] (2 x 32) and 32-wide values.  One CTA per pair, one WARP per head; keys in blocks of 32 with the lane
// as the key for the scores (its 2 x 128-byte key segments live in registers for all LQ queries - every key byte is read
// once) and the lane as the output dimension for P V (value rows read coalesced, straight from global memory).  Online
// softmax per query; no block-wide barrier after the query tile is staged.
// ---------------------------------------------------------------------------------------------------------------
template <int LQ>
__global__ void __launch_bounds__(256) dec_cross_kernel(const MhaSmallArgs a) {
    extern __shared__ float smem[];
    float* Qs = smem;                                  // [LQ][2][256]  scaled content / sine queries
    float* Ps = Qs + LQ * 512;                         // [8 warps][LQ][32]  probabilities of the current key block
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int S = a.S;
    long long kfirst = (long long)b * a.k_bs + a.k_off, kpad0 = (long long)b * a.S;
    if (a.k_cu) {
        const int c0 = a.k_cu[b] - a.k_cu[0];
        S = a.k_cu[b + 1] - a.k_cu[b];
        kfirst = (long long)c0 + (a.k_enc ? b : 0) + a.k_off;
        kpad0 = c0;
    }
    for (int idx = threadIdx.x; idx < LQ * 128; idx += 256) {
        const int i = idx >> 7, part = (idx >> 6) & 1, c4 = (idx & 63) * 4;
        const long long row = (long long)b * a.q_bs + (long long)i * a.q_is;
        float4 v = *reinterpret_cast<const float4*>((part ? a.q2 + row * a.ldq2 : a.q + row * a.ldq) + c4);
        v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
        *reinterpret_cast<float4*>(Qs + i * 512 + part * 256 + c4) = v;
    }
    __syncthreads();
    float* P = Ps + h * (LQ * 32);
    float m[LQ], l[LQ], acc[LQ];
#pragma unroll
    for (int i = 0; i < LQ; ++i) { m[i] = -CUDART_INF_F; l[i] = 0.f; acc[i] = 0.f; }

    for (int k0 = 0; k0 < S; k0 += 32) {
        const int key = k0 + lane;
        const bool valid = key < S && !(a.k_pad && a.k_pad[kpad0 + key]);
        float s[LQ];
#pragma unroll
        for (int i = 0; i < LQ; ++i) s[i] = 0.f;
        if (valid) {
            const long long krow = kfirst + (long long)key * a.k_is;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const float* kp = part == 0 ? a.k + krow * a.ldk + h * 32
                                            : a.k2 + (a.k2_table ? (long long)a.k2_table[krow] : krow) * a.ldk2 + h * 32;
                float kr[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 t = *reinterpret_cast<const float4*>(kp + 4 * j);
                    kr[4 * j] = t.x; kr[4 * j + 1] = t.y; kr[4 * j + 2] = t.z; kr[4 * j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < LQ; ++i) {
                    const float4* q4 = reinterpret_cast<const float4*>(Qs + i * 512 + part * 256 + h * 32);   // broadcast reads
                    float d = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = q4[j];
                        d = fmaf(t.x, kr[4 * j], d); d = fmaf(t.y, kr[4 * j + 1], d);
                        d = fmaf(t.z, kr[4 * j + 2], d); d = fmaf(t.w, kr[4 * j + 3], d);
                    }
                    s[i] += d;
                }
            }
        }
        __syncwarp();                                  // the previous block's P has been consumed
#pragma unroll
        for (int i = 0; i < LQ; ++i) {
            const float sv = valid ? s[i] : -CUDART_INF_F;
            const float mn = fmaxf(m[i], warp_max(sv));
            float corr = 1.f, p = 0.f;
            if (mn != -CUDART_INF_F) { corr = __expf(m[i] - mn); p = valid ? __expf(sv - mn) : 0.f; }   // m = -inf -> corr 0
            l[i] = l[i] * corr + warp_sum(p);
            acc[i] *= corr;
            m[i] = mn;
            P[i * 32 + lane] = p;
        }
        __syncwarp();
        const int nk = min(32, S - k0);
        for (int j0 = 0; j0 < nk; j0 += 4) {           // lane = output dim; 4 keys per step, P read as broadcast float4
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int jj = j0 + u;
                v[u] = jj < nk ? a.v[(kfirst + (long long)(k0 + jj) * a.k_is) * a.ldv + h * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < LQ; ++i) {
                const float4 p4 = *reinterpret_cast<const float4*>(P + i * 32 + j0);
                acc[i] = fmaf(p4.x, v[0], fmaf(p4.y, v[1], fmaf(p4.z, v[2], fmaf(p4.w, v[3], acc[i]))));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LQ; ++i)
        a.out[((long long)b * a.q_bs + (long long)i * a.q_is) * a.ldo + h * 32 + lane] = acc[i] / l[i];   // l == 0 -> NaN like the reference
}

static bool dec_cross_eligible(const MhaSmallArgs& a) {
    auto al16 = [](const float* p, int ld) { return p && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 4 == 0); };
    return a.L == 10 && a.nheads == NH && a.hq == HD && a.hv == HD && a.q2 && a.k2 && !a.attn_w && al16(a.q, a.ldq) &&
           al16(a.q2, a.ldq2) && al16(a.k, a.ldk) && al16(a.k2, a.ldk2) && a.S >= 1;
}

cudaError_t launch_mha_small(const MhaSmallArgs& a, cudaStream_t s) {
    if (a.B <= 0 || a.L <= 0) return cudaSuccess;
    static int cross_on = -1;
    if (cross_on < 0) { const char* e = getenv("MESM_DEC_CROSS"); cross_on = (e && e[0] == '0') ? 0 : 1; }
    static int cross_mma = -1;
    if (cross_mma < 0) { const char* e = getenv("MESM_DEC_CROSS_MMA"); cross_mma = (e && e[0] == '0') ? 0 : 1; }
    if (cross_on && cross_mma && dec_cross_mma_eligible(a)) return launch_dec_cross_mma(a, s);      // decoder cross- and self-attention (<= 16 queries)
    if (cross_on && !a.causal && dec_cross_eligible(a)) {
        ProfScope _ps("dec_cross", s);
        const size_t smem = (size_t)(10 * 512 + 8 * 10 * 32) * sizeof(float);
        dec_cross_kernel<10><<<a.B, 256, smem, s>>>(a);
        g_stats.launches++;
        return cudaGetLastError();
    }
    ProfScope _ps("mha_small", s);
    const int E = a.hq * (a.q2 ? 2 : 1);
    const size_t smem = ((size_t)a.L * E + (size_t)a.L * a.S) * sizeof(float);
    if (smem > 220 * 1024 || (a.hq & 3)) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) MESM_CHECK(cudaFuncSetAttribute(mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(a.nheads, a.B);
    mha_small_kernel<<<grid, 128, smem, s>>>(a);
    g_stats.launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// recon_pool: one masked sentence slot per pair attends over the projected clips of its video group (SegSenRecon,
// model/model.py:467-488) — evaluated without projecting the clips:
//   score[h][k] = qk[b,h,:] . x_k          (qk = Wk_h^T q_h, built by the caller; the q.bk term cancels in softmax)
//   pooled[b,h,:] = sum_k softmax_k(score[h][.]) x_k     (caller applies Wv_h and bv)
// Key set of pair b: its own Lv rows (charades / tacos branch, model.py:186-189) or the valid rows of every pair of
// its group, concatenated (qvhighlights branch, model.py:191-195).  Mask = reference quirk (see t2v_mask in the
// oracle): masked = kpad[b,k] | (qpad[b',slot_b] & kpad[b',k]), b' = (b*H+h) % B, qpad[b',s] = s >= num_clips[g(b')].
// ---------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) recon_pool_kernel(const ReconPoolArgs a) {
    extern __shared__ float smem[];
    float* Sc = smem;                                  // [8][max_keys]
    int* krow = reinterpret_cast<int*>(Sc + 8 * a.max_keys);   // [max_keys] global row index of key k (-1 = padded)
    __shared__ int s_nkeys;
    const int bl = blockIdx.x, b = a.b0 + blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = a.pair_group[b], slot = a.pair_slot[b];
    const int Lv = a.Lv;

    // key list
    if (a.qvh) {
        if (threadIdx.x == 0) {
            int n = 0;
            for (int p = a.group_start[g]; p < a.group_start[g + 1]; ++p)
                for (int i = 0; i < Lv; ++i)
                    if (a.vmask[(long long)p * Lv + i]) krow[n++] = a.x_start ? a.x_start[p] + i : (p - a.b0) * Lv + i;
            s_nkeys = n;
        }
    } else {
        for (int i = threadIdx.x; i < Lv; i += blockDim.x) krow[i] = a.vmask[(long long)b * Lv + i] ? (a.x_start ? a.x_start[b] + i : bl * Lv + i) : -1;
        if (threadIdx.x == 0) s_nkeys = Lv;
    }
    __syncthreads();
    const int nk = s_nkeys;

    // quirk partner of this (pair, head)
    const int bp = (int)(((long long)b * NH + h) % a.Btot);
    const int gp = a.pair_group[bp];
    const bool qpad_oth = slot >= (a.group_start[gp + 1] - a.group_start[gp]);

    float qv[8];
    {
        const float4* qp = reinterpret_cast<const float4*>(a.qk + ((long long)bl * NH + h) * D + lane * 8);
        const float4 t0 = qp[0], t1 = qp[1];
        qv[0] = t0.x; qv[1] = t0.y; qv[2] = t0.z; qv[3] = t0.w; qv[4] = t1.x; qv[5] = t1.y; qv[6] = t1.z; qv[7] = t1.w;
    }
    float* sc = Sc + h * a.max_keys;
    for (int k = 0; k < nk; ++k) {
        const int r = krow[k];
        bool masked = r < 0;
        if (!masked && qpad_oth) {
            const bool kpad_oth = a.qvh ? (k >= a.group_len[gp]) : (a.vmask[(long long)bp * Lv + k] == 0);
            masked = kpad_oth;
        }
        float s = -CUDART_INF_F;
        if (!masked) {
            const float4* xp = reinterpret_cast<const float4*>(a.x + (long long)r * a.ldx + lane * 8);
            const float4 t0 = xp[0], t1 = xp[1];
            float d = qv[0] * t0.x;
            d = fmaf(qv[1], t0.y, d); d = fmaf(qv[2], t0.z, d); d = fmaf(qv[3], t0.w, d);
            d = fmaf(qv[4], t1.x, d); d = fmaf(qv[5], t1.y, d); d = fmaf(qv[6], t1.z, d); d = fmaf(qv[7], t1.w, d);
            s = warp_sum(d);
        }
        if (lane == 0) sc[k] = s;
    }
    __syncwarp();
    float m = -CUDART_INF_F;
    for (int k = lane; k < nk; k += 32) m = fmaxf(m, sc[k]);
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < nk; k += 32) { const float p = __expf(sc[k] - m); sc[k] = p; sum += p; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nk; ++k) {
        const float p = sc[k];
        if (p == 0.f) continue;                        // warp-uniform
        const float4* xp = reinterpret_cast<const float4*>(a.x + (long long)krow[k] * a.ldx + lane * 8);
        const float4 t0 = xp[0], t1 = xp[1];
        acc[0] = fmaf(p, t0.x, acc[0]); acc[1] = fmaf(p, t0.y, acc[1]); acc[2] = fmaf(p, t0.z, acc[2]); acc[3] = fmaf(p, t0.w, acc[3]);
        acc[4] = fmaf(p, t1.x, acc[4]); acc[5] = fmaf(p, t1.y, acc[5]); acc[6] = fmaf(p, t1.z, acc[6]); acc[7] = fmaf(p, t1.w, acc[7]);
    }
    float4* op = reinterpret_cast<float4*>(a.pooled + ((long long)bl * NH + h) * D + lane * 8);
    op[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    op[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

// Tiled variant (round 2).  The kernel above keeps one warp per head walking the keys one at a time: every key costs a 1 KB row
// load per head (8x per pair, twice: scores and pooling) and a serial 5-step warp reduction.  Here the CTA stages 32 clip rows
// ONCE in shared memory for all 8 heads and both uses, computes the 8 x 32 scores with the Wk^T q vectors in registers (warp w:
// 4 keys; lane: 8 columns x 8 heads = 64 FMAs per key, then a 9-shuffle transpose-reduce instead of 8 x 5), runs the softmax
// online per tile (warp h = head h), and pools with thread = column, 8 heads in registers.
constexpr int RP_TK = 32, RP_LD = 260;
__global__ void __launch_bounds__(256) recon_pool_tiled_kernel(const ReconPoolArgs a) {
    extern __shared__ float smem[];
    float* xt = smem;                                   // [32][260]  clip rows of the tile (row stride 260: LDS.128 conflict-free)
    float* sc = xt + RP_TK * RP_LD;                     // [32 keys][8 heads] scores, then probabilities
    int* krow = reinterpret_cast<int*>(sc + RP_TK * 8); // [max_keys] global row of key k (-1 = padded)
    __shared__ int s_nkeys, s_bp[8], s_qpo[8], s_glp[8];
    __shared__ float s_alpha[8], s_l[8];
    const int bl = blockIdx.x, b = a.b0 + blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = a.pair_group[b], slot = a.pair_slot[b];
    const int Lv = a.Lv;
    if (a.qvh) {
        if (threadIdx.x == 0) {
            int n = 0;
            for (int p = a.group_start[g]; p < a.group_start[g + 1]; ++p)
                for (int i = 0; i < Lv; ++i)
                    if (a.vmask[(long long)p * Lv + i]) krow[n++] = a.x_start ? a.x_start[p] + i : (p - a.b0) * Lv + i;
            s_nkeys = n;
        }
    } else {
        for (int i = threadIdx.x; i < Lv; i += blockDim.x) krow[i] = a.vmask[(long long)b * Lv + i] ? (a.x_start ? a.x_start[b] + i : bl * Lv + i) : -1;
        if (threadIdx.x == 0) s_nkeys = Lv;
    }
    if (threadIdx.x < 8) {                              // quirk partner of (pair, head h): see the header comment of this section
        const int h = threadIdx.x;
        const int bp = (int)(((long long)b * NH + h) % a.Btot);
        const int gp = a.pair_group[bp];
        s_bp[h] = bp;
        s_qpo[h] = slot >= (a.group_start[gp + 1] - a.group_start[gp]) ? 1 : 0;
        s_glp[h] = a.qvh ? a.group_len[gp] : 0;
    }
    __syncthreads();
    const int nk = s_nkeys;
    float qk[8][8];                                     // (Wk_h^T q_h)[lane * 8 + j] for the 8 heads
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        const float4* qp = reinterpret_cast<const float4*>(a.qk + ((long long)bl * NH + h) * D + lane * 8);
        const float4 t0 = qp[0], t1 = qp[1];
        qk[h][0] = t0.x; qk[h][1] = t0.y; qk[h][2] = t0.z; qk[h][3] = t0.w; qk[h][4] = t1.x; qk[h][5] = t1.y; qk[h][6] = t1.z; qk[h][7] = t1.w;
    }
    float acc[8];                                       // pooled[h][c], c = threadIdx.x
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[h] = 0.f;
    float m_run = -CUDART_INF_F, l_run = 0.f;           // online softmax state of head w (uniform over the warp)
    const int c = threadIdx.x;

    for (int k0 = 0; k0 < nk; k0 += RP_TK) {
        const int tk = min(RP_TK, nk - k0);
        {   // stage the tile: thread -> (row t >> 3, float4 columns (t & 7) + 8 i)
            const int r = threadIdx.x >> 3;
            const int gr = (r < tk) ? krow[k0 + r] : -1;
            const float4* src = gr >= 0 ? reinterpret_cast<const float4*>(a.x + (long long)gr * a.ldx) : nullptr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c4 = (threadIdx.x & 7) + 8 * i;
                const float4 v = src ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(&xt[r * RP_LD + c4 * 4]) = v;
            }
        }
        __syncthreads();
        // ---- scores: warp w -> keys 4w .. 4w+3 of the tile, all 8 heads ----
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
            const int kk = 4 * w + u;
            if (kk >= tk) break;                        // warp-uniform
            const float4 x0 = *reinterpret_cast<const float4*>(&xt[kk * RP_LD + lane * 8]);
            const float4 x1 = *reinterpret_cast<const float4*>(&xt[kk * RP_LD + lane * 8 + 4]);
            float p[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                float d = qk[h][0] * x0.x;
                d = fmaf(qk[h][1], x0.y, d); d = fmaf(qk[h][2], x0.z, d); d = fmaf(qk[h][3], x0.w, d);
                d = fmaf(qk[h][4], x1.x, d); d = fmaf(qk[h][5], x1.y, d); d = fmaf(qk[h][6], x1.z, d); d = fmaf(qk[h][7], x1.w, d);
                p[h] = d;
            }
            // transpose-reduce: 8 partials x 32 lanes -> every lane ends with the full sum of head (bit4, bit3, bit2 of its lane id)
            {
                const bool up = (lane & 16) != 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float send = up ? p[i] : p[i + 4]; const float recv = __shfl_xor_sync(0xffffffffu, send, 16); p[i] = (up ? p[i + 4] : p[i]) + recv; }
            }
            {
                const bool up = (lane & 8) != 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) { const float send = up ? p[i] : p[i + 2]; const float recv = __shfl_xor_sync(0xffffffffu, send, 8); p[i] = (up ? p[i + 2] : p[i]) + recv; }
            }
            {
                const bool up = (lane & 4) != 0;
                const float send = up ? p[0] : p[1];
                const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
                p[0] = (up ? p[1] : p[0]) + recv;
            }
            p[0] += __shfl_xor_sync(0xffffffffu, p[0], 2);
            p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
            if ((lane & 3) == 0) {
                const int h = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                const int k = k0 + kk;
                bool masked = krow[k] < 0;
                if (!masked && s_qpo[h]) masked = a.qvh ? (k >= s_glp[h]) : (a.vmask[(long long)s_bp[h] * Lv + k] == 0);
                sc[kk * 8 + h] = masked ? -CUDART_INF_F : p[0];
            }
        }
        __syncthreads();
        // ---- online softmax of head w over the tile's keys (lane = key) ----
        {
            const float sv = lane < tk ? sc[lane * 8 + w] : -CUDART_INF_F;
            const float tm = warp_max(sv);
            float alpha = 1.f, pv = 0.f;
            if (tm != -CUDART_INF_F) {
                const float mn = fmaxf(m_run, tm);
                alpha = __expf(m_run - mn);             // first tile: exp(-inf) = 0
                pv = sv == -CUDART_INF_F ? 0.f : __expf(sv - mn);
                l_run = l_run * alpha + warp_sum(pv);
                m_run = mn;
            }
            if (lane < tk) sc[lane * 8 + w] = pv;
            if (lane == 0) s_alpha[w] = alpha;
        }
        __syncthreads();
        // ---- pooling: thread = column c, 8 heads in registers ----
#pragma unroll
        for (int h = 0; h < 8; ++h) acc[h] *= s_alpha[h];
        for (int kk = 0; kk < tk; ++kk) {
            const float x = xt[kk * RP_LD + c];
            const float4 p0 = *reinterpret_cast<const float4*>(&sc[kk * 8]);
            const float4 p1 = *reinterpret_cast<const float4*>(&sc[kk * 8 + 4]);
            acc[0] = fmaf(p0.x, x, acc[0]); acc[1] = fmaf(p0.y, x, acc[1]); acc[2] = fmaf(p0.z, x, acc[2]); acc[3] = fmaf(p0.w, x, acc[3]);
            acc[4] = fmaf(p1.x, x, acc[4]); acc[5] = fmaf(p1.y, x, acc[5]); acc[6] = fmaf(p1.z, x, acc[6]); acc[7] = fmaf(p1.w, x, acc[7]);
        }
        __syncthreads();                                // the next tile overwrites xt / sc
    }
    if (lane == 0) s_l[w] = l_run;
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 8; ++h) a.pooled[((long long)bl * NH + h) * D + c] = acc[h] * (1.f / s_l[h]);     // l == 0 -> NaN like the reference
}

cudaError_t launch_recon_pool(const ReconPoolArgs& a, cudaStream_t s) {
    ProfScope _ps("recon_pool", s);
    if (a.B <= 0) return cudaSuccess;
    static int tiled = -1;
    if (tiled < 0) { const char* e = getenv("MESM_RECON_TILED"); tiled = (e && e[0] == '0') ? 0 : 1; }
    if (tiled && (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {
        const size_t smem = (size_t)(RP_TK * RP_LD + RP_TK * 8) * sizeof(float) + (size_t)a.max_keys * sizeof(int);
        if (smem <= 200 * 1024) {
            if (smem > 48 * 1024) MESM_CHECK(cudaFuncSetAttribute(recon_pool_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            recon_pool_tiled_kernel<<<a.B, 256, smem, s>>>(a);
            g_stats.launches++;
            return cudaGetLastError();
        }
    }
    const size_t smem = (size_t)a.max_keys * (8 * sizeof(float) + sizeof(int));
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) MESM_CHECK(cudaFuncSetAttribute(recon_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    recon_pool_kernel<<<a.B, 256, smem, s>>>(a);
    g_stats.launches++;
    return cudaGetLastError();
}

}  // namespace mesm
