// Row-wise / element-wise kernels of the MESM path (all HBM- or latency-bound; vectorised, coalesced, one warp per row
// where a row reduction is needed).
#include "kernels.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

namespace mesm {

#define LAUNCH_END()            \
    g_stats.launches++;         \
    return cudaGetLastError()

static inline int blocks_for(long long n, int per) { return (int)((n + per - 1) / per); }

// ---- text post-processing: F.normalize(eps 1e-5), mask = rowsum != 0 (model/model.py:145-152) + LN statistics ------
__global__ void text_prep_kernel(const float* __restrict__ x, int R, int Dt, float* __restrict__ y,
                                 uint8_t* __restrict__ mask, float* __restrict__ rowstat) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + (long long)r * Dt;
    float ss = 0.f;
    for (int c = lane; c < Dt; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-5f);
    float s = 0.f;
    for (int c = lane; c < Dt; c += 32) s += xr[c] * inv;
    s = warp_sum(s);
    const float mean = s / Dt;
    float q = 0.f;
    for (int c = lane; c < Dt; c += 32) { const float v = xr[c] * inv; y[(long long)r * Dt + c] = v; const float d = v - mean; q = fmaf(d, d, q); }
    q = warp_sum(q);
    if (lane == 0) {
        mask[r] = (s != 0.f) ? 1 : 0;
        rowstat[2 * r] = mean;
        rowstat[2 * r + 1] = rsqrtf(q / Dt + 1e-5f);
    }
}
cudaError_t launch_text_prep(const float* x, int R, int Dt, float* y, uint8_t* mask, float* rowstat, cudaStream_t s) {
    ProfScope _ps("text_prep", s);
    if (R <= 0) return cudaSuccess;
    text_prep_kernel<<<blocks_for(R, 8), 256, 0, s>>>(x, R, Dt, y, mask, rowstat);
    LAUNCH_END();
}

// ---- LayerNorm statistics of raw feature rows (LinearLayer's LayerNorm(in_hsz), model/model.py:427-431) -----------
__global__ void row_stats_kernel(const float* __restrict__ x, long long R, int Dv, int ldx, float* __restrict__ rowstat,
                                 const int* __restrict__ table) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + (table ? (long long)table[r] : r) * ldx;      // table: statistics of gathered rows (packed layout)
    float s = 0.f;
    if ((ldx & 1) == 0 && (Dv & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) {
        const float2* x2 = reinterpret_cast<const float2*>(xr);
        const int n2 = Dv >> 1;
        for (int c = lane; c < n2; c += 32) { const float2 v = x2[c]; s += v.x + v.y; }
        s = warp_sum(s);
        const float mean = s / Dv;
        float q = 0.f;
        for (int c = lane; c < n2; c += 32) { const float2 v = x2[c]; const float a = v.x - mean, b = v.y - mean; q = fmaf(a, a, q); q = fmaf(b, b, q); }
        q = warp_sum(q);
        if (lane == 0) { rowstat[2 * r] = mean; rowstat[2 * r + 1] = rsqrtf(q / Dv + 1e-5f); }
    } else {
        for (int c = lane; c < Dv; c += 32) s += xr[c];
        s = warp_sum(s);
        const float mean = s / Dv;
        float q = 0.f;
        for (int c = lane; c < Dv; c += 32) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
        q = warp_sum(q);
        if (lane == 0) { rowstat[2 * r] = mean; rowstat[2 * r + 1] = rsqrtf(q / Dv + 1e-5f); }
    }
}
cudaError_t launch_row_stats(const float* x, long long R, int Dv, int ldx, float* rowstat, cudaStream_t s, const int* table) {
    ProfScope _ps("row_stats", s);
    if (R <= 0) return cudaSuccess;
    row_stats_kernel<<<blocks_for(R, 8), 256, 0, s>>>(x, R, Dv, ldx, rowstat, table);
    LAUNCH_END();
}

// ---- PositionEmbeddingSine(normalize=True) (model/position_encoding.py:51-72) + encoder-layout side buffers --------
// posV [B,Lv,256]; posE [B,Lv+1,256] with global_rep_pos in row 0; padV [B,Lv], padE [B,Lv+1] (1 = pad; the global
// token's flag is 1: model/transformer.py:185-186); tokE row 0 of the [B,Lv+1,256] encoder buffer = global_rep_token.
__global__ void __launch_bounds__(256) pos_embed_kernel(const PosArgs a) {
    extern __shared__ float xemb[];     // [Lv]
    const int b = blockIdx.x;
    long long vrow0; int Lv;                                   // this pair's clip rows in the (possibly packed) buffers
    pair_rows(a.cu, 0, b, a.Lv, vrow0, Lv);
    const long long erow0 = vrow0 + b;                         // its encoder rows: global token first (uniform: b * (Lv + 1))
    if (a.enc_hi && threadIdx.x < D) {                         // the global token also as bf16 hi / lo planes of the encoder buffer
        const float g = a.gtok[threadIdx.x];
        const __nv_bfloat16 h = __float2bfloat16_rn(g);
        a.enc_hi[erow0 * D + threadIdx.x] = __bfloat16_as_ushort(h);
        a.enc_lo[erow0 * D + threadIdx.x] = __bfloat16_as_ushort(__float2bfloat16_rn(g - __bfloat162float(h)));
    }
    if (!a.posV && !a.posE && !a.padV && !a.padE) {          // only the global token row of the encoder buffer
        if (a.encbuf) a.encbuf[erow0 * D + threadIdx.x] = a.gtok[threadIdx.x];
        return;
    }
    const uint8_t* m = a.vmask + (long long)b * a.Lv;          // the mask itself is always the zero-padded [B, Lv] array
    const bool need_pos = a.posV || a.posE;
    for (int i = threadIdx.x; i < Lv; i += blockDim.x) {
        int c = 0;
        if (need_pos) for (int j = 0; j <= i; ++j) c += m[j] ? 1 : 0;
        xemb[i] = (float)c;
        const uint8_t pad = m[i] ? 0 : 1;
        if (a.padV) a.padV[vrow0 + i] = pad;
        if (a.padE) a.padE[erow0 + 1 + i] = pad;
    }
    if (threadIdx.x == 0 && a.padE) a.padE[erow0] = 1;
    if (threadIdx.x < D && a.encbuf && !a.posV && !a.posE) a.encbuf[erow0 * D + threadIdx.x] = a.gtok[threadIdx.x];
    if (!need_pos) return;                                     // pads (+ global token) only: positions come from the table
    __syncthreads();
    const float last = (Lv > 0 ? xemb[Lv - 1] : 0.f) + 1e-6f;   // packed: rows past this pair's length hold no valid clip
    const int c = threadIdx.x;          // 256 threads = 256 feature dims
    const float dim_t = powf(10000.f, (float)(2 * (c / 2)) / 256.f);
    for (int i = 0; i < Lv; ++i) {
        const float xe = xemb[i] / last * 6.283185307179586f;
        const float arg = xe / dim_t;
        const float v = (c & 1) ? cosf(arg) : sinf(arg);
        if (a.posV) a.posV[(vrow0 + i) * D + c] = v;
        if (a.posE) a.posE[(erow0 + 1 + i) * D + c] = v;
    }
    if (a.posE) a.posE[erow0 * D + c] = a.gpos[c];
    if (a.encbuf) a.encbuf[erow0 * D + c] = a.gtok[c];
}
cudaError_t launch_pos_embed(const PosArgs& a, cudaStream_t s) {
    ProfScope _ps("pos_embed", s);
    if (a.B <= 0) return cudaSuccess;
    pos_embed_kernel<<<a.B, 256, a.Lv * sizeof(float), s>>>(a);
    LAUNCH_END();
}

// ---- generic small helpers ---------------------------------------------------------------------------------------
__global__ void invert_mask_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] ? 0 : 1;
}
cudaError_t launch_invert_mask(const uint8_t* in, uint8_t* out, long long n, cudaStream_t s) {
    ProfScope _ps("invert_mask", s);
    if (n <= 0) return cudaSuccess;
    invert_mask_kernel<<<blocks_for(n, 256), 256, 0, s>>>(in, out, n);
    LAUNCH_END();
}

// expanded words mask / pad: [B,Lt+1] with a leading always-valid recon slot (model/model.py:218-219)
__global__ void expand_mask_kernel(const uint8_t* __restrict__ wmask, int B, int Lt, uint8_t* __restrict__ emask,
                                   uint8_t* __restrict__ epad, uint8_t* __restrict__ wpad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * (Lt + 1)) return;
    const int b = idx / (Lt + 1), j = idx % (Lt + 1);
    const uint8_t v = j == 0 ? 1 : wmask[b * Lt + j - 1];
    if (emask) emask[idx] = v;
    if (epad) epad[idx] = v ? 0 : 1;
    if (wpad && j > 0) wpad[b * Lt + j - 1] = v ? 0 : 1;
}
cudaError_t launch_expand_mask(const uint8_t* wmask, int B, int Lt, uint8_t* emask, uint8_t* epad, uint8_t* wpad, cudaStream_t s) {
    ProfScope _ps("expand_mask", s);
    if (B <= 0) return cudaSuccess;
    expand_mask_kernel<<<blocks_for((long long)B * (Lt + 1), 256), 256, 0, s>>>(wmask, B, Lt, emask, epad, wpad);
    LAUNCH_END();
}

__global__ void wpad_from_epad_kernel(const uint8_t* __restrict__ epad, int B, int Lt, uint8_t* __restrict__ wpad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * Lt) return;
    const int b = idx / Lt, j = idx % Lt;
    wpad[idx] = epad[b * (Lt + 1) + 1 + j];
}
cudaError_t launch_expand_mask_from_epad(const uint8_t* epad, int B, int Lt, uint8_t* wpad, cudaStream_t s) {
    ProfScope _ps("expand_mask_from_epad", s);
    if (B <= 0) return cudaSuccess;
    wpad_from_epad_kernel<<<blocks_for((long long)B * Lt, 256), 256, 0, s>>>(epad, B, Lt, wpad);
    LAUNCH_END();
}

// dst row block b <- src row block idx[b]   (rows_per x cols floats), plus an optional byte mask of rows_per entries
__global__ void gather_blocks_kernel(const float* __restrict__ src, float* __restrict__ dst, const int64_t* __restrict__ idx,
                                     int B, long long block_elems, const uint8_t* __restrict__ msrc,
                                     uint8_t* __restrict__ mdst, int mlen) {
    const int b = blockIdx.y;
    long long sb = idx[b];
    sb = sb < 0 ? 0 : (sb >= B ? B - 1 : sb);        // an out-of-range index must not read outside the batch (Engine.forward validates it in debug mode)
    const float4* s4 = reinterpret_cast<const float4*>(src + sb * block_elems);
    float4* d4 = reinterpret_cast<float4*>(dst + (long long)b * block_elems);
    const long long n4 = block_elems >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) d4[i] = s4[i];
    if (msrc && blockIdx.x == 0)
        for (int i = threadIdx.x; i < mlen; i += blockDim.x) mdst[(long long)b * mlen + i] = msrc[sb * mlen + i];
}
cudaError_t launch_gather_blocks(const float* src, float* dst, const int64_t* idx, int B, long long block_elems,
                                 const uint8_t* msrc, uint8_t* mdst, int mlen, cudaStream_t s) {
    ProfScope _ps("gather_blocks", s);
    if (B <= 0) return cudaSuccess;
    if (block_elems & 3) return cudaErrorInvalidValue;
    dim3 grid((unsigned)min((long long)8, (block_elems / 4 + 255) / 256), B);
    gather_blocks_kernel<<<grid, 256, 0, s>>>(src, dst, idx, B, block_elems, msrc, mdst, mlen);
    LAUNCH_END();
}

// strided row copy: dst[omap(r)] = src[imap(r)] for r < R, 256-wide rows
__global__ void copy_rows_kernel(const float* __restrict__ src, int lds, RowMap imap, float* __restrict__ dst, int ldd,
                                 RowMap omap, long long R) {
    const long long r = (long long)blockIdx.x * 4 + (threadIdx.x >> 6);
    const int c = (threadIdx.x & 63) * 4;
    if (r >= R) return;
    *reinterpret_cast<float4*>(dst + omap((int)r) * ldd + c) = *reinterpret_cast<const float4*>(src + imap((int)r) * lds + c);
}
cudaError_t launch_copy_rows(const float* src, int lds, RowMap imap, float* dst, int ldd, RowMap omap, long long R, cudaStream_t s) {
    ProfScope _ps("copy_rows", s);
    if (R <= 0) return cudaSuccess;
    copy_rows_kernel<<<blocks_for(R, 4), 256, 0, s>>>(src, lds, imap, dst, ldd, omap, R);
    LAUNCH_END();
}

// broadcast one 256-vector to R rows (masked_sent_token -> recon state)
__global__ void broadcast_row_kernel(const float* __restrict__ vec, float* __restrict__ dst, long long R) {
    const long long r = (long long)blockIdx.x * 4 + (threadIdx.x >> 6);
    const int c = (threadIdx.x & 63) * 4;
    if (r >= R) return;
    *reinterpret_cast<float4*>(dst + r * D + c) = *reinterpret_cast<const float4*>(vec + c);
}
cudaError_t launch_broadcast_row(const float* vec, float* dst, long long R, cudaStream_t s) {
    ProfScope _ps("broadcast_row", s);
    if (R <= 0) return cudaSuccess;
    broadcast_row_kernel<<<blocks_for(R, 4), 256, 0, s>>>(vec, dst, R);
    LAUNCH_END();
}

// F.normalize(x, dim=1) with eps 1e-12 (model/model.py:486): out1 [R,256] (optional) and out2 rows via map (optional)
__global__ void l2norm_rows_kernel(const float* __restrict__ x, long long R, float* __restrict__ out1, float* __restrict__ out2,
                                   int ld2, RowMap map2) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float4* xp = reinterpret_cast<const float4*>(x + r * D + lane * 8);
    float4 a = xp[0], b = xp[1];
    float ss = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv; b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv;
    if (out1) { float4* o = reinterpret_cast<float4*>(out1 + r * D + lane * 8); o[0] = a; o[1] = b; }
    if (out2) { float4* o = reinterpret_cast<float4*>(out2 + map2((int)r) * ld2 + lane * 8); o[0] = a; o[1] = b; }
}
cudaError_t launch_l2norm_rows(const float* x, long long R, float* out1, float* out2, int ld2, RowMap map2, cudaStream_t s) {
    ProfScope _ps("l2norm_rows", s);
    if (R <= 0) return cudaSuccess;
    l2norm_rows_kernel<<<blocks_for(R, 8), 256, 0, s>>>(x, R, out1, out2, ld2, map2);
    LAUNCH_END();
}

// saliency = <proj1(memory[b,i]), proj2(memory_global[b])> / sqrt(256)  (model/model.py:301)
__global__ void saliency_kernel(const float* __restrict__ p1, RowMap map1, const float* __restrict__ p2, int B, int Lv,
                                float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= (long long)B * Lv) return;
    const int b = (int)(r / Lv);
    const float4* x = reinterpret_cast<const float4*>(p1 + map1((int)r) * D + lane * 8);
    const float4* y = reinterpret_cast<const float4*>(p2 + (long long)b * D + lane * 8);
    const float4 a0 = x[0], a1 = x[1], b0 = y[0], b1 = y[1];
    float d = a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w + a1.x * b1.x + a1.y * b1.y + a1.z * b1.z + a1.w * b1.w;
    d = warp_sum(d);
    if (lane == 0) out[r] = d / 16.f;
}
cudaError_t launch_saliency(const float* p1, RowMap map1, const float* p2, int B, int Lv, float* out, cudaStream_t s) {
    ProfScope _ps("saliency", s);
    if (B <= 0) return cudaSuccess;
    saliency_kernel<<<blocks_for((long long)B * Lv, 8), 256, 0, s>>>(p1, map1, p2, B, Lv, out);
    LAUNCH_END();
}

__global__ void saliency_packed_kernel(const float* __restrict__ p1, const int* __restrict__ cu, const float* __restrict__ p2, int B,
                                       int Lv, float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= (long long)B * Lv) return;
    const int b = (int)(r / Lv), i = (int)(r - (long long)b * Lv);
    long long e0; int n;
    pair_rows(cu, 1, b, 0, e0, n);                       // encoder rows of pair b: global token, then n - 1 clips
    if (i >= n - 1) { if (lane == 0) out[r] = 0.f; return; }
    const float4* x = reinterpret_cast<const float4*>(p1 + (e0 + 1 + i) * D + lane * 8);
    const float4* y = reinterpret_cast<const float4*>(p2 + (long long)b * D + lane * 8);
    const float4 a0 = x[0], a1 = x[1], b0 = y[0], b1 = y[1];
    float d = a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w + a1.x * b1.x + a1.y * b1.y + a1.z * b1.z + a1.w * b1.w;
    d = warp_sum(d);
    if (lane == 0) out[r] = d / 16.f;
}
cudaError_t launch_saliency_packed(const float* p1, const int* cu, const float* p2, int B, int Lv, float* out, cudaStream_t s) {
    ProfScope _ps("saliency", s);
    if (B <= 0) return cudaSuccess;
    saliency_packed_kernel<<<blocks_for((long long)B * Lv, 8), 256, 0, s>>>(p1, cu, p2, B, Lv, out);
    LAUNCH_END();
}

// ---- packed (variable-length) clip layout: gather tables ------------------------------------------------------------
__global__ void pack_table_kernel(const int* __restrict__ cu, int Lv, int* __restrict__ t_pad, const int* __restrict__ pair_group,
                                  const int* __restrict__ group_start) {
    const int b = blockIdx.x, c0 = cu[b], n = cu[b + 1] - c0;
    const int src = pair_group ? group_start[pair_group[b]] : b;      // shared group video: read the group's first pair
    for (int i = threadIdx.x; i < n; i += blockDim.x) t_pad[c0 + i] = src * Lv + i;
}
cudaError_t launch_pack_table(const int* cu, int B, int Lv, int* t_pad, cudaStream_t s, const int* pair_group, const int* group_start) {
    ProfScope _ps("pack_tables", s);
    if (B <= 0) return cudaSuccess;
    pack_table_kernel<<<B, 128, 0, s>>>(cu, Lv, t_pad, pair_group, group_start);
    LAUNCH_END();
}
// per-video tables of a batch whose groups share their video: t_vin[vcu[g] + i] = first pair of g * Lv + i (video row -> row
// of the zero-padded input), t_p2v[cu[b] + i] = vcu[group(b)] + i (packed pair row -> video row)
__global__ void video_tables_kernel(const int* __restrict__ cu, const int* __restrict__ vcu, const int* __restrict__ pair_group,
                                    const int* __restrict__ group_start, int B, int G, int Lv, int* __restrict__ t_vin, int* __restrict__ t_p2v) {
    const int j = blockIdx.x;
    if (j < G) {
        const int v0 = vcu[j], n = vcu[j + 1] - v0, src = group_start[j] * Lv;
        for (int i = threadIdx.x; i < n; i += blockDim.x) t_vin[v0 + i] = src + i;
    } else {
        const int b = j - G, c0 = cu[b], n = cu[b + 1] - c0, v0 = vcu[pair_group[b]];
        for (int i = threadIdx.x; i < n; i += blockDim.x) t_p2v[c0 + i] = v0 + i;
    }
}
cudaError_t launch_video_tables(const int* cu, const int* vcu, const int* pair_group, const int* group_start, int B, int G, int Lv,
                                int* t_vin, int* t_p2v, cudaStream_t s) {
    ProfScope _ps("pack_tables", s);
    if (B <= 0) return cudaSuccess;
    video_tables_kernel<<<G + B, 128, 0, s>>>(cu, vcu, pair_group, group_start, B, G, Lv, t_vin, t_p2v);
    LAUNCH_END();
}

// 16-bit stored clip features -> the TMA-addressable operand of the first projection: rows gathered through `table` (video
// row -> row of the padded [B, Lv, Dv] fp16 input, whose 2 * Dv-byte pitch is not 16-byte aligned for the shipped Dv) are
// copied to a compact [R, ldo] fp16 buffer (ldo % 8 == 0, tail columns zeroed) and their LayerNorm statistics (mean, rstd
// over the Dv values, fp32, two passes) are written to rowstat.  One warp per row; the second pass re-reads the row from L1/L2.
__global__ void __launch_bounds__(256) repack_f16_rows_kernel(const uint16_t* __restrict__ x, const int* __restrict__ table, long long r0,
                                                               long long R, int Dv, uint16_t* __restrict__ out, int ldo, float* __restrict__ rowstat) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const long long src = table ? (long long)table[r0 + r] : r0 + r;
    const uint16_t* xr = x + src * Dv;
    uint16_t* orow = out + r * ldo;
    float s = 0.f;
    const bool pair_ok = (Dv & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0;
    const int n2 = Dv >> 1;
    if (pair_ok) {
        const uint32_t* x2 = reinterpret_cast<const uint32_t*>(xr);
        for (int c = lane; c < n2; c += 32) {
            const uint32_t u = __ldg(x2 + c);
            const __half2 h = *reinterpret_cast<const __half2*>(&u);
            const float2 f = __half22float2(h);
            s += f.x + f.y;
        }
    } else {
        for (int c = lane; c < Dv; c += 32) s += __half2float(__ushort_as_half(xr[c]));
    }
    s = warp_sum(s);
    const float mean = s / Dv;
    float q = 0.f;
    if (pair_ok) {
        const uint32_t* x2 = reinterpret_cast<const uint32_t*>(xr);
        uint32_t* o2 = reinterpret_cast<uint32_t*>(orow);
        for (int c = lane; c < n2; c += 32) {
            const uint32_t u = __ldg(x2 + c);
            o2[c] = u;
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
            const float a = f.x - mean, b = f.y - mean;
            q = fmaf(a, a, q); q = fmaf(b, b, q);
        }
    } else {
        for (int c = lane; c < Dv; c += 32) {
            const uint16_t u = xr[c];
            orow[c] = u;
            const float d = __half2float(__ushort_as_half(u)) - mean;
            q = fmaf(d, d, q);
        }
    }
    for (int c = Dv + lane; c < ldo; c += 32) orow[c] = 0;
    q = warp_sum(q);
    if (lane == 0) { rowstat[2 * r] = mean; rowstat[2 * r + 1] = rsqrtf(q / Dv + 1e-5f); }
}
cudaError_t launch_repack_f16_rows(const uint16_t* x, const int* table, long long r0, long long R, int Dv, uint16_t* out, int ldo,
                                   float* rowstat, cudaStream_t s) {
    ProfScope _ps("repack_f16_rows", s);
    if (R <= 0) return cudaSuccess;
    repack_f16_rows_kernel<<<blocks_for(R, 8), 256, 0, s>>>(x, table, r0, R, Dv, out, ldo, rowstat);
    LAUNCH_END();
}

__global__ void chunk_tables_kernel(const int* __restrict__ cu, int* __restrict__ t_c2e, int* __restrict__ t_g,
                                    const int* __restrict__ len_off, int* __restrict__ t_posV, int* __restrict__ t_posE) {
    const int b = blockIdx.x, c0 = cu[b] - cu[0], n = cu[b + 1] - cu[b];
    const int po = len_off ? len_off[n] : 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        t_c2e[c0 + i] = c0 + b + 1 + i;
        if (len_off) { t_posV[c0 + i] = po + i; t_posE[c0 + b + 1 + i] = po + i; }
    }
    if (threadIdx.x == 0) { t_g[b] = c0 + b; if (len_off) t_posE[c0 + b] = 0; }
}
cudaError_t launch_chunk_tables(const int* cu, int Bc, int* t_c2e, int* t_g, cudaStream_t s, const int* len_off, int* t_posV,
                                int* t_posE) {
    ProfScope _ps("pack_tables", s);
    if (Bc <= 0) return cudaSuccess;
    chunk_tables_kernel<<<Bc, 128, 0, s>>>(cu, t_c2e, t_g, len_off, t_posV, t_posE);
    LAUNCH_END();
}
// same arithmetic as pos_embed_kernel for a prefix mask of n valid clips: x_embed = (i + 1) / (n + 1e-6) * 2 pi
__global__ void __launch_bounds__(256) pos_table_kernel(const int* __restrict__ dl_len, const int* __restrict__ dl_off,
                                                        const float* __restrict__ gpos, float* __restrict__ PT) {
    const int c = threadIdx.x;
    if (blockIdx.x == 0) { PT[c] = gpos[c]; return; }
    const int n = dl_len[blockIdx.x - 1];
    float* out = PT + (long long)dl_off[blockIdx.x - 1] * D;
    const float last = (float)n + 1e-6f;
    const float dim_t = powf(10000.f, (float)(2 * (c / 2)) / 256.f);
    for (int i = 0; i < n; ++i) {
        const float xe = (float)(i + 1) / last * 6.283185307179586f;
        const float arg = xe / dim_t;
        out[(long long)i * D + c] = (c & 1) ? cosf(arg) : sinf(arg);
    }
}
cudaError_t launch_pos_table(const int* dl_len, const int* dl_off, int nd, const float* gpos, float* PT, cudaStream_t s) {
    ProfScope _ps("pos_table", s);
    pos_table_kernel<<<nd + 1, 256, 0, s>>>(dl_len, dl_off, gpos, PT);
    LAUNCH_END();
}
__global__ void zero_masked_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask, long long R, int width) {
    const long long r = (long long)blockIdx.x * 4 + (threadIdx.x >> 6);
    if (r >= R || mask[r]) return;
    for (int c = threadIdx.x & 63; c < width; c += 64) x[r * width + c] = 0.f;
}
cudaError_t launch_zero_masked_rows(float* x, const uint8_t* mask, long long R, int width, cudaStream_t s) {
    ProfScope _ps("zero_masked_rows", s);
    if (R <= 0) return cudaSuccess;
    zero_masked_rows_kernel<<<blocks_for(R, 4), 256, 0, s>>>(x, mask, R, width);
    LAUNCH_END();
}

// ---- DAB-DETR decoder element-wise pieces (model/transformer.py:43-59, 344-397) ------------------------------------
__device__ __forceinline__ float inv_sigmoid(float x) {          // transformer.py:36-40, eps = 1e-3
    x = fminf(fmaxf(x, 0.f), 1.f);
    return logf(fmaxf(x, 1e-3f) / fmaxf(1.f - x, 1e-3f));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// ref0[b,q,:] = sigmoid(query_embed[q,:])
__global__ void dec_init_ref_kernel(const float* __restrict__ qe, int B, int nq, float* __restrict__ ref) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nq * 2) return;
    ref[idx] = sigmoidf(qe[idx % (nq * 2)]);
}
cudaError_t launch_dec_init_ref(const float* qe, int B, int nq, float* ref, cudaStream_t s) {
    ProfScope _ps("dec_init_ref", s);
    dec_init_ref_kernel<<<blocks_for((long long)B * nq * 2, 256), 256, 0, s>>>(qe, B, nq, ref);
    LAUNCH_END();
}

// sine[r, :] = gen_sineembed_for_position(ref[r])  (256 dims: 128 for the center, 128 for the width)
// scaled[r, :] = sine * (pos_trans[r,:] or 1) * (sigmoid(anchor[r]) / ref_width[r])      (transformer.py:357-376)
__global__ void dec_sine_kernel(const float* __restrict__ ref, long long R, const float* __restrict__ pos_trans,
                                const float* __restrict__ anchor, float* __restrict__ sine, float* __restrict__ scaled) {
    const long long r = (long long)blockIdx.x;
    const int c = threadIdx.x;
    if (r >= R) return;
    const int part = c >> 7, j = c & 127;
    const float x = ref[r * 2 + part] * 6.283185307179586f;
    const float dim_t = powf(10000.f, (float)(2 * (j / 2)) / 128.f);
    const float arg = x / dim_t;
    const float v = (j & 1) ? cosf(arg) : sinf(arg);
    if (sine) sine[r * D + c] = v;
    if (scaled) {
        float sv = v;
        if (pos_trans) sv *= pos_trans[r * D + c];
        sv *= sigmoidf(anchor[r]) / ref[r * 2 + 1];
        scaled[r * D + c] = sv;
    }
}
cudaError_t launch_dec_sine(const float* ref, long long R, const float* pos_trans, const float* anchor, float* sine,
                            float* scaled, cudaStream_t s) {
    ProfScope _ps("dec_sine", s);
    if (R <= 0) return cudaSuccess;
    dec_sine_kernel<<<(unsigned)R, 256, 0, s>>>(ref, R, pos_trans, anchor, sine, scaled);
    LAUNCH_END();
}

// out[r,0:2] = sigmoid(delta[r,0:2] + inverse_sigmoid(ref[r,0:2]))   (ref update 387-397 and span head model.py:247-252)
__global__ void ref_update_kernel(const float* __restrict__ delta, int ldd, const float* __restrict__ ref, long long n,
                                  float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = i >> 1; const int c = (int)(i & 1);
    out[i] = sigmoidf(delta[r * ldd + c] + inv_sigmoid(ref[i]));
}
cudaError_t launch_ref_update(const float* delta, int ldd, const float* ref, long long rows, float* out, cudaStream_t s) {
    ProfScope _ps("ref_update", s);
    if (rows <= 0) return cudaSuccess;
    ref_update_kernel<<<blocks_for(rows * 2, 256), 256, 0, s>>>(delta, ldd, ref, rows * 2, out);
    LAUNCH_END();
}

// LayerNorm over 256-wide rows (decoder output norm, transformer.py:399-406)
__global__ void layernorm_rows_kernel(const float* __restrict__ x, long long R, const float* __restrict__ g,
                                      const float* __restrict__ bta, float* __restrict__ out) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float4* xp = reinterpret_cast<const float4*>(x + r * D + lane * 8);
    const float4 a = xp[0], b = xp[1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    const float mu = warp_sum(s) * (1.f / 256.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - mu; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) * (1.f / 256.f) + 1e-5f);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mu) * rs * g[lane * 8 + j] + bta[lane * 8 + j];
    float4* op = reinterpret_cast<float4*>(out + r * D + lane * 8);
    op[0] = make_float4(o[0], o[1], o[2], o[3]);
    op[1] = make_float4(o[4], o[5], o[6], o[7]);
}
cudaError_t launch_layernorm_rows(const float* x, long long R, const float* g, const float* b, float* out, cudaStream_t s) {
    ProfScope _ps("layernorm_rows", s);
    if (R <= 0) return cudaSuccess;
    layernorm_rows_kernel<<<blocks_for(R, 8), 256, 0, s>>>(x, R, g, b, out);
    LAUNCH_END();
}

// ---- qvhighlights grouping: valid clips per video group (model/model.py:191) ----------------------------------------
__global__ void group_len_kernel(const uint8_t* __restrict__ vmask, int Lv, const int* __restrict__ group_start, int G,
                                 int* __restrict__ group_len) {
    const int g = blockIdx.x;
    if (g >= G) return;
    const long long lo = (long long)group_start[g] * Lv, hi = (long long)group_start[g + 1] * Lv;
    int c = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) c += vmask[i] ? 1 : 0;
    c = (int)warp_sum((float)c);   // exact for counts < 2^24
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) group_len[g] = tot;
}
cudaError_t launch_group_len(const uint8_t* vmask, int Lv, const int* group_start, int G, int* group_len, cudaStream_t s) {
    ProfScope _ps("group_len", s);
    if (G <= 0) return cudaSuccess;
    group_len_kernel<<<G, 128, 0, s>>>(vmask, Lv, group_start, G, group_len);
    LAUNCH_END();
}

// ---- alignment scores: masked mean + L2 normalise (model/criterion.py:241-259) --------------------------------------
// out row b (ld = ldo) or, if transposed, column b of a [256, ldo] matrix.
__global__ void masked_mean_norm_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int B, int L,
                                        float* __restrict__ out, int ldo, int transposed) {
    const int b = blockIdx.x, c = threadIdx.x;      // 256 threads
    float s = 0.f; int cnt = 0;
    for (int i = 0; i < L; ++i) {
        if (mask[(long long)b * L + i]) { s += x[((long long)b * L + i) * D + c]; ++cnt; }
    }
    const float v = s / (float)cnt;
    __shared__ float red[8];
    float ss = warp_sum(v * v);
    if ((c & 31) == 0) red[c >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float o = v / fmaxf(sqrtf(tot), 1e-12f);
    if (transposed) out[(long long)c * ldo + b] = o; else out[(long long)b * ldo + c] = o;
}
cudaError_t launch_masked_mean_norm(const float* x, const uint8_t* mask, int B, int L, float* out, int ldo, int transposed, cudaStream_t s) {
    ProfScope _ps("masked_mean_norm", s);
    if (B <= 0) return cudaSuccess;
    masked_mean_norm_kernel<<<B, 256, 0, s>>>(x, mask, B, L, out, ldo, transposed);
    LAUNCH_END();
}

__global__ void fill_kernel(float* p, long long n, float v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
cudaError_t launch_fill(float* p, long long n, float v, cudaStream_t s) {
    ProfScope _ps("fill", s);
    if (n <= 0) return cudaSuccess;
    fill_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, n, v);
    LAUNCH_END();
}

// Pair/group tables live in pinned host memory; the device pulls them over PCIe with a kernel so that the transfer never
// queues on the copy engine behind a caller's bulk host->device prefetch.
__global__ void pull_ints_kernel(const int* __restrict__ host_src, int* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = host_src[i];
}
cudaError_t launch_pull_ints(const int* host_src, int* dst, long long n, cudaStream_t s) {
    ProfScope _ps("pull_tables", s);
    if (n <= 0) return cudaSuccess;
    pull_ints_kernel<<<blocks_for(n, 256), 256, 0, s>>>(host_src, dst, n);
    LAUNCH_END();
}

}  // namespace mesm
