// PTX wrappers shared by the tcgen05 kernels (linear_tc.cu, attn_tc.cu): mbarrier, bulk async copy, UMMA descriptors for
// K-major SWIZZLE_64B operand tiles (rows of 32 bf16), tcgen05.mma / commit / ld / st, bf16 hi/lo splitting.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace mesm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Watchdog: a wait that lasts ~2 s records {tag, block, thread, barrier, parity} once and gives up,
// so that a protocol bug surfaces as a diagnosable wrong answer instead of a hung GPU (host: mesm_debug_watchdog()).
#ifndef MBAR_SUSPEND_NS
#define MBAR_SUSPEND_NS 4000u      // try_wait suspend-time hint: a waiting thread sleeps in hardware instead of spinning on issue slots
#endif
__device__ unsigned long long g_tc_watchdog[64];      // [0] = number of records; record i at [1 + 3i]: tag, block, thread|parity
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
        if (!done && (spins & 255u) == 255u && clock64() - t_start > 4000000000ll) {      // ~2 s
            const unsigned long long slot = atomicAdd(&g_tc_watchdog[0], 1ull);
            if (slot < 21) {
                g_tc_watchdog[1 + 3 * slot] = (unsigned long long)tag;
                g_tc_watchdog[2 + 3 * slot] = ((unsigned long long)blockIdx.y << 32) | blockIdx.x;
                g_tc_watchdog[3 + 3 * slot] = ((unsigned long long)threadIdx.x << 32) | ((unsigned long long)(bar & 0xffffff) << 8) | parity;
            }
            return;
        }
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14),
// LBO (unused for swizzled K-major) [16,30), SBO = 512 B (8 rows x 64 B) [32,46), version 1 [46,48), layout type
// SWIZZLE_64B = 4 in [61,64).  Rows are 64 bytes (32 bf16 of K); the 16-byte chunk index is XORed with (row >> 1) & 3.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128 (256 for a CTA pair), N given.
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = 128) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pair (cluster of 2, cta_group::2): one thread of the leader CTA issues M = 256 MMAs that read A rows and half of
// the B rows from EACH CTA's shared memory (same offsets in both) and write each CTA's 128 accumulator lanes in its TMEM.
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_saddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_saddr) : "memory");
}
// wait on a LOCAL barrier whose arrivals come from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag = 0) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
        if (!done && (spins & 255u) == 255u && clock64() - t_start > 4000000000ll) {
            const unsigned long long slot = atomicAdd(&g_tc_watchdog[0], 1ull);
            if (slot < 21) {
                g_tc_watchdog[1 + 3 * slot] = (unsigned long long)tag;
                g_tc_watchdog[2 + 3 * slot] = ((unsigned long long)blockIdx.y << 32) | blockIdx.x;
                g_tc_watchdog[3 + 3 * slot] = ((unsigned long long)threadIdx.x << 32) | ((unsigned long long)(bar & 0xffffff) << 8) | parity;
            }
            return;
        }
    }
}

// Latency-critical single-thread waits (producer / MMA issuer): plain polling, no suspend-time hint.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity, int tag = 0, bool cluster = false) {
    uint32_t done = 0;
    const long long t_start = clock64();
    for (uint32_t spins = 0; !done; ++spins) {
        if (cluster)
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && (spins & 1023u) == 1023u && clock64() - t_start > 4000000000ll) {
            const unsigned long long slot = atomicAdd(&g_tc_watchdog[0], 1ull);
            if (slot < 21) {
                g_tc_watchdog[1 + 3 * slot] = (unsigned long long)tag;
                g_tc_watchdog[2 + 3 * slot] = ((unsigned long long)blockIdx.y << 32) | blockIdx.x;
                g_tc_watchdog[3 + 3 * slot] = ((unsigned long long)threadIdx.x << 32) | ((unsigned long long)(bar & 0xffffff) << 8) | parity;
            }
            return;
        }
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    __syncwarp();                       // .sync.aligned: the whole warp must be converged
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// byte offset of element (row, k) inside a K-major SWIZZLE_64B tile whose rows hold BK = 32 bf16
__device__ __forceinline__ int sw64(int row, int k) { return row * 64 + ((((k >> 3) ^ ((row >> 1) & 3)) << 4) | ((k & 7) << 1)); }

// Two floats -> packed bf16 hi and lo words (element a in the low half).  One F2FP.PACK_AB per word: the scalar
// __float2bfloat16_rn compiles to F2F.BF16.F32, which issues on the quarter-rate conversion unit (16 lanes/clk/SM) - 64 of them
// per K block and thread made the operand converters of linear_tc XU-bound.  Bit-identical to split_bf16 per element.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}


}  // namespace tc
}  // namespace mesm
