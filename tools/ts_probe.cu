// Developer probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, bf16, M = 128).  Checks the assumed A layout:
// lane = row, 32-bit column c holds K elements (2c, 2c+1) as packed bf16x2 (low half = even k).  D[128 x 64] = A[128 x 64] . B[64 x 64]^T.
#include "../mesm_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace mesm::tc;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128, 1) ts_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    // B: 64 rows (N) x 64 K as two K blocks of 32 (SWIZZLE_64B tiles of 64 rows x 64 B = 4 KB each)
    uint32_t* tptr = reinterpret_cast<uint32_t*>(smem + 8192);
    const uint32_t bar = sbase + 8192 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        const __nv_bfloat16 v = __float2bfloat16_rn(B[n * 64 + k]);
        *reinterpret_cast<__nv_bfloat16*>(smem + (k / 32) * 4096 + sw64(n, k % 32)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    // A row `r` = threadIdx.x: 64 bf16 -> 32 packed columns at TMEM columns [64, 96) (D uses columns [0, 64))
    {
        const int r = threadIdx.x;
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const __nv_bfloat162 p = __floats2bfloat162_rn(A[r * 64 + 2 * c], A[r * 64 + 2 * c + 1]);
            v[c] = __uint_as_float(*reinterpret_cast<const uint32_t*>(&p));
        }
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64, v);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(64);
#pragma unroll
        for (int k = 0; k < 4; ++k) {            // K = 16 per instruction: 8 packed A columns, 32 bytes of the B rows
            const uint64_t bdesc = make_desc(sbase + (k / 2) * 4096 + (k % 2) * 32);
            umma_ts(tmem, tmem + 64 + k * 8, bdesc, k > 0 ? 1u : 0u, idesc);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0, 1);
    tc_fence_after();
    for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        for (int j = 0; j < 32; ++j) D[threadIdx.x * 64 + c * 32 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64), R(128 * 64);
    for (size_t i = 0; i < A.size(); ++i) A[i] = (float)((int)((i * 2654435761u) % 17) - 8) / 8.f;      // exactly representable in bf16
    for (size_t i = 0; i < B.size(); ++i) B[i] = (float)((int)((i * 40503u + 7) % 13) - 6) / 4.f;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)A[m * 64 + k] * B[n * 64 + k]; R[m * 64 + n] = (float)s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    ts_probe_kernel<<<1, 128, 16384>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < D.size(); ++i) { const double d = fabs((double)D[i] - R[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    printf("TS-MMA max |err| = %g, mismatches = %d / %zu   D[0..3] = %g %g %g %g  ref %g %g %g %g\n", maxerr, bad, D.size(), D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    return 0;
}
