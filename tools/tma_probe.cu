// Developer probe: how fast can one SM stream 16 KB weight slots from an L2-resident 2 MB image through TMA, as ffn_pair_kernel does?
//   mode 0: every CTA completes its loads on its OWN mbarrier (cta_group::1 form)
//   mode 1: CTA pairs; both CTAs complete on the LEADER's mbarrier (cta_group::2 form, as in ffn_pair_kernel); the leader releases the slot
//           to both CTAs with remote arrives
//   mode 2: like 0 but plain cp.async.bulk (1-D, 16 KB contiguous)
// Usage: tma_probe [mode] [nslot] [slots_per_cta] [nbytes_image_MB]
#include "../mesm_b200/csrc/tc_common.cuh"
#include "../mesm_b200/csrc/tma_host.h"
#include <cstdio>
#include <cstdlib>
using namespace mesm; using namespace mesm::tc;
namespace mesm { thread_local LaunchStats g_stats; }
constexpr int SLOT = 16384;
__device__ long long g_cycles[512];

__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* img, int mode, int nslot, int nloads, int img_slots) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = sbase + nslot * SLOT;           // full[nslot] then empty[nslot]
    const uint32_t bar_full = bars, bar_empty = bars + 64;
    const uint32_t rank = blockIdx.x & 1u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nslot; ++s) { mbar_init(bar_full + 8 * s, mode == 1 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {                               // producer
        const uint32_t full_leader = map_to_cta(bar_full, 0);
        for (int g = 0; g < nloads; ++g) {
            const int s = g % nslot; const uint32_t ph = (g / nslot) & 1;
            mbar_wait_spin(bar_empty + 8 * s, ph ^ 1, 100);
            const int src_slot = (int)(((unsigned)g * 7u + (blockIdx.x >> 1) * 13u + rank * 4u) % (unsigned)img_slots);
            const uint32_t dst = sbase + s * SLOT;
            if (mode == 1) {
                asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(full_leader + 8 * s), "r"(SLOT) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                             "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(src_slot * 128), "r"(full_leader + 8 * s) : "memory");
            } else if (mode == 0) {
                mbar_arrive_expect_tx(bar_full + 8 * s, SLOT);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                             "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(src_slot * 128), "r"(bar_full + 8 * s) : "memory");
            } else {
                mbar_arrive_expect_tx(bar_full + 8 * s, SLOT);
                bulk_copy_g2s(dst, img + (size_t)src_slot * SLOT, SLOT, bar_full + 8 * s);
            }
        }
    } else if (threadIdx.x == 32) {                       // consumer: frees a slot as soon as it is full
        if (mode == 1) {
            if (rank == 0) {
                const uint32_t e0 = map_to_cta(bar_empty, 0), e1 = map_to_cta(bar_empty, 1);
                for (int g = 0; g < nloads; ++g) {
                    const int s = g % nslot; const uint32_t ph = (g / nslot) & 1;
                    mbar_wait_spin(bar_full + 8 * s, ph, 200, true);
                    mbar_arrive_remote(e0 + 8 * s); mbar_arrive_remote(e1 + 8 * s);
                }
            }
        } else {
            for (int g = 0; g < nloads; ++g) {
                const int s = g % nslot; const uint32_t ph = (g / nslot) & 1;
                mbar_wait_spin(bar_full + 8 * s, ph, 200);
                mbar_arrive(bar_empty + 8 * s);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0 && blockIdx.x < 512) g_cycles[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0, nslot = argc > 2 ? atoi(argv[2]) : 5, nloads = argc > 3 ? atoi(argv[3]) : 2000, mb = argc > 4 ? atoi(argv[4]) : 2;
    const size_t bytes = (size_t)mb << 20;
    uint8_t* img; cudaMalloc(&img, bytes); cudaMemset(img, 1, bytes);
    CUtensorMap tm;
    EncodeTiledFn enc = tma_encode_fn();
    const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 128)}; const cuuint64_t gstr[1] = {128}; const cuuint32_t box[2] = {128, 128}; const cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); return 1; }
    const int smem = nslot * SLOT + 256 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148, 1, 1); cfg.blockDim = dim3(64, 1, 1); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int it = 0; it < 3; ++it) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, tm, (const uint8_t*)img, mode, nslot, nloads, (int)(bytes / SLOT));
        if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    long long h[148]; cudaMemcpyFromSymbol(h, g_cycles, sizeof(h));
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("mode %d nslot %d image %d MB: %.0f cycles per 16 KB slot per CTA = %.1f B/clk/SM, chip %.0f B/clk\n", mode, nslot, mb, avg / nloads, SLOT / (avg / nloads), 148.0 * SLOT / (avg / nloads));
    return 0;
}
