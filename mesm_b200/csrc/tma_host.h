// Host-side tensor-map (TMA descriptor) construction shared by the TMA-fed kernels (linear_tma.cu, ffn_tc.cu).
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint): the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace mesm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tma_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major tensor of 16-bit elements: `inner` elements per row (the contiguous dimension), `outer` rows, `ld_bytes` between rows
// (a multiple of 16; base 16-byte aligned).  Box = box_inner x box_outer elements; out-of-bounds box elements are zero-filled.
static inline bool tma_map_2d_16bit(CUtensorMap* tm, const void* base, unsigned long long inner, unsigned long long outer, unsigned long long ld_bytes,
                                    unsigned box_inner, unsigned box_outer, CUtensorMapSwizzle swz, bool fp16 = false) {
    EncodeTiledFn enc = tma_encode_fn();
    if (!enc || (reinterpret_cast<unsigned long long>(base) & 15) || (ld_bytes & 15)) return false;
    const cuuint64_t gdim[2] = {inner, outer};
    const cuuint64_t gstr[1] = {ld_bytes};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mesm
