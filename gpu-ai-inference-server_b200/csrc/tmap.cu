// tmap.cu — cuTensorMapEncodeTiled without linking libcuda: the driver entry point is resolved through the runtime.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "kernels.h"

namespace b200 {
namespace kernels {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn GetEncodeTiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}
}  // namespace

static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "TensorMap must mirror CUtensorMap");

int MakeTensorMap(TensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128) {
    EncodeTiledFn fn = GetEncodeTiled();
    if (!fn) return -1;
    if (rank < 1 || rank > 5) return -2;
    cuuint64_t d[5];
    cuuint64_t s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                             : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    // 128-byte promotion: activation boxes read a 128-byte channel slice of a wider pixel; with L2_256B the
                    // TMA unit pulled the neighbouring 128 bytes as well (ncu: 205 MB instead of 103 MB of DRAM reads for a
                    // 64..128-channel layer of block 1)
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace kernels
}  // namespace b200
