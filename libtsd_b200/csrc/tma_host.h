// Host side of the tensor-map (TMA) kernels: cuTensorMapEncodeTiled reached through the runtime's driver entry point
// (no link against libcuda), and a helper for the 2-D float32 row maps of the tensor-core kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace tsdgpu {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tma_encode_fn()
{
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn) f;
  }();
  return fn;
}

// [rows][row_floats] float32 with a row pitch of pitch_bytes (multiple of 16), box {box_floats, box_rows}; out-of-bounds
// elements read as zero / are not written.  Returns false when the driver refuses the description.
inline bool tma_map_rows(CUtensorMap *map, const void *base, unsigned long long row_floats, unsigned long long rows,
                         unsigned long long pitch_bytes, unsigned box_floats, unsigned box_rows, bool swizzle128)
{
  EncodeTiledFn fn = tma_encode_fn();
  if(!fn) return false;
  const cuuint64_t dims[2] = {row_floats, rows};
  const cuuint64_t strides[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_floats, box_rows}, estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace tsdgpu
