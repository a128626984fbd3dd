// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the driver-API types the kernel sources name.
#pragma once
#include <cstdint>

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
enum : int { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };

// the 128-byte opaque descriptor; the emulated encoder keeps its own record inside (emu::TensorMapRec, emu_async.h)
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };

enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 = 11 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B = 1, CU_TENSOR_MAP_SWIZZLE_64B = 2, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B = 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
