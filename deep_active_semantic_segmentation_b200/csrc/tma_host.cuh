// Host-side tensor-map encoding (cuTensorMapEncodeTiled resolved through the runtime, no -lcuda).  Defined in handle.cu.
#pragma once
#include <cuda.h>

struct das_handle;

namespace das {
// rank <= 3; dims / box in elements (fastest first); strides in bytes for dims 1.. (rank - 1 entries)
int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle);
// the same through the handle's descriptor cache (keyed by every argument); nullptr + *rc on failure
const CUtensorMap* cached_tensor_map(das_handle* h, CUtensorMapDataType dtype, int rank, const void* base,
                                     const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                     CUtensorMapSwizzle swizzle, int* rc);
}  // namespace das
