// Host-side tensor-map encoding (cuTensorMapEncodeTiled resolved through the runtime, no -lcuda).
#pragma once
#include <cuda.h>

namespace das {
// rank <= 3; dims / box in elements (fastest first); strides in bytes for dims 1.. (rank - 1 entries)
int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle);
}  // namespace das
