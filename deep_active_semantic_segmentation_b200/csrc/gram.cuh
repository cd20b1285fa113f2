// Shared between gram.cu (tensor-core distance filter) and kcenter.cu (exact float64 greedy update).
#pragma once
#include <stddef.h>

namespace das {

// |d2~ - d2| <= kFilterDelta * (|a|^2 + |b|^2): twice the proven bf16 + fp32-accumulate bound (gram.cu header)
constexpr double kFilterDelta = 1.0 / 128.0;

// device blob built by das_kcenter_filter_build (offsets in bytes)
struct KcFilterLayout {
    size_t fb;     // bf16 [N, Dp] features, zero padded to Dp = ceil(D / 64) * 64
    size_t nrm64;  // f64 [N] exact squared norms of the float32 rows
    size_t nrm32;  // f32 [N] the same, rounded (epilogue operand)
    size_t stats;  // u64 [2]: rows re-evaluated exactly, rows screened
    size_t dt;     // f32 [N, ld]: dt[c * ld + (i - row_begin)] ~ |f_i - f_c|^2, i in the rank's row shard
    size_t total;
    int Dp, ld;
};
KcFilterLayout kc_filter_layout(int N, int D, int rows);

}  // namespace das
