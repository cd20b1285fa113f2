// Shared device/host helpers for libdas_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "das_b200.h"

namespace das {

// ---- host side bookkeeping -----------------------------------------------------------------
extern int g_last_cuda_error;
extern unsigned long long g_launch_count;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return DAS_ERR_CUDA;
}

// every kernel launch goes through this so that das_launch_count() is an honest count
#define DAS_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        ++::das::g_launch_count;                                           \
    } while (0)

#define DAS_CHECK_LAUNCH()                                                 \
    do {                                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                           \
        if (e__ != cudaSuccess) return ::das::cuda_fail(cudaGetLastError()); \
    } while (0)

#define DAS_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return ::das::cuda_fail(e__);              \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// state layout shared by accumulate / finalize (offsets in bytes, all 256-byte aligned)
struct McLayout {
    size_t sum_p, sum_ent, votes, partials, total;
    int blocks_per_image;  // K2 blocks per image (256 threads)
    int blocks_fused;      // fused K1+K2 blocks per image (128 threads); partials are sized for this
};
McLayout mc_layout(const das_mc_desc& d);
int mc_validate(const das_mc_desc* d);
constexpr int kFinalizeThreads = 256;

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kEps = 1e-12f;  // mc_dropout.py:48 / ceal.py:118

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// L2 cache policies: logits are touched once (evict-first); running state should survive in
// the 126 MB L2 between the launches of consecutive passes (evict-last).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// streaming 128-bit / 32-bit loads: read-only path, no L1 allocation, L2 policy hint
__device__ __forceinline__ float4 ldg_stream(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ldg_hint(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_hint(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void stg_hint(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// order-preserving float -> uint32 (handles sign); -0 is folded into +0 first
__device__ __forceinline__ uint32_t float_orderable(float f) {
    f += 0.0f;
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_orderable(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

#endif  // __CUDACC__

}  // namespace das
