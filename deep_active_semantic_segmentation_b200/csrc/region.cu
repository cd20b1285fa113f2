// Region scoring: labelled-rect suppression, sliding RxR box sums, pool min/max normalisation and the
// image-local greedy NMS sequences (reference active_selection/mc_dropout.py:82-158).
#include <math.h>

#include "das_common.cuh"

namespace das {

// ---------------------------------------------------------------------------------------------
// suppress_labeled_entropy (mc_dropout.py:110-121): one block per rect record (image,r,c,h,w)
// ---------------------------------------------------------------------------------------------
__global__ void suppress_rects_kernel(float* maps, int B, int H, int W, const int32_t* rects, int n) {
    const int32_t* q = rects + (size_t)blockIdx.x * 5;
    const int img = q[0];
    if (img < 0 || img >= B) return;
    // python slice semantics for [r:r+h, c:c+w] with non-negative r, c
    const int r0 = min(max(q[1], 0), H), c0 = min(max(q[2], 0), W);
    const int r1 = min(max(q[1] + q[3], r0), H), c1 = min(max(q[2] + q[4], c0), W);
    const int w = c1 - c0, cells = (r1 - r0) * w;
    float* m = maps + (size_t)img * H * W;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) m[(size_t)(r0 + i / w) * W + c0 + i % w] = 0.f;
}

// the same with the records passed BY VALUE in the kernel parameters (host rects: no host-to-device copy, which for a
// pageable source would be a synchronous one and stall the batch pipeline of create_region_maps)
constexpr int kRectsPerLaunch = 128;
struct RectBlock {
    int32_t v[kRectsPerLaunch * 5];
};
__global__ void suppress_rects_param_kernel(float* maps, int B, int H, int W, const __grid_constant__ RectBlock rb) {
    const int32_t* q = rb.v + blockIdx.x * 5;
    const int img = q[0];
    if (img < 0 || img >= B) return;
    const int r0 = min(max(q[1], 0), H), c0 = min(max(q[2], 0), W);
    const int r1 = min(max(q[1] + q[3], r0), H), c1 = min(max(q[2] + q[4], c0), W);
    const int w = c1 - c0, cells = (r1 - r0) * w;
    float* m = maps + (size_t)img * H * W;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) m[(size_t)(r0 + i / w) * W + c0 + i % w] = 0.f;
}

__global__ void add_maps_kernel(float* a, const float* b, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a[i] = a[i] + b[i];
}

// ---------------------------------------------------------------------------------------------
// box sum, step 1: vertical sliding sums in fp64.  V[b,r,c] = sum_{i<R} M[b,r+i,c], r < H2.
// A thread owns one column and kSeg consecutive output rows: R loads to start, then add/subtract.
// Threads of a warp sit on consecutive columns, so every load is coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int kSeg = 32;
__global__ void box_vertical_kernel(const float* __restrict__ maps, int H, int W, int R, int H2, double* __restrict__ V) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= W) return;
    const int r0 = blockIdx.y * kSeg;
    const int b = blockIdx.z;
    const float* m = maps + (size_t)b * H * W + c;
    double* v = V + (size_t)b * H2 * W + c;
    double s = 0.0;
    for (int i = 0; i < R; ++i) s += (double)m[(size_t)(r0 + i) * W];
    const int r_end = min(r0 + kSeg, H2);
    v[(size_t)r0 * W] = s;
    for (int r = r0 + 1; r < r_end; ++r) {
        s += (double)m[(size_t)(r + R - 1) * W] - (double)m[(size_t)(r - 1) * W];
        v[(size_t)r * W] = s;
    }
}

// atomic float min/max via the integer trick (initial values +inf / -inf)
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f)
        atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ---------------------------------------------------------------------------------------------
// box sum, step 2: one block per output row.  The row of V goes to shared memory, is prefix-summed
// in fp64 (warp shuffles over 32-wide rows + a scan of the row totals), and
// out[c] = P[c+R-1] - P[c-1] is rounded once to float32.  Block min/max -> atomics on minmax[2].
// ---------------------------------------------------------------------------------------------
constexpr int kBoxThreads = 256;
__global__ void __launch_bounds__(kBoxThreads) box_horizontal_kernel(const double* __restrict__ V, int W, int R, int H2,
                                                                     int W2, float* __restrict__ out, float* minmax) {
    extern __shared__ double sm[];  // W prefix sums + ceil(W/32) row totals
    const int r = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nrows = (W + 31) / 32;
    double* P = sm;
    double* tot = sm + W;
    const double* v = V + ((size_t)b * H2 + r) * W;
    for (int row = wid; row < nrows; row += kBoxThreads / 32) {
        const int i = row * 32 + lane;
        double x = i < W ? v[i] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (i < W) P[i] = x;
        if (lane == 31) tot[row] = x;
    }
    __syncthreads();
    if (wid == 0) {  // exclusive scan of the row totals, 32 at a time with a carry
        double carry = 0.0;
        for (int base = 0; base < nrows; base += 32) {
            const int i = base + lane;
            const double t = i < nrows ? tot[i] : 0.0;
            double x = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (i < nrows) tot[i] = carry + x - t;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    __syncthreads();
    float mn = INFINITY, mx = -INFINITY;
    float* o = out + ((size_t)b * H2 + r) * W2;
    for (int c = tid; c < W2; c += kBoxThreads) {
        const int hi = c + R - 1;
        const double a = P[hi] + tot[hi >> 5];
        const double l = c > 0 ? P[c - 1] + tot[(c - 1) >> 5] : 0.0;
        const float s = (float)(a - l);
        o[c] = s;
        mn = fminf(mn, s);
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, k));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
    }
    __shared__ float smn[kBoxThreads / 32], smx[kBoxThreads / 32];
    if (lane == 0) smn[wid] = mn, smx[wid] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kBoxThreads / 32; ++w) mn = fminf(mn, smn[w]), mx = fmaxf(mx, smx[w]);
        if (mn <= mx) {
            atomic_min_float(minmax, mn);
            atomic_max_float(minmax + 1, mx);
        }
    }
}

__global__ void minmax_init_kernel(float* minmax) {
    minmax[0] = INFINITY;
    minmax[1] = -INFINITY;
}

// x.add_(-min).mul_(1.0 / (max - min)) in float32 (mc_dropout.py:154-155)
__global__ void minmax_normalise_kernel(float* x, size_t n, const float* minmax) {
    const float neg_min = -minmax[0];
    const float inv = __fdiv_rn(1.0f, __fadd_rn(minmax[1], -minmax[0]));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = __fmul_rn(__fadd_rn(x[i], neg_min), inv);
}

// ---------------------------------------------------------------------------------------------
// Image-local greedy NMS (mc_dropout.py:87-106 restricted to one image): one CTA per image.
// argmax key = (orderable(score) << 32) | ~flat_index  -> max key == first flat index among the
// maximal scores, exactly torch's argmax tie rule.
//
// The reference re-scans the WHOLE pool for every pick (mc_dropout.py:91).  Here the map is covered by 32 x 32 tiles
// whose arg-max keys live in shared memory: a pick is the arg-max over the tile keys; zeroing its [r-R, r+R) x [c-R, c+R)
// window only invalidates the tiles the window touches, and of those the ones it covers completely need no re-scan
// (all zero: key of 0.0 at the tile's first element).  Per pick: 64 K cells zeroed + ~32 edge tiles re-read instead of
// the 345 K-cell map (config 3: 385 x 897, R = 128) - 3.7x less traffic per image, same picks bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 512;
constexpr int kNmsTile = 32;

__device__ __forceinline__ unsigned long long nms_key(float v, int flat) {
    return ((unsigned long long)float_orderable(v) << 32) | (uint32_t)(~(uint32_t)flat);
}
// arg-max key of tile (tr, tc) by one warp: lane = column, coalesced 128-byte row segments
__device__ __forceinline__ unsigned long long nms_tile_scan(const float* m, int H2, int W2, int tr, int tc, int lane) {
    const int c = tc * kNmsTile + lane;
    const int r_end = min((tr + 1) * kNmsTile, H2);
    unsigned long long best = 0ull;
    if (c < W2)
        for (int r = tr * kNmsTile; r < r_end; ++r) {
            const int flat = r * W2 + c;
            const unsigned long long key = nms_key(m[flat], flat);
            best = key > best ? key : best;
        }
    return warp_max_u64(best);
}

__global__ void __launch_bounds__(kNmsThreads) nms_sequences_kernel(float* score_maps, int H2, int W2, int R, int kmax,
                                                                    float stop, float* cand_score, int32_t* cand_rc,
                                                                    int32_t* cand_count, long long image_offset,
                                                                    int64_t* cand_flat) {
    extern __shared__ unsigned long long tile_key[];  // [TH * TW]
    __shared__ unsigned long long wbest[kNmsThreads / 32];
    __shared__ unsigned long long best_s;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = kNmsThreads / 32;
    float* m = score_maps + (size_t)img * H2 * W2;
    const int TH = (H2 + kNmsTile - 1) / kNmsTile, TW = (W2 + kNmsTile - 1) / kNmsTile, nt = TH * TW;
    for (int t = wid; t < nt; t += NW) {
        const unsigned long long k = nms_tile_scan(m, H2, W2, t / TW, t % TW, lane);
        if (lane == 0) tile_key[t] = k;
    }
    __syncthreads();
    int picks = 0;
    while (picks < kmax) {
        unsigned long long best = 0ull;
        for (int t = tid; t < nt; t += kNmsThreads) best = tile_key[t] > best ? tile_key[t] : best;
        best = warp_max_u64(best);
        if (lane == 0) wbest[wid] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NW; ++w) best = wbest[w] > best ? wbest[w] : best;
            best_s = best;
        }
        __syncthreads();
        best = best_s;
        const float s = float_from_orderable((uint32_t)(best >> 32));
        const int flat = (int)(~(uint32_t)best);
        if (picks > 0 && s < stop) break;  // the pool loop checks the max AFTER a pick (mc_dropout.py:105)
        const int r = flat / W2, c = flat % W2;
        if (tid == 0) {
            cand_score[(size_t)img * kmax + picks] = s;
            cand_rc[((size_t)img * kmax + picks) * 2 + 0] = r;
            cand_rc[((size_t)img * kmax + picks) * 2 + 1] = c;
            if (cand_flat != nullptr)  // flat index in the un-sharded pool: the reference's arg-max tie rule
                cand_flat[(size_t)img * kmax + picks] = ((image_offset + img) * H2 + r) * (long long)W2 + c;
        }
        ++picks;
        const int r0 = max(0, r - R), r1 = min(H2, r + R), c0 = max(0, c - R), c1 = min(W2, c + R);
        const int w = c1 - c0, cells = (r1 - r0) * w;
        for (int i = tid; i < cells; i += kNmsThreads) m[(size_t)(r0 + i / w) * W2 + c0 + i % w] = 0.f;
        __syncthreads();  // the zeroed window is visible to the re-scans below
        // refresh the keys of the tiles the window touches
        const int tr_lo = r0 / kNmsTile, tr_hi = (r1 - 1) / kNmsTile, tc_lo = c0 / kNmsTile, tc_hi = (c1 - 1) / kNmsTile;
        const int tcols = tc_hi - tc_lo + 1, touched = (tr_hi - tr_lo + 1) * tcols;
        for (int k = wid; k < touched; k += NW) {
            const int tr = tr_lo + k / tcols, tc = tc_lo + k % tcols;
            const int tr0 = tr * kNmsTile, tc0 = tc * kNmsTile;
            const bool inside = tr0 >= r0 && min(tr0 + kNmsTile, H2) <= r1 && tc0 >= c0 && min(tc0 + kNmsTile, W2) <= c1;
            // completely inside the window: every cell is +0.0, the first flat index is the tile's top-left cell
            const unsigned long long key = inside ? nms_key(0.f, tr0 * W2 + tc0) : nms_tile_scan(m, H2, W2, tr, tc, lane);
            if (lane == 0) tile_key[tr * TW + tc] = key;
        }
        __syncthreads();
    }
    if (tid == 0) cand_count[img] = picks;
    // unused slots rank below every real candidate (das_topk over the flattened table)
    for (int j = picks + tid; j < kmax; j += kNmsThreads) {
        cand_score[(size_t)img * kmax + j] = -INFINITY;
        if (cand_flat != nullptr) cand_flat[(size_t)img * kmax + j] = -1;
    }
}

}  // namespace das

using namespace das;

extern "C" {

int das_suppress_rects(das_handle* h, float* maps, int B, int H, int W, const int32_t* rects, int n, void* stream) {
    DAS_ENTER(h);
    if (maps == nullptr || B <= 0 || H <= 0 || W <= 0 || n < 0 || (n > 0 && rects == nullptr)) return DAS_ERR_INVALID_ARG;
    if (n == 0) return DAS_OK;
    DAS_LAUNCH(suppress_rects_kernel, n, 256, 0, (cudaStream_t)stream, maps, B, H, W, rects, n);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_suppress_rects_host(das_handle* h, float* maps, int B, int H, int W, const int32_t* host_rects, int n, void* stream) {
    DAS_ENTER(h);
    if (maps == nullptr || B <= 0 || H <= 0 || W <= 0 || n < 0 || (n > 0 && host_rects == nullptr)) return DAS_ERR_INVALID_ARG;
    for (int i0 = 0; i0 < n; i0 += kRectsPerLaunch) {
        const int m = n - i0 < kRectsPerLaunch ? n - i0 : kRectsPerLaunch;
        RectBlock rb;
        for (int j = 0; j < m * 5; ++j) rb.v[j] = host_rects[(size_t)i0 * 5 + j];
        DAS_LAUNCH(suppress_rects_param_kernel, m, 256, 0, (cudaStream_t)stream, maps, B, H, W, rb);
        DAS_CHECK_LAUNCH();
    }
    return DAS_OK;
}

int das_add_maps(das_handle* h, float* a, const float* b, size_t n, void* stream) {
    DAS_ENTER(h);
    if (a == nullptr || b == nullptr) return DAS_ERR_INVALID_ARG;
    if (n == 0) return DAS_OK;
    const int grid = (int)((n + 255) / 256 < (size_t)(h->num_sms * 8) ? (n + 255) / 256 : (size_t)(h->num_sms * 8));
    DAS_LAUNCH(add_maps_kernel, grid, 256, 0, (cudaStream_t)stream, a, b, n);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_box_sum_workspace_bytes(int B, int H, int W, int R, size_t* bytes) {
    if (bytes == nullptr || B <= 0 || H <= 0 || W <= 0 || R <= 0 || R > H || R > W) return DAS_ERR_INVALID_ARG;
    *bytes = align_up((size_t)B * (H - R + 1) * W * sizeof(double), 256);
    return DAS_OK;
}

int das_minmax_init(das_handle* h, float* minmax, void* stream) {
    DAS_ENTER(h);
    if (minmax == nullptr) return DAS_ERR_INVALID_ARG;
    DAS_LAUNCH(minmax_init_kernel, 1, 1, 0, (cudaStream_t)stream, minmax);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_box_sum(das_handle* h, const float* maps, int B, int H, int W, int R, float* out, float* minmax, void* workspace,
                void* stream) {
    DAS_ENTER(h);
    if (maps == nullptr || out == nullptr || minmax == nullptr || workspace == nullptr) return DAS_ERR_INVALID_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || R <= 0 || R > H || R > W) return DAS_ERR_INVALID_ARG;
    const int H2 = H - R + 1, W2 = W - R + 1;
    const size_t smem = ((size_t)W + (W + 31) / 32) * sizeof(double);
    if (smem > 200 * 1024 || B > 65535) return DAS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    double* V = static_cast<double*>(workspace);
    dim3 g1((W + 127) / 128, (H2 + kSeg - 1) / kSeg, B);
    DAS_LAUNCH(box_vertical_kernel, g1, 128, 0, st, maps, H, W, R, H2, V);
    DAS_CHECK_LAUNCH();
    if (smem > 48 * 1024)
        DAS_CUDA(cudaFuncSetAttribute(box_horizontal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g2(H2, B);
    DAS_LAUNCH(box_horizontal_kernel, g2, kBoxThreads, smem, st, V, W, R, H2, W2, out, minmax);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_minmax_normalise(das_handle* h, float* score_maps, size_t n, const float* minmax, void* stream) {
    DAS_ENTER(h);
    if (score_maps == nullptr || minmax == nullptr) return DAS_ERR_INVALID_ARG;
    if (n == 0) return DAS_OK;
    const int grid = (int)((n + 255) / 256 < (size_t)(h->num_sms * 8) ? (n + 255) / 256 : (size_t)(h->num_sms * 8));
    DAS_LAUNCH(minmax_normalise_kernel, grid, 256, 0, (cudaStream_t)stream, score_maps, n, minmax);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_nms_sequences(das_handle* h, float* score_maps, int N, int H2, int W2, int R, int kmax, float stop, float* cand_score,
                      int32_t* cand_rc, int32_t* cand_count, long long image_offset, int64_t* cand_flat, void* stream) {
    DAS_ENTER(h);
    if (score_maps == nullptr || cand_score == nullptr || cand_rc == nullptr || cand_count == nullptr)
        return DAS_ERR_INVALID_ARG;
    if (N <= 0 || H2 <= 0 || W2 <= 0 || R <= 0 || kmax <= 0) return DAS_ERR_INVALID_ARG;
    if ((long long)H2 * W2 > 0x7fffffffLL) return DAS_ERR_UNSUPPORTED;
    if (image_offset < 0) return DAS_ERR_INVALID_ARG;
    const size_t tiles = (size_t)((H2 + kNmsTile - 1) / kNmsTile) * ((W2 + kNmsTile - 1) / kNmsTile);
    const size_t smem = tiles * sizeof(unsigned long long);     // one arg-max key per 32 x 32 tile
    if (smem > 200 * 1024) return DAS_ERR_UNSUPPORTED;           // maps beyond ~5000 x 5000
    if (smem > 48 * 1024)
        DAS_CUDA(cudaFuncSetAttribute(nms_sequences_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DAS_LAUNCH(nms_sequences_kernel, N, kNmsThreads, smem, (cudaStream_t)stream, score_maps, H2, W2, R, kmax, stop,
               cand_score, cand_rc, cand_count, image_offset, cand_flat);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

}  // extern "C"
