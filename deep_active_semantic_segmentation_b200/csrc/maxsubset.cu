// Max-subset representativeness (reference active_selection/max_subset.py:17-39): greedy facility location.
//   D[n,i] = ||x_n - y_i||  (sklearn pairwise_distances: float64 arithmetic; float32 inputs give float32 distances)
//   repeat k times: score_i = sum_n min(min_d[n], D[n,i]) for every unselected candidate i; pick the FIRST i with
//   the smallest score (the reference maximises -score with a strict '>'); min_d = min(min_d, D[:, i]).
// The reference spends O(k M N) numpy calls on the host; here the distance matrix is built once (candidate major,
// dist[i][n], so a candidate's distances to the whole pool are one contiguous stream) and every pick is two
// launches: a bandwidth-bound sweep over dist (one CTA per candidate, fixed-order fp64 reduction) that also folds
// the previous pick into min_d, and a one-CTA arg-min.
#include <math.h>

#include "das_common.cuh"

namespace das {

constexpr int kMsThreads = 256;

// ---- distances: 32 x 32 tile per block, 16 x 16 threads x (2 x 2) outputs, fp64 accumulation of exact differences
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) ms_dist_kernel(const TIn* __restrict__ X, const TIn* __restrict__ Y, int N, int M, int D,
                                                      TOut* __restrict__ dist) {
    __shared__ double xs[32][33], ys[32][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int d0 = 0; d0 < D; d0 += 32) {
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            xs[r][c] = (n0 + r < N && d0 + c < D) ? (double)X[(size_t)(n0 + r) * D + d0 + c] : 0.0;
            ys[r][c] = (i0 + r < M && d0 + c < D) ? (double)Y[(size_t)(i0 + r) * D + d0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int d = 0; d < 32; ++d) {
            const double a0 = xs[tx][d], a1 = xs[tx + 16][d], b0 = ys[ty][d], b1 = ys[ty + 16][d];
            double t;
            t = a0 - b0, acc[0][0] = fma(t, t, acc[0][0]);
            t = a1 - b0, acc[0][1] = fma(t, t, acc[0][1]);
            t = a0 - b1, acc[1][0] = fma(t, t, acc[1][0]);
            t = a1 - b1, acc[1][1] = fma(t, t, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + ty + 16 * a, n = n0 + tx + 16 * b;
            if (i < M && n < N) dist[(size_t)i * N + n] = (TOut)sqrt(acc[a][b]);
        }
}

// ---- one pick, part 1: fold the previous pick into min_d (block 0 writes it back) and score every candidate
template <typename T>
__global__ void __launch_bounds__(kMsThreads) ms_score_kernel(const T* __restrict__ dist, int N, int M, const double* __restrict__ min_in,
                                                              double* __restrict__ min_out, const int32_t* __restrict__ picks,
                                                              int step, const uint8_t* __restrict__ selected,
                                                              double* __restrict__ scores) {
    __shared__ double red[kMsThreads / 32];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // picks[step - 1] == -1: every candidate was already taken (k > M), there is no previous pick to fold in
    const int last = step > 0 ? picks[step - 1] : -1;
    const T* prev = last >= 0 ? dist + (size_t)last * N : nullptr;
    const bool skip = selected[i] != 0 || last == i;
    const T* mine = dist + (size_t)i * N;
    double s = 0.0;
    for (int n = tid; n < N; n += kMsThreads) {
        double m = min_in[n];
        if (prev != nullptr) m = fmin(m, (double)prev[n]);
        if (i == 0) min_out[n] = m;  // the running minimum after the previous pick
        if (!skip) s += fmin(m, (double)mine[n]);
    }
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kMsThreads / 32; ++w) s += red[w];
        scores[i] = skip ? INFINITY : s;
    }
}

// ---- part 2: first index of the smallest score among the unselected candidates
__global__ void __launch_bounds__(1024) ms_pick_kernel(const double* __restrict__ scores, int M, uint8_t* selected, int32_t* picks,
                                                       int step) {
    __shared__ double sv[32];
    __shared__ int si[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double bv = INFINITY;
    int bi = 0x7fffffff;
    if (step > 0 && tid == 0 && picks[step - 1] >= 0) selected[picks[step - 1]] = 1;
    for (int i = tid; i < M; i += 1024) {
        const double v = scores[i];
        if (v < bv || (v == bv && i < bi)) bv = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    if (lane == 0) sv[wid] = bv, si[wid] = bi;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; ++w)
            if (sv[w] < bv || (sv[w] == bv && si[w] < bi)) bv = sv[w], bi = si[w];
        picks[step] = isinf(bv) ? -1 : bi;  // -1: every candidate is already selected (the reference appends None)
    }
}

__global__ void ms_init_kernel(double* min_d, int N, uint8_t* selected, int M) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) min_d[i] = INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) selected[i] = 0;
}

struct MsLayout {
    size_t dist, min_a, min_b, scores, selected, total;
};
static MsLayout ms_layout(int N, int M, int f64) {
    MsLayout L;
    size_t off = 0;
    L.dist = off, off += align_up((size_t)N * M * (f64 ? 8 : 4), 256);
    L.min_a = off, off += align_up((size_t)N * 8, 256);
    L.min_b = off, off += align_up((size_t)N * 8, 256);
    L.scores = off, off += align_up((size_t)M * 8, 256);
    L.selected = off, off += align_up((size_t)M, 256);
    L.total = off;
    return L;
}

template <typename TIn, typename TDist>
static int ms_run(const TIn* X, const TIn* Y, int N, int M, int D, int k, int32_t* picks, char* ws, const MsLayout& L,
                  cudaStream_t st) {
    TDist* dist = reinterpret_cast<TDist*>(ws + L.dist);
    double* mins[2] = {reinterpret_cast<double*>(ws + L.min_a), reinterpret_cast<double*>(ws + L.min_b)};
    double* scores = reinterpret_cast<double*>(ws + L.scores);
    uint8_t* selected = reinterpret_cast<uint8_t*>(ws + L.selected);
    DAS_LAUNCH(ms_init_kernel, 64, 256, 0, st, mins[0], N, selected, M);
    DAS_CHECK_LAUNCH();
    DAS_LAUNCH((ms_dist_kernel<TIn, TDist>), dim3((N + 31) / 32, (M + 31) / 32), 256, 0, st, X, Y, N, M, D, dist);
    DAS_CHECK_LAUNCH();
    for (int s = 0; s < k; ++s) {
        DAS_LAUNCH((ms_score_kernel<TDist>), M, kMsThreads, 0, st, dist, N, M, mins[s & 1], mins[(s + 1) & 1], picks, s, selected,
                   scores);
        DAS_CHECK_LAUNCH();
        DAS_LAUNCH(ms_pick_kernel, 1, 1024, 0, st, scores, M, selected, picks, s);
        DAS_CHECK_LAUNCH();
    }
    return DAS_OK;
}

}  // namespace das

using namespace das;

extern "C" {

int das_maxsubset_workspace_bytes(int N, int M, int D, int is_f64, size_t* bytes) {
    if (bytes == nullptr || N <= 0 || M <= 0 || D <= 0) return DAS_ERR_INVALID_ARG;
    *bytes = ms_layout(N, M, is_f64).total;
    return DAS_OK;
}

int das_maxsubset_greedy(das_handle* h, const void* X, const void* Y, int N, int M, int D, int is_f64, int k, int32_t* picks,
                         void* workspace, void* stream) {
    DAS_ENTER(h);
    if (X == nullptr || Y == nullptr || picks == nullptr || workspace == nullptr) return DAS_ERR_INVALID_ARG;
    if (N <= 0 || M <= 0 || D <= 0 || k < 0) return DAS_ERR_INVALID_ARG;
    if (M > 65535 * 32) return DAS_ERR_UNSUPPORTED;
    const MsLayout L = ms_layout(N, M, is_f64);
    char* ws = static_cast<char*>(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f64)
        return ms_run<double, double>(static_cast<const double*>(X), static_cast<const double*>(Y), N, M, D, k, picks, ws, L, st);
    return ms_run<float, float>(static_cast<const float*>(X), static_cast<const float*>(Y), N, M, D, k, picks, ws, L, st);
}

}  // extern "C"
