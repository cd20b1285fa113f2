// K4: core-set k-center greedy (reference active_selection/core_set.py:17-38).
//   min_d2[i] = min over centres of ||f_i - f_c||^2, accumulated in fp64 from float32 features widened
//   exactly (the reference widens the same float32 network outputs to float64, core_set.py:50,63);
//   pick = first argmax (np.argmax, core_set.py:22); update with the new centre (core_set.py:26,37-38).
// One warp per feature row; a step streams the N x D feature matrix once (L2 resident for the
// BASELINE size: 10k x 2048 x 4 B = 82 MB < 126 MB).  Every block leaves its best (value,row) in a small
// table; the NEXT launch starts by reducing that table, so the greedy loop is a chain of launches with
// no host round trip, no atomics and a deterministic tie rule.
//
// With a tensor-core distance filter (gram.cu, das_kcenter_filter_build) the launches screen every row with
// its bf16 distance first and re-evaluate in float64 only the rows whose minimum can change; the values
// written to min_d2 and the picks are bit-identical to the unfiltered path.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "das_common.cuh"
#include "gram.cuh"

namespace das {

constexpr int kKcThreads = 256;
constexpr int kKcWarps = kKcThreads / 32;

struct KcBest {
    double v;
    int idx;
    int pad;
};

__device__ __forceinline__ bool kc_better(double v, int i, double bv, int bi) {
    return v > bv || (v == bv && i < bi);  // larger distance, then lower row index
}

// Squared distance between rows a and b (length D), fp64 accumulation of exact float32 differences.
// The VALUE is defined as (Q0 + Q1) + (Q2 + Q3), where quarter Q_q sums the elements of the loop iterations it with
// it % 4 == q (a warp covers 32 float4 / floats per iteration) - per lane two chains (x,z | y,w components) added
// once, then the warp xor-tree.  One warp can evaluate all four quarters (warp_dist2: init / step / chain kernels) or
// four warps one quarter each (warp_dist2_quarter: the cluster kernel, where the exact re-evaluation of ~11 rows per
// CTA sits on the critical path of every greedy step - measured 4.3 of its 8.7 us): same association, same bits.
template <bool VEC4>
__device__ __forceinline__ double lane_quarter(const float* __restrict__ a, const float* __restrict__ b, int D, int lane, int q) {
    if (VEC4) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        double s0 = 0.0, s1 = 0.0;
        const int n4 = D / 4;
#pragma unroll 4
        for (int i = lane + 32 * q; i < n4; i += 128) {
            const float4 x = a4[i], y = b4[i];
            const double d0 = (double)x.x - (double)y.x, d1 = (double)x.y - (double)y.y;
            const double d2 = (double)x.z - (double)y.z, d3 = (double)x.w - (double)y.w;
            s0 = fma(d0, d0, s0);
            s1 = fma(d1, d1, s1);
            s0 = fma(d2, d2, s0);
            s1 = fma(d3, d3, s1);
        }
        return s0 + s1;
    }
    double s = 0.0;
    for (int i = lane + 32 * q; i < D; i += 128) {
        const double d = (double)a[i] - (double)b[i];
        s = fma(d, d, s);
    }
    return s;
}
template <bool VEC4>
__device__ __forceinline__ double warp_dist2_quarter(const float* __restrict__ a, const float* __restrict__ b, int D, int lane,
                                                     int q) {
    return warp_sum(lane_quarter<VEC4>(a, b, D, lane, q));
}
template <bool VEC4>
__device__ __forceinline__ double warp_dist2(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
    if (VEC4) {
        // all four quarters in one sweep (eight independent chains, 8 loads in flight per iteration)
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        double s[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        const int n4 = D / 4;
        int i = lane;
#pragma unroll 2
        for (; i + 96 < n4; i += 128) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = a4[i + 32 * q], y = b4[i + 32 * q];
                const double d0 = (double)x.x - (double)y.x, d1 = (double)x.y - (double)y.y;
                const double d2 = (double)x.z - (double)y.z, d3 = (double)x.w - (double)y.w;
                s[q][0] = fma(d0, d0, s[q][0]);
                s[q][1] = fma(d1, d1, s[q][1]);
                s[q][0] = fma(d2, d2, s[q][0]);
                s[q][1] = fma(d3, d3, s[q][1]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // ragged tail: every remaining iteration stays in its quarter
            if (i + 32 * q < n4) {
                const float4 x = a4[i + 32 * q], y = b4[i + 32 * q];
                const double d0 = (double)x.x - (double)y.x, d1 = (double)x.y - (double)y.y;
                const double d2 = (double)x.z - (double)y.z, d3 = (double)x.w - (double)y.w;
                s[q][0] = fma(d0, d0, s[q][0]);
                s[q][1] = fma(d1, d1, s[q][1]);
                s[q][0] = fma(d2, d2, s[q][0]);
                s[q][1] = fma(d3, d3, s[q][1]);
            }
        }
        const double q0 = warp_sum(s[0][0] + s[0][1]), q1 = warp_sum(s[1][0] + s[1][1]);
        const double q2 = warp_sum(s[2][0] + s[2][1]), q3 = warp_sum(s[3][0] + s[3][1]);
        return (q0 + q1) + (q2 + q3);
    }
    const double q0 = warp_sum(lane_quarter<false>(a, b, D, lane, 0)), q1 = warp_sum(lane_quarter<false>(a, b, D, lane, 1));
    const double q2 = warp_sum(lane_quarter<false>(a, b, D, lane, 2)), q3 = warp_sum(lane_quarter<false>(a, b, D, lane, 3));
    return (q0 + q1) + (q2 + q3);
}

// MODE 0: init  - min over the L given centres              (core_set.py:19)
// MODE 1: step  - centre index read from *centre_ptr        (multi-GPU: host all-reduced it)
// MODE 2: chain - centre = argmax of the previous launch's block table; block 0 records the pick
template <bool VEC4, int MODE>
__global__ void __launch_bounds__(kKcThreads) kcenter_kernel(const float* __restrict__ feats, int D, int row_begin,
                                                             int row_end, const int32_t* __restrict__ centres, int L,
                                                             double* __restrict__ min_d2,
                                                             const KcBest* __restrict__ prev_best, int n_prev,
                                                             KcBest* __restrict__ next_best, int32_t* picks, int step) {
    __shared__ double sv[kKcWarps];
    __shared__ int si[kKcWarps];
    __shared__ int centre_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    int centre = -1;
    if (MODE == 1) centre = centres[0];
    if (MODE == 2) {
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int i = tid; i < n_prev; i += kKcThreads) {
            const KcBest q = prev_best[i];
            if (q.idx >= 0 && kc_better(q.v, q.idx, bv, bi)) bv = q.v, bi = q.idx;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (kc_better(ov, oi, bv, bi)) bv = ov, bi = oi;
        }
        if (lane == 0) sv[wid] = bv, si[wid] = bi;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kKcWarps; ++w)
                if (kc_better(sv[w], si[w], bv, bi)) bv = sv[w], bi = si[w];
            centre_s = bi;
            if (blockIdx.x == 0 && picks != nullptr) picks[step] = bi;
        }
        __syncthreads();
        centre = centre_s;
        __syncthreads();
    }

    double bv = -1.0;
    int bi = -1;
    for (int row = row_begin + blockIdx.x * kKcWarps + wid; row < row_end; row += gridDim.x * kKcWarps) {
        const float* fr = feats + (size_t)row * D;
        double m;
        if (MODE == 0) {
            m = INFINITY;
            for (int l = 0; l < L; ++l) m = fmin(m, warp_dist2<VEC4>(fr, feats + (size_t)centres[l] * D, D, lane));
        } else {
            m = fmin(min_d2[row - row_begin], warp_dist2<VEC4>(fr, feats + (size_t)centre * D, D, lane));
        }
        if (lane == 0) min_d2[row - row_begin] = m;
        if (bi < 0 || kc_better(m, row, bv, bi)) bv = m, bi = row;
    }
    if (lane == 0) sv[wid] = bv, si[wid] = bi;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kKcWarps; ++w)
            if (si[w] >= 0 && (bi < 0 || kc_better(sv[w], si[w], bv, bi))) bv = sv[w], bi = si[w];
        KcBest q;
        q.v = bv, q.idx = bi, q.pad = 0;
        next_best[blockIdx.x] = q;
    }
}


// ---- tensor-core filtered variants --------------------------------------------------------------
struct KcFilter {
    const float* dt;            // [N, ld] screening distances, dt[c * ld + (row - row_begin)]
    const double* nrm;          // [N] exact squared norms
    unsigned long long* stats;  // [0] exact re-evaluations, [1] rows screened
    int ld;
};

// arg-max of the previous launch's block table (every block computes the same answer)
__device__ __forceinline__ int kc_table_argmax(const KcBest* __restrict__ prev_best, int n_prev, double* sv, int* si,
                                               int* centre_s, int32_t* picks, int step) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int i = tid; i < n_prev; i += kKcThreads) {
        const KcBest q = prev_best[i];
        if (q.idx >= 0 && kc_better(q.v, q.idx, bv, bi)) bv = q.v, bi = q.idx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (kc_better(ov, oi, bv, bi)) bv = ov, bi = oi;
    }
    if (lane == 0) sv[wid] = bv, si[wid] = bi;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kKcWarps; ++w)
            if (kc_better(sv[w], si[w], bv, bi)) bv = sv[w], bi = si[w];
        *centre_s = bi;
        if (blockIdx.x == 0 && picks != nullptr) picks[step] = bi;
    }
    __syncthreads();
    const int c = *centre_s;
    __syncthreads();
    return c;
}

// Filtered init (core_set.py:19): one warp per row.  ub = min_l (d2~ + margin) bounds the true minimum from
// above; only centres with d2~ - margin <= ub can attain it and are evaluated exactly.
template <bool VEC4>
__global__ void __launch_bounds__(kKcThreads) kcenter_finit_kernel(const float* __restrict__ feats, int D, int row_begin,
                                                                   int row_end, const int32_t* __restrict__ centres, int L,
                                                                   double* __restrict__ min_d2, KcBest* __restrict__ next_best,
                                                                   const KcFilter f) {
    __shared__ double sv[kKcWarps];
    __shared__ int si[kKcWarps];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double bv = -1.0;
    int bi = -1;
    unsigned long long n_exact = 0;
    for (int row = row_begin + blockIdx.x * kKcWarps + wid; row < row_end; row += gridDim.x * kKcWarps) {
        const float* fr = feats + (size_t)row * D;
        const double nr = f.nrm[row];
        double ub = INFINITY;
        for (int l = lane; l < L; l += 32) {
            const int c = centres[l];
            const double mg = kFilterDelta * (nr + f.nrm[c]);
            ub = fmin(ub, (double)f.dt[(size_t)c * f.ld + (row - row_begin)] + mg);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ub = fmin(ub, __shfl_xor_sync(0xffffffffu, ub, o));
        double m = INFINITY;
        for (int l0 = 0; l0 < L; l0 += 32) {
            const int l = l0 + lane;
            bool need = false;
            if (l < L) {
                const int c = centres[l];
                const double mg = kFilterDelta * (nr + f.nrm[c]);
                need = (double)f.dt[(size_t)c * f.ld + (row - row_begin)] - mg <= ub;
            }
            unsigned mask = __ballot_sync(0xffffffffu, need);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                m = fmin(m, warp_dist2<VEC4>(fr, feats + (size_t)centres[l0 + j] * D, D, lane));
                ++n_exact;
            }
        }
        if (lane == 0) min_d2[row - row_begin] = m;
        if (bi < 0 || kc_better(m, row, bv, bi)) bv = m, bi = row;
    }
    if (lane == 0) sv[wid] = bv, si[wid] = bi;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kKcWarps; ++w)
            if (si[w] >= 0 && (bi < 0 || kc_better(sv[w], si[w], bv, bi))) bv = sv[w], bi = si[w];
        KcBest q;
        q.v = bv, q.idx = bi, q.pad = 0;
        next_best[blockIdx.x] = q;
    }
    if (lane == 0 && n_exact) atomicAdd(f.stats, n_exact);
}

// Filtered greedy step (core_set.py:26,37-38): one THREAD per row screens with the tensor-core distance;
// rows that may change go to a shared-memory work list and are evaluated exactly by whole warps.
template <bool VEC4, int MODE>
__global__ void __launch_bounds__(kKcThreads) kcenter_fstep_kernel(const float* __restrict__ feats, int D, int row_begin,
                                                                   int row_end, const int32_t* __restrict__ centres,
                                                                   double* __restrict__ min_d2,
                                                                   const KcBest* __restrict__ prev_best, int n_prev,
                                                                   KcBest* __restrict__ next_best, int32_t* picks, int step,
                                                                   const KcFilter f, int stage_centre) {
    __shared__ double sv[kKcWarps];
    __shared__ int si[kKcWarps];
    __shared__ int centre_s;
    __shared__ int wl_row[kKcThreads];
    __shared__ double wl_val[kKcThreads];
    __shared__ int wl_n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int centre = MODE == 1 ? centres[0] : kc_table_argmax(prev_best, n_prev, sv, si, &centre_s, picks, step);
    const double nc = f.nrm[centre];
    const float* dtc = f.dt + (size_t)centre * f.ld;
    // every exact evaluation of this launch is against the same centre row: stage it in shared memory once
    extern __shared__ float4 centre_row[];
    const float* fc = feats + (size_t)centre * D;
    if (stage_centre) {
        if (VEC4) {
            const float4* src = reinterpret_cast<const float4*>(fc);
            for (int i = tid; i < D / 4; i += kKcThreads) centre_row[i] = src[i];
        } else {
            float* dst = reinterpret_cast<float*>(centre_row);
            for (int i = tid; i < D; i += kKcThreads) dst[i] = fc[i];
        }
        fc = reinterpret_cast<const float*>(centre_row);  // visible after the first __syncthreads() of the row loop
    }

    double bv = -1.0;
    int bi = -1;
    for (int base = row_begin + blockIdx.x * kKcThreads; base < row_end; base += gridDim.x * kKcThreads) {
        if (tid == 0) wl_n = 0;
        __syncthreads();
        const int row = base + tid;
        const bool in = row < row_end;
        double m = 0.0;
        bool need = false;
        if (in) {
            m = min_d2[row - row_begin];
            need = (double)dtc[row - row_begin] - kFilterDelta * (f.nrm[row] + nc) <= m;
        }
        int slot = -1;
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&wl_n, __popc(mask));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (need) {
                slot = wbase + __popc(mask & ((1u << lane) - 1u));
                wl_row[slot] = row;
            }
        }
        __syncthreads();
        const int n_work = wl_n;
        for (int k = wid; k < n_work; k += kKcWarps) {
            const double d = warp_dist2<VEC4>(feats + (size_t)wl_row[k] * D, fc, D, lane);
            if (lane == 0) wl_val[k] = d;
        }
        __syncthreads();
        if (need) {
            m = fmin(m, wl_val[slot]);
            min_d2[row - row_begin] = m;
        }
        if (in && (bi < 0 || kc_better(m, row, bv, bi))) bv = m, bi = row;
        if (tid == 0 && n_work) atomicAdd(f.stats, (unsigned long long)n_work);
        __syncthreads();
    }
    // block arg-max
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || kc_better(ov, oi, bv, bi))) bv = ov, bi = oi;
    }
    if (lane == 0) sv[wid] = bv, si[wid] = bi;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kKcWarps; ++w)
            if (si[w] >= 0 && (bi < 0 || kc_better(sv[w], si[w], bv, bi))) bv = sv[w], bi = si[w];
        KcBest q;
        q.v = bv, q.idx = bi, q.pad = 0;
        next_best[blockIdx.x] = q;
        if (blockIdx.x == 0) atomicAdd(f.stats + 1, (unsigned long long)(row_end - row_begin));
    }
}


// ---- whole greedy loop in ONE thread-block cluster (single-GPU path with a filter) -----------------
// 500 dependent steps of a few microseconds each are launch-latency bound as a chain of kernels (~9 us per
// step measured).  Here 4 or 8 CTAs x 1024 threads of one cluster keep min_d2 and the row norms in REGISTERS
// (up to kClRows rows per thread), stage the centre row in shared memory, exchange each CTA's arg-max
// through distributed shared memory (st.shared::cluster) and meet at one hardware cluster barrier per
// step: no kernel boundary, no global-memory round trip for the arg-max.
constexpr int kClCtasMax = 8;   // CTAs per cluster: 4 or 8 (the portable maximum), chosen from N at launch
constexpr int kClThreads = 1024;
constexpr int kClWarps = kClThreads / 32;
constexpr int kClRows = 4;  // rows per thread -> N <= 8 * 1024 * 4 = 32768

// What a CTA publishes per step, in every peer's shared memory: ONE 16-byte remote store {value, row, tag = step + 1}.
// A 16-byte aligned vector store to shared memory is a single transaction, so a reader that sees the tag of the step
// also sees the value and the row that came with it: the peers poll their own copy of the table instead of meeting at a
// hardware cluster barrier (whose release half is a MEMBAR.ALL.GPU per step).  Two tables (step parity): a CTA can be
// at most one step ahead of the slowest one, because its next record needs every peer's record of this step.
struct __align__(16) ClBest {
    double v;
    int idx;
    unsigned int tag;
};

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool VEC4, int kClCtas>
__global__ void __launch_bounds__(kClThreads, 1)
    kcenter_cluster_kernel(const float* __restrict__ feats, int N, int D, double* __restrict__ min_d2, int K,
                           int32_t* __restrict__ picks, double* __restrict__ min_d_out, const KcFilter f, int stage_centre) {
    extern __shared__ float4 cl_dyn[];  // [centre row][wl_val f64 x 4 quarters x 4096][wl_row i32 x 4096]
    __shared__ ClBest slots[2][kClCtas];
    __shared__ double sv[kClWarps];
    __shared__ int si[kClWarps];
    __shared__ int wl_n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const size_t row_f4 = stage_centre ? ((size_t)D * sizeof(float) + 15) / 16 : 0;
    float* centre_row = reinterpret_cast<float*>(cl_dyn);
    double* wl_val = reinterpret_cast<double*>(cl_dyn + row_f4);           // [slot][quarter]
    int* wl_row = reinterpret_cast<int*>(wl_val + 4 * kClThreads * kClRows);
    if (tid == 0) wl_n = 0;
    if (tid < 2 * kClCtas) {
        ClBest z;
        z.v = 0.0, z.idx = -1, z.tag = 0u;
        (&slots[0][0])[tid] = z;
    }
    __syncthreads();

    // rows of this thread: g + j * 8192 (coalesced across the cluster for every j)
    const int g = (int)rank * kClThreads + tid;
    double m[kClRows], nr[kClRows];
#pragma unroll
    for (int j = 0; j < kClRows; ++j) {
        const int row = g + j * kClCtas * kClThreads;
        m[j] = row < N ? min_d2[row] : -1.0;
        nr[j] = row < N ? f.nrm[row] : 0.0;
    }
    unsigned long long n_exact = 0;
    int parity = 0;
    cluster_barrier();  // every CTA of the cluster is resident before the first remote store

    for (int s = 0; s < K; ++s) {
        // ---- local arg-max of the current min_d2 ----
        double bv = -1.0;
        int bi = -1;
#pragma unroll
        for (int j = 0; j < kClRows; ++j) {
            const int row = g + j * kClCtas * kClThreads;
            if (row < N && (bi < 0 || kc_better(m[j], row, bv, bi))) bv = m[j], bi = row;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || kc_better(ov, oi, bv, bi))) bv = ov, bi = oi;
        }
        if (lane == 0) sv[wid] = bv, si[wid] = bi;
        __syncthreads();
        if (wid == 0) {
            bv = sv[lane], bi = si[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || kc_better(ov, oi, bv, bi))) bv = ov, bi = oi;
            }
            if (lane < kClCtas) {  // lane t publishes this CTA's best in CTA t's slot table: one 16-byte store
                const uint32_t local = (uint32_t)__cvta_generic_to_shared(&slots[parity][rank]);
                uint32_t remote;
                asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(lane));
                const unsigned long long vb = (unsigned long long)__double_as_longlong(bv);
                asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"((uint32_t)vb),
                             "r"((uint32_t)(vb >> 32)), "r"((uint32_t)bi), "r"((uint32_t)(s + 1))
                             : "memory");
            }
        }
        // ---- every thread waits for the 8 records of this step and reduces them: the next centre ----
        double cv = -1.0;
        int centre = -1;
        {
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(&slots[parity][0]);
#pragma unroll
            for (int r = 0; r < kClCtas; ++r) {
                uint32_t w0, w1, w2, w3;
                do {
                    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                 : "r"(base + 16u * r)
                                 : "memory");
                } while (w3 != (uint32_t)(s + 1));
                const double qv = __longlong_as_double((long long)(((unsigned long long)w1 << 32) | w0));
                const int qi = (int)w2;
                if (qi >= 0 && (centre < 0 || kc_better(qv, qi, cv, centre))) cv = qv, centre = qi;
            }
        }
        parity ^= 1;
        const double nc = f.nrm[centre];  // L2 resident; overlaps the screening loads below
        if (rank == 0 && tid == 0) picks[s] = centre;

        // ---- screen own rows with the tensor-core distances, exact float64 for the rest ----
        // (the screening distances are requested first so that their HBM latency overlaps the centre-row copy)
        const float* dtc = f.dt + (size_t)centre * f.ld;
        float dtv[kClRows];
#pragma unroll
        for (int j = 0; j < kClRows; ++j) {
            const int row = g + j * kClCtas * kClThreads;
            dtv[j] = row < N ? dtc[row] : 0.f;
        }
        const float* fc = feats + (size_t)centre * D;
        if (stage_centre) {
            if (VEC4) {
                const float4* src = reinterpret_cast<const float4*>(fc);
                for (int i = tid; i < D / 4; i += kClThreads) cl_dyn[i] = src[i];
            } else {
                for (int i = tid; i < D; i += kClThreads) centre_row[i] = fc[i];
            }
            fc = centre_row;
        }
        __syncthreads();  // (kept on purpose: without it - wl_n is already reset, the staged row is only read after the
                          //  next barrier - the loop measured 10 % SLOWER, 3.34 vs 3.03 ms: the warps then drift apart)
        int slot[kClRows];
#pragma unroll
        for (int j = 0; j < kClRows; ++j) {
            const int row = g + j * kClCtas * kClThreads;
            bool need = false;
            if (row < N) need = (double)dtv[j] - kFilterDelta * (nr[j] + nc) <= m[j];
            slot[j] = -1;
            const unsigned mask = __ballot_sync(0xffffffffu, need);
            if (mask) {
                int wbase = 0;
                if (lane == 0) wbase = atomicAdd(&wl_n, __popc(mask));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (need) {
                    slot[j] = wbase + __popc(mask & ((1u << lane) - 1u));
                    wl_row[slot[j]] = row;
                }
            }
        }
        __syncthreads();
        const int n_work = wl_n;
        // four warps per work-list row, one quarter of the distance each (see warp_dist2): a quarter of the latency
        for (int k4 = wid; k4 < 4 * n_work; k4 += kClWarps) {
            const double d = warp_dist2_quarter<VEC4>(feats + (size_t)wl_row[k4 >> 2] * D, fc, D, lane, k4 & 3);
            if (lane == 0) wl_val[k4] = d;
        }
        __syncthreads();
        if (tid == 0) n_exact += (unsigned long long)n_work, wl_n = 0;  // every thread has read n_work; next use is after a barrier
#pragma unroll
        for (int j = 0; j < kClRows; ++j)
            if (slot[j] >= 0) {
                const double* p4 = wl_val + 4 * slot[j];
                m[j] = fmin(m[j], (p4[0] + p4[1]) + (p4[2] + p4[3]));
            }
    }

#pragma unroll
    for (int j = 0; j < kClRows; ++j) {
        const int row = g + j * kClCtas * kClThreads;
        if (row < N) {
            min_d2[row] = m[j];
            min_d_out[row] = sqrt(m[j]);
        }
    }
    if (tid == 0) {
        if (n_exact) atomicAdd(f.stats, n_exact);
        if (rank == 0) atomicAdd(f.stats + 1, (unsigned long long)N * K);
    }
    cluster_barrier();  // no CTA leaves while a peer could still address its shared memory
}

// reduce a block table to the packed pair key2 = {fp64 bits of the max, row index}
__global__ void kcenter_key_kernel(const KcBest* best, int n, unsigned long long* key2, int32_t* picks, int step) {
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int i = 0; i < n; ++i)
        if (best[i].idx >= 0 && kc_better(best[i].v, best[i].idx, bv, bi)) bv = best[i].v, bi = best[i].idx;
    if (key2 != nullptr) {
        key2[0] = (unsigned long long)__double_as_longlong(bv < 0.0 ? 0.0 : bv);
        key2[1] = (unsigned long long)(unsigned int)bi;
    }
    if (picks != nullptr) picks[step] = bi;
}

__global__ void kcenter_sqrt_kernel(const double* d2, int n, double* d) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = sqrt(d2[i]);
}

static int kc_grid(const das_handle* h, int rows) {
    const int want = (rows + kKcWarps - 1) / kKcWarps;
    const int cap = h->num_sms * 4;
    return want < cap ? (want < 1 ? 1 : want) : cap;
}
// filtered step: one thread per row
static int kc_fgrid(const das_handle* h, int rows) {
    const int want = (rows + kKcThreads - 1) / kKcThreads;
    const int cap = h->num_sms * 4;
    return want < cap ? (want < 1 ? 1 : want) : cap;
}
static bool kc_vec4(const float* feats, int D) { return D % 4 == 0 && aligned16(feats); }

struct KcWorkspace {
    KcBest* best[2];
    double* d2;
};
static KcWorkspace kc_carve(const das_handle* h, void* ws, int N) {
    KcWorkspace w;
    char* p = static_cast<char*>(ws);
    const size_t tbl = align_up((size_t)h->num_sms * 4 * sizeof(KcBest), 256);
    w.best[0] = reinterpret_cast<KcBest*>(p);
    w.best[1] = reinterpret_cast<KcBest*>(p + tbl);
    w.d2 = reinterpret_cast<double*>(p + 2 * tbl);
    (void)N;
    return w;
}

static KcFilter kc_filter_view(const void* filter, int N, int D, int rows) {
    const KcFilterLayout L = kc_filter_layout(N, D, rows);
    char* base = static_cast<char*>(const_cast<void*>(filter));
    KcFilter f;
    f.dt = reinterpret_cast<const float*>(base + L.dt);
    f.nrm = reinterpret_cast<const double*>(base + L.nrm64);
    f.stats = reinterpret_cast<unsigned long long*>(base + L.stats);
    f.ld = L.ld;
    return f;
}

template <int MODE>
static int kc_launch(bool v4, int grid, cudaStream_t st, const float* feats, int D, int rb, int re, const int32_t* centres,
                     int L, double* min_d2, const KcBest* prev, int n_prev, KcBest* next, int32_t* picks, int step) {
    if (v4)
        DAS_LAUNCH((kcenter_kernel<true, MODE>), grid, kKcThreads, 0, st, feats, D, rb, re, centres, L, min_d2, prev,
                   n_prev, next, picks, step);
    else
        DAS_LAUNCH((kcenter_kernel<false, MODE>), grid, kKcThreads, 0, st, feats, D, rb, re, centres, L, min_d2, prev,
                   n_prev, next, picks, step);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

static int kc_launch_finit(bool v4, int grid, cudaStream_t st, const float* feats, int D, int rb, int re,
                           const int32_t* centres, int L, double* min_d2, KcBest* next, const KcFilter& f) {
    if (v4)
        DAS_LAUNCH((kcenter_finit_kernel<true>), grid, kKcThreads, 0, st, feats, D, rb, re, centres, L, min_d2, next, f);
    else
        DAS_LAUNCH((kcenter_finit_kernel<false>), grid, kKcThreads, 0, st, feats, D, rb, re, centres, L, min_d2, next, f);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

template <int MODE>
static int kc_launch_fstep(bool v4, int grid, cudaStream_t st, const float* feats, int D, int rb, int re,
                           const int32_t* centres, double* min_d2, const KcBest* prev, int n_prev, KcBest* next,
                           int32_t* picks, int step, const KcFilter& f) {
    const size_t row_bytes = (size_t)D * sizeof(float);
    const int stage_centre = row_bytes <= 40 * 1024 ? 1 : 0;  // fits the default dynamic shared-memory limit
    const size_t smem = stage_centre ? align_up(row_bytes, 16) : 0;
    if (v4)
        DAS_LAUNCH((kcenter_fstep_kernel<true, MODE>), grid, kKcThreads, smem, st, feats, D, rb, re, centres, min_d2, prev,
                   n_prev, next, picks, step, f, stage_centre);
    else
        DAS_LAUNCH((kcenter_fstep_kernel<false, MODE>), grid, kKcThreads, smem, st, feats, D, rb, re, centres, min_d2, prev,
                   n_prev, next, picks, step, f, stage_centre);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

// One thread-block cluster of NC CTAs (cluster dimension as a launch attribute) runs the whole greedy loop.
template <bool VEC4, int NC>
static int kc_launch_cluster_t(size_t smem, cudaStream_t st, const float* feats, int N, int D, double* d2, int K, int32_t* picks,
                               double* min_d, const KcFilter& f, int stage_centre) {
    auto kernel = kcenter_cluster_kernel<VEC4, NC>;
    DAS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(NC), cfg.blockDim = dim3(kClThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    DAS_CUDA(cudaLaunchKernelEx(&cfg, kernel, feats, N, D, d2, K, picks, min_d, f, stage_centre));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return DAS_OK;
}
static int kc_launch_cluster(bool v4, int nc, size_t smem, cudaStream_t st, const float* feats, int N, int D, double* d2, int K,
                             int32_t* picks, double* min_d, const KcFilter& f, int stage_centre) {
    if (nc == 4)
        return v4 ? kc_launch_cluster_t<true, 4>(smem, st, feats, N, D, d2, K, picks, min_d, f, stage_centre)
                  : kc_launch_cluster_t<false, 4>(smem, st, feats, N, D, d2, K, picks, min_d, f, stage_centre);
    return v4 ? kc_launch_cluster_t<true, 8>(smem, st, feats, N, D, d2, K, picks, min_d, f, stage_centre)
              : kc_launch_cluster_t<false, 8>(smem, st, feats, N, D, d2, K, picks, min_d, f, stage_centre);
}

}  // namespace das

using namespace das;

// scratch table of the step-wise (multi-GPU) entry points: owned by the handle, allocated on first use
static int ensure_step_table(das_handle* h, KcBest** table) {
    if (h->kc_step_table == nullptr) DAS_CUDA(cudaMalloc(&h->kc_step_table, (size_t)h->num_sms * 4 * sizeof(KcBest)));
    *table = static_cast<KcBest*>(h->kc_step_table);
    return DAS_OK;
}

extern "C" {

int das_kcenter_init(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end, const int32_t* centers,
                     int L, double* min_d2, unsigned long long* key2, const void* filter, void* stream) {
    DAS_ENTER(h);
    if (feats == nullptr || centers == nullptr || min_d2 == nullptr || key2 == nullptr) return DAS_ERR_INVALID_ARG;
    if (N <= 0 || D <= 0 || L <= 0 || row_begin < 0 || row_end > N || row_begin >= row_end) return DAS_ERR_INVALID_ARG;
    KcBest* g_step_table = nullptr;
    int rc = ensure_step_table(h, &g_step_table);
    if (rc != DAS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = kc_grid(h, row_end - row_begin);
    if (filter != nullptr)
        rc = kc_launch_finit(kc_vec4(feats, D), grid, st, feats, D, row_begin, row_end, centers, L, min_d2, g_step_table,
                             kc_filter_view(filter, N, D, row_end - row_begin));
    else
        rc = kc_launch<0>(kc_vec4(feats, D), grid, st, feats, D, row_begin, row_end, centers, L, min_d2, nullptr, 0,
                          g_step_table, nullptr, 0);
    if (rc != DAS_OK) return rc;
    DAS_LAUNCH(kcenter_key_kernel, 1, 1, 0, st, g_step_table, grid, key2, nullptr, 0);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_kcenter_step(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end,
                     const int32_t* centre_idx, double* min_d2, unsigned long long* key2, const void* filter,
                     void* stream) {
    DAS_ENTER(h);
    if (feats == nullptr || centre_idx == nullptr || min_d2 == nullptr || key2 == nullptr) return DAS_ERR_INVALID_ARG;
    if (N <= 0 || D <= 0 || row_begin < 0 || row_end > N || row_begin >= row_end) return DAS_ERR_INVALID_ARG;
    KcBest* g_step_table = nullptr;
    int rc = ensure_step_table(h, &g_step_table);
    if (rc != DAS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int grid;
    if (filter != nullptr) {
        grid = kc_fgrid(h, row_end - row_begin);
        rc = kc_launch_fstep<1>(kc_vec4(feats, D), grid, st, feats, D, row_begin, row_end, centre_idx, min_d2, nullptr, 0,
                                g_step_table, nullptr, 0, kc_filter_view(filter, N, D, row_end - row_begin));
    } else {
        grid = kc_grid(h, row_end - row_begin);
        rc = kc_launch<1>(kc_vec4(feats, D), grid, st, feats, D, row_begin, row_end, centre_idx, 1, min_d2, nullptr, 0,
                          g_step_table, nullptr, 0);
    }
    if (rc != DAS_OK) return rc;
    DAS_LAUNCH(kcenter_key_kernel, 1, 1, 0, st, g_step_table, grid, key2, nullptr, 0);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_kcenter_workspace_bytes(const das_handle* h, int N, int D, size_t* bytes) {
    if (h == nullptr || h->magic != kDasHandleMagic) return DAS_ERR_INVALID_ARG;
    if (bytes == nullptr || N <= 0 || D <= 0) return DAS_ERR_INVALID_ARG;
    *bytes = 2 * align_up((size_t)h->num_sms * 4 * sizeof(KcBest), 256) + align_up((size_t)N * sizeof(double), 256);
    return DAS_OK;
}

int das_kcenter_greedy(das_handle* h, const float* feats, int N, int D, const int32_t* centers, int L, int K,
                       int32_t* picks, double* min_d, void* workspace, const void* filter, void* stream) {
    DAS_ENTER(h);
    if (feats == nullptr || centers == nullptr || picks == nullptr || min_d == nullptr || workspace == nullptr)
        return DAS_ERR_INVALID_ARG;
    if (N <= 0 || D <= 0 || L <= 0 || K < 0) return DAS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const KcWorkspace w = kc_carve(h, workspace, N);
    const bool v4 = kc_vec4(feats, D);
    int rc;
    if (filter != nullptr) {
        const KcFilter f = kc_filter_view(filter, N, D, N);
        int grid = kc_grid(h, N);
        rc = kc_launch_finit(v4, grid, st, feats, D, 0, N, centers, L, w.d2, w.best[0], f);
        if (rc != DAS_OK) return rc;
        if (N <= kClCtasMax * kClThreads * kClRows && h->opt[DAS_OPT_KC_CLUSTER]) {
            const size_t row_bytes = (size_t)D * sizeof(float);
            const int stage_centre = row_bytes <= 64 * 1024 ? 1 : 0;  // + 144 KB of work-list tables <= 227 KB
            const size_t smem = (stage_centre ? align_up(row_bytes, 16) : 0) + (size_t)kClThreads * kClRows * (4 * 8 + 4);
            // the smallest cluster that holds the rows: 4 CTAs up to 16 384 rows (measured 2.94 ms against 3.03 ms with 8
            // and 3.57 ms with a non-portable 16-CTA cluster at N = 10 000 - every record more costs exchange latency)
            const int nc = N <= 4 * kClThreads * kClRows ? 4 : 8;
            return kc_launch_cluster(v4, nc, smem, st, feats, N, D, w.d2, K, picks, min_d, f, stage_centre);
        }
        int n_prev = grid;
        grid = kc_fgrid(h, N);
        for (int s = 0; s < K; ++s) {
            rc = kc_launch_fstep<2>(v4, grid, st, feats, D, 0, N, nullptr, w.d2, w.best[s & 1], n_prev, w.best[(s + 1) & 1],
                                    picks, s, f);
            if (rc != DAS_OK) return rc;
            n_prev = grid;
        }
    } else {
        const int grid = kc_grid(h, N);
        rc = kc_launch<0>(v4, grid, st, feats, D, 0, N, centers, L, w.d2, nullptr, 0, w.best[0], nullptr, 0);
        if (rc != DAS_OK) return rc;
        for (int s = 0; s < K; ++s) {
            rc = kc_launch<2>(v4, grid, st, feats, D, 0, N, nullptr, 0, w.d2, w.best[s & 1], grid, w.best[(s + 1) & 1],
                              picks, s);
            if (rc != DAS_OK) return rc;
        }
    }
    DAS_LAUNCH(kcenter_sqrt_kernel, kc_grid(h, N), kKcThreads, 0, st, w.d2, N, min_d);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

/* host copy of the filter counters: stats[0] = rows re-evaluated exactly, stats[1] = rows screened */
int das_kcenter_filter_stats(das_handle* h, const void* filter, int N, int D, int rows, unsigned long long* stats2,
                             void* stream) {
    DAS_ENTER(h);
    if (filter == nullptr || stats2 == nullptr || N <= 0 || D <= 0 || rows <= 0 || rows > N) return DAS_ERR_INVALID_ARG;
    const KcFilter f = kc_filter_view(filter, N, D, rows);
    DAS_CUDA(cudaMemcpyAsync(stats2, f.stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DAS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return DAS_OK;
}

}  // extern "C"
