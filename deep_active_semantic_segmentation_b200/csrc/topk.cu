// K3: block-level radix top-k over (score, position) pairs.
// Replaces the host-side `sorted(zip(scores, images), key=score, reverse=...)[:k]` of the reference
// (mc_dropout.py:195, ceal.py:69,97,130, mc_noise.py:59,128,147).  Python's sort is stable in both
// directions, so ties keep input order; the 64-bit key (orderable(score) << 32 | ~position) makes every
// key unique and its descending order identical to that stable sort.
#include "das_common.cuh"

namespace das {

constexpr int kTopkThreads = 1024;

__device__ __forceinline__ unsigned long long topk_key(float s, uint32_t pos, bool descending) {
    uint32_t o = float_orderable(s);
    if (!descending) o = ~o;
    return ((unsigned long long)o << 32) | (uint32_t)(~pos);
}

// One CTA: (1) MSB-first 8-bit radix select of the k-th largest key (8 passes over the scores, keys are
// recomputed on the fly so no key array is materialised), (2) compaction of the k winners into shared
// memory, (3) bitonic sort, (4) ordered write-out.
__global__ void __launch_bounds__(kTopkThreads) topk_kernel(const float* __restrict__ scores,
                                                            const int64_t* __restrict__ ids, int n, int k,
                                                            int descending, float* __restrict__ out_scores,
                                                            int64_t* __restrict__ out_ids) {
    extern __shared__ unsigned long long keys[];  // next_pow2(k) entries
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long prefix_s;
    __shared__ int krem_s;
    __shared__ unsigned int fill_s;
    const int tid = threadIdx.x;
    const bool desc = descending != 0;

    unsigned long long prefix = 0ull;  // the bits fixed so far
    int krem = k;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i = tid; i < 256; i += kTopkThreads) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kTopkThreads) {
            const unsigned long long key = topk_key(scores[i], (uint32_t)i, desc);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, d = 255;
            for (; d > 0; --d) {
                if (acc + (int)hist[d] >= krem) break;
                acc += hist[d];
            }
            prefix_s = prefix | ((unsigned long long)d << shift);
            krem_s = krem - acc;
        }
        __syncthreads();
        prefix = prefix_s;
        krem = krem_s;
        __syncthreads();
    }
    const unsigned long long kth = prefix;  // exactly k keys are >= kth (keys are unique)

    int cap = 1;
    while (cap < k) cap <<= 1;
    if (tid == 0) fill_s = 0;
    for (int i = tid; i < cap; i += kTopkThreads) keys[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += kTopkThreads) {
        const unsigned long long key = topk_key(scores[i], (uint32_t)i, desc);
        if (key >= kth) keys[atomicAdd(&fill_s, 1u)] = key;
    }
    __syncthreads();
    // bitonic sort, descending
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < cap; i += kTopkThreads) {
                const int j = i ^ stride;
                if (j > i) {
                    const unsigned long long a = keys[i], b = keys[j];
                    const bool up = (i & size) == 0;  // descending blocks first
                    if (up ? (a < b) : (a > b)) keys[i] = b, keys[j] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += kTopkThreads) {
        const uint32_t pos = ~(uint32_t)keys[i];
        out_scores[i] = scores[pos];
        out_ids[i] = ids != nullptr ? ids[pos] : (int64_t)pos;
    }
}

}  // namespace das

using namespace das;

extern "C" {

int das_topk_workspace_bytes(int n, int k, size_t* bytes) {
    if (bytes == nullptr || n < 0 || k < 0) return DAS_ERR_INVALID_ARG;
    *bytes = 0;  // keys are recomputed from the scores; the k winners live in shared memory
    return DAS_OK;
}

int das_topk(const float* scores, const int64_t* ids, int n, int k, int descending, float* out_scores,
             int64_t* out_ids, void* workspace, void* stream) {
    (void)workspace;
    if (n < 0 || k < 0) return DAS_ERR_INVALID_ARG;
    if (k > n) k = n;
    if (k == 0) return DAS_OK;
    if (scores == nullptr || out_scores == nullptr || out_ids == nullptr) return DAS_ERR_INVALID_ARG;
    if (k > DAS_TOPK_MAX_K) return DAS_ERR_UNSUPPORTED;
    int cap = 1;
    while (cap < k) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    DAS_LAUNCH(topk_kernel, 1, kTopkThreads, smem, (cudaStream_t)stream, scores, ids, n, k, descending, out_scores,
               out_ids);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

}  // extern "C"
