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

// Input / output shapes of one launch.  Plain arrays (sstride = istride = 1) or candidate RECORDS: int64 pairs
// {float32 score bits in the low word, id}, the unit the ranks exchange (one all-gather of [k, 2] int64 per rank):
// a record table is read with sstride = 4 (floats), istride = 2 and `neg_ids_last` (padding records carry id -1 and
// rank below every real candidate); it is written when out_records != nullptr (all k_slots slots: the winners with
// id + id_offset, then padding).
struct TopkIo {
    const float* scores;
    const int64_t* ids;
    int sstride, istride, neg_ids_last;
    float* out_scores;
    int64_t* out_ids;
    int64_t* out_records;
    int k_slots;
    long long id_offset;
};

// One CTA: (1) MSB-first 8-bit radix select of the k-th largest key (8 passes over the scores, keys are
// recomputed on the fly so no key array is materialised), (2) compaction of the k winners into shared
// memory, (3) bitonic sort, (4) ordered write-out.
__global__ void __launch_bounds__(kTopkThreads) topk_kernel(const TopkIo io, int n, int k, int descending) {
    const float* __restrict__ scores = io.scores;
    const int64_t* __restrict__ ids = io.ids;
    const size_t ss = (size_t)io.sstride, is = (size_t)io.istride;
    const bool pad_last = io.neg_ids_last != 0 && ids != nullptr;
    auto key_of = [&](int i, bool desc_) -> unsigned long long {
        if (pad_last && ids[(size_t)i * is] < 0) return (unsigned long long)(uint32_t)(~(uint32_t)i);  // below every score
        return topk_key(scores[(size_t)i * ss], (uint32_t)i, desc_);
    };
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
            const unsigned long long key = key_of(i, desc);
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
        const unsigned long long key = key_of(i, desc);
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
        const float sc = scores[(size_t)pos * ss];
        const int64_t id = ids != nullptr ? ids[(size_t)pos * is] : (int64_t)pos;
        if (io.out_records != nullptr) {
            io.out_records[2 * i] = (int64_t)(unsigned long long)__float_as_uint(sc);
            io.out_records[2 * i + 1] = id < 0 ? id : id + io.id_offset;
        } else {
            io.out_scores[i] = sc;
            io.out_ids[i] = id;
        }
    }
    if (io.out_records != nullptr) {  // padding records: never selected by a merge (neg_ids_last), dropped by the host
        const float pad = desc ? -INFINITY : INFINITY;
        for (int i = k + tid; i < io.k_slots; i += kTopkThreads) {
            io.out_records[2 * i] = (int64_t)(unsigned long long)__float_as_uint(pad);
            io.out_records[2 * i + 1] = -1;
        }
    }
}

__global__ void topk_pad_records_kernel(int64_t* records, int k_slots, int descending) {
    const float pad = descending ? -INFINITY : INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k_slots; i += gridDim.x * blockDim.x) {
        records[2 * i] = (int64_t)(unsigned long long)__float_as_uint(pad);
        records[2 * i + 1] = -1;
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

static int topk_launch(das_handle* h, const TopkIo& io, int n, int k, int descending, void* stream) {
    if (k > DAS_TOPK_MAX_K) return DAS_ERR_UNSUPPORTED;
    DAS_ENTER(h);
    int cap = 1;
    while (cap < k) cap <<= 1;
    const size_t smem = (size_t)cap * sizeof(unsigned long long);
    DAS_LAUNCH(topk_kernel, 1, kTopkThreads, smem, (cudaStream_t)stream, io, n, k, descending);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

int das_topk(das_handle* h, const float* scores, const int64_t* ids, int n, int k, int descending, float* out_scores,
             int64_t* out_ids, void* workspace, void* stream) {
    (void)workspace;
    if (n < 0 || k < 0) return DAS_ERR_INVALID_ARG;
    if (k > n) k = n;
    if (k == 0) return DAS_OK;
    if (scores == nullptr || out_scores == nullptr || out_ids == nullptr) return DAS_ERR_INVALID_ARG;
    const TopkIo io = {scores, ids, 1, 1, 0, out_scores, out_ids, nullptr, 0, 0};
    return topk_launch(h, io, n, k, descending, stream);
}

int das_topk_records(das_handle* h, const float* scores, const int64_t* ids, int n, int k_slots, int descending,
                     long long id_offset, int64_t* records, void* stream) {
    if (n < 0 || k_slots <= 0 || records == nullptr) return DAS_ERR_INVALID_ARG;
    if (n > 0 && scores == nullptr) return DAS_ERR_INVALID_ARG;
    if (k_slots > DAS_TOPK_MAX_K) return DAS_ERR_UNSUPPORTED;
    const int k = k_slots < n ? k_slots : n;
    if (k == 0) {  // an empty shard still contributes a full block of padding records to the exchange
        DAS_ENTER(h);
        DAS_LAUNCH(topk_pad_records_kernel, 1, 256, 0, (cudaStream_t)stream, records, k_slots, descending);
        DAS_CHECK_LAUNCH();
        return DAS_OK;
    }
    const TopkIo io = {scores, ids, 1, 1, 0, nullptr, nullptr, records, k_slots, id_offset};
    return topk_launch(h, io, n, k, descending, stream);
}

int das_topk_merge(das_handle* h, const int64_t* records, int n_records, int k_slots, int descending,
                   int64_t* out_records, void* stream) {
    if (records == nullptr || out_records == nullptr || n_records <= 0 || k_slots <= 0) return DAS_ERR_INVALID_ARG;
    if (k_slots > DAS_TOPK_MAX_K) return DAS_ERR_UNSUPPORTED;
    const int k = k_slots < n_records ? k_slots : n_records;
    const TopkIo io = {reinterpret_cast<const float*>(records), records + 1, 4, 2, 1, nullptr, nullptr, out_records, k_slots, 0};
    return topk_launch(h, io, n_records, k, descending, stream);
}

}  // extern "C"
