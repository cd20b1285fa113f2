// das_handle: the per-device state behind the C ABI (include/das_b200.h) - device ordinal and SM count, tuning options
// (the environment is read once, here), lazily allocated k-center step scratch, a cache of encoded TMA descriptors - plus
// the library-wide bookkeeping (launch counter, last CUDA error of the calling thread, status strings).
#include <stdlib.h>
#include <string.h>

#include <new>

#include "das_common.cuh"
#include "tma_host.cuh"

namespace das {

thread_local int g_last_cuda_error = 0;
std::atomic<unsigned long long> g_launch_count{0};

// DAS_OPT_MC_UP_WARPS: 4 | 15 = pixel-pair kernel, 220 | 216 = one-pixel-per-lane kernel (the variants mc_inst.cu builds)
static bool up_warps_option_ok(int v) { return v == 4 || v == 15 || v == 220 || v == 216; }

static int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    if (e == nullptr || e[0] == '\0') return fallback;
    return atoi(e);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<EncodeTiledFn> fn{nullptr};
    EncodeTiledFn f = fn.load(std::memory_order_acquire);
    if (f == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            f = reinterpret_cast<EncodeTiledFn>(p);
            fn.store(f, std::memory_order_release);
        }
    }
    return f;
}

int make_tensor_map(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return DAS_ERR_CUDA;
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = (int)r;
        return DAS_ERR_CUDA;
    }
    return DAS_OK;
}

// Descriptor cache: a pool is scored batch after batch out of the same few logits buffers (the selectors recycle the T
// pass tensors through torch's caching allocator), so the (pointer, shape) -> CUtensorMap encoding is looked up
// instead of re-encoded 20 times per batch.  A descriptor only depends on its key: a stale entry can never be wrong.
const CUtensorMap* cached_tensor_map(das_handle* h, CUtensorMapDataType dtype, int rank, const void* base,
                                     const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                     CUtensorMapSwizzle swizzle, int* rc) {
    das_tmap_entry key;
    memset(&key, 0, sizeof(key));
    key.base = base;
    for (int i = 0; i < rank; ++i) key.dims[i] = dims[i], key.box[i] = box[i];
    for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
    key.rank = rank, key.dtype = (int)dtype, key.swizzle = (int)swizzle;
    for (int s = 0; s < das_handle::kTmapSlots; ++s) {
        const das_tmap_entry& e = h->tmaps[s];
        if (e.valid && e.base == key.base && e.rank == key.rank && e.dtype == key.dtype && e.swizzle == key.swizzle &&
            memcmp(e.dims, key.dims, sizeof(key.dims)) == 0 && memcmp(e.strides, key.strides, sizeof(key.strides)) == 0 &&
            memcmp(e.box, key.box, sizeof(key.box)) == 0) {
            ++h->tmap_hits;
            *rc = DAS_OK;
            return &e.map;
        }
    }
    das_tmap_entry& slot = h->tmaps[h->tmap_next];
    h->tmap_next = (h->tmap_next + 1) % das_handle::kTmapSlots;
    slot.valid = 0;
    *rc = make_tensor_map(&key.map, dtype, rank, base, dims, strides, box, swizzle);
    if (*rc != DAS_OK) return nullptr;
    key.valid = 1;
    slot = key;
    ++h->tmap_misses;
    return &slot.map;
}

L2Window::L2Window(das_handle* h, cudaStream_t stream, const void* base, size_t bytes) : st(stream) {
    if (h == nullptr || !h->opt[DAS_OPT_MC_L2_PERSIST] || base == nullptr || bytes == 0) return;
    if (h->l2_persist_max == 0 || h->l2_window_max == 0) return;
    if (h->l2_persist_set < h->l2_persist_max) {  // carve the persisting share of L2 out once per handle
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        h->l2_persist_set = h->l2_persist_max;
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    // the window covers the HEAD of the range, as much as the carve-out holds, with hitRatio 1: a state slightly larger
    // than the carve-out (B = 2: 83.9 MB vs 82.9 MB) keeps all but its tail resident; a hitRatio < 1 over the whole range
    // was measured worse (the persisting lines are then chosen at random per access and thrash)
    size_t win = bytes < h->l2_window_max ? bytes : h->l2_window_max;
    if (win > h->l2_persist_set) win = h->l2_persist_set;
    v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    active = true;
}

L2Window::~L2Window() {
    if (!active) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));  // num_bytes = 0 disables the window for whatever is enqueued next on this stream
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

}  // namespace das

using namespace das;

extern "C" {

const char* das_strerror(int status) {
    switch (status) {
        case DAS_OK: return "ok";
        case DAS_ERR_INVALID_ARG: return "invalid argument";
        case DAS_ERR_UNSUPPORTED: return "unsupported size (see DAS_MAX_* in das_b200.h)";
        case DAS_ERR_CUDA: return "CUDA runtime error (see das_last_cuda_error)";
        case DAS_ERR_MISALIGNED: return "pointer not aligned as documented";
        default: return "unknown das_status";
    }
}
int das_abi_version(void) { return DAS_ABI_VERSION; }
int das_last_cuda_error(void) { return g_last_cuda_error; }
uint64_t das_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int das_handle_create(int device, das_handle** out) {
    if (out == nullptr) return DAS_ERR_INVALID_ARG;
    *out = nullptr;
    if (device < 0) DAS_CUDA(cudaGetDevice(&device));
    int n_dev = 0;
    DAS_CUDA(cudaGetDeviceCount(&n_dev));
    if (device >= n_dev) return DAS_ERR_INVALID_ARG;
    int sms = 0, l2 = 0, persist = 0, window = 0, cc_major = 0;
    DAS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DAS_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device));
    DAS_CUDA(cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, device));
    DAS_CUDA(cudaDeviceGetAttribute(&window, cudaDevAttrMaxAccessPolicyWindowSize, device));
    DAS_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    if (cc_major != 10) return DAS_ERR_UNSUPPORTED;  // the library holds sm_100a code only
    das_handle* h = new (std::nothrow) das_handle;
    if (h == nullptr) return DAS_ERR_INVALID_ARG;
    memset(h, 0, sizeof(*h));
    h->magic = kDasHandleMagic;
    h->device = device;
    h->num_sms = sms;
    h->l2_bytes = (size_t)l2;
    h->l2_persist_max = (size_t)persist;
    h->l2_window_max = (size_t)window;
    h->opt[DAS_OPT_MC_TMA] = env_int("DAS_MC_TMA", 1) != 0;
    {
        const int v = env_int("DAS_MC_TMA_CTAS", 0);
        h->opt[DAS_OPT_MC_TMA_CTAS] = v >= 1 && v <= 8 ? v : 0;
    }
    {
        const int v = env_int("DAS_MC_UP_WARPS", 0);
        h->opt[DAS_OPT_MC_UP_WARPS] = up_warps_option_ok(v) ? v : 0;
    }
    h->opt[DAS_OPT_GEMM_2CTA] = env_int("DAS_GEMM_2CTA", 1) != 0;
    h->opt[DAS_OPT_KC_CLUSTER] = env_int("DAS_KC_CLUSTER", 1) != 0;
    h->opt[DAS_OPT_MC_L2_PERSIST] = env_int("DAS_MC_L2_PERSIST", 1) != 0;
    *out = h;
    return DAS_OK;
}

int das_handle_destroy(das_handle* h) {
    if (h == nullptr) return DAS_OK;
    DAS_ENTER(h);
    if (h->kc_step_table != nullptr) cudaFree(h->kc_step_table);
    h->magic = 0;
    delete h;
    return DAS_OK;
}

int das_handle_device(const das_handle* h) {
    return (h == nullptr || h->magic != kDasHandleMagic) ? DAS_ERR_INVALID_ARG : h->device;
}
int das_handle_sm_count(const das_handle* h) {
    return (h == nullptr || h->magic != kDasHandleMagic) ? DAS_ERR_INVALID_ARG : h->num_sms;
}

int das_handle_l2_info(const das_handle* h, size_t* l2_bytes, size_t* persisting_max, size_t* window_max) {
    if (h == nullptr || h->magic != kDasHandleMagic) return DAS_ERR_INVALID_ARG;
    if (l2_bytes != nullptr) *l2_bytes = h->l2_bytes;
    if (persisting_max != nullptr) *persisting_max = h->l2_persist_max;
    if (window_max != nullptr) *window_max = h->l2_window_max;
    return DAS_OK;
}

int das_handle_set_option(das_handle* h, int option, int value) {
    if (h == nullptr || h->magic != kDasHandleMagic || option < 0 || option >= DAS_OPT_COUNT) return DAS_ERR_INVALID_ARG;
    switch (option) {
        case DAS_OPT_MC_TMA_CTAS:
            if (value < 0 || value > 8) return DAS_ERR_INVALID_ARG;
            break;
        case DAS_OPT_MC_UP_WARPS:
            if (value != 0 && !up_warps_option_ok(value)) return DAS_ERR_INVALID_ARG;
            break;
        default:
            value = value != 0;
    }
    h->opt[option] = value;
    return DAS_OK;
}

int das_handle_get_option(const das_handle* h, int option, int* value) {
    if (h == nullptr || h->magic != kDasHandleMagic || option < 0 || option >= DAS_OPT_COUNT || value == nullptr)
        return DAS_ERR_INVALID_ARG;
    *value = h->opt[option];
    return DAS_OK;
}

}  // extern "C"
