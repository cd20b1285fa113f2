// extern "C" entry points of the Monte-Carlo reduction (K1/K2): validation, state layout, dispatch.
#include <math.h>

#include <stdlib.h>

#include "mc_kernels.cuh"
#include "tma_host.cuh"

namespace das {

// class-count ranges compiled in separate translation units (see build.py)
#define DAS_DECL_RANGE(LO, HI)                                                                         \
    int dispatch_accumulate_##LO##_##HI(const McAccParams&, int, int, int, cudaStream_t);             \
    int dispatch_finalize_##LO##_##HI(const McFinParams&, int, int, int, cudaStream_t);               \
    int dispatch_score_##LO##_##HI(const McScoreParams&, int, int, int, cudaStream_t);                \
    int dispatch_score_tma_##LO##_##HI(const McTmaParams&, int, int, cudaStream_t);                   \
    int dispatch_score_up_##LO##_##HI(const McUpParams&, int, int, cudaStream_t);
DAS_DECL_RANGE(2, 9)
DAS_DECL_RANGE(10, 16)
DAS_DECL_RANGE(17, 20)
DAS_DECL_RANGE(21, 24)
DAS_DECL_RANGE(25, 28)
DAS_DECL_RANGE(29, 32)

int dispatch_accumulate(const McAccParams& p, int B, int v4, int f, cudaStream_t st) {
    if (p.C <= 9) return dispatch_accumulate_2_9(p, B, v4, f, st);
    if (p.C <= 16) return dispatch_accumulate_10_16(p, B, v4, f, st);
    if (p.C <= 20) return dispatch_accumulate_17_20(p, B, v4, f, st);
    if (p.C <= 24) return dispatch_accumulate_21_24(p, B, v4, f, st);
    if (p.C <= 28) return dispatch_accumulate_25_28(p, B, v4, f, st);
    return dispatch_accumulate_29_32(p, B, v4, f, st);
}
int dispatch_finalize(const McFinParams& p, int B, int v4, int f, cudaStream_t st) {
    if (p.C <= 9) return dispatch_finalize_2_9(p, B, v4, f, st);
    if (p.C <= 16) return dispatch_finalize_10_16(p, B, v4, f, st);
    if (p.C <= 20) return dispatch_finalize_17_20(p, B, v4, f, st);
    if (p.C <= 24) return dispatch_finalize_21_24(p, B, v4, f, st);
    if (p.C <= 28) return dispatch_finalize_25_28(p, B, v4, f, st);
    return dispatch_finalize_29_32(p, B, v4, f, st);
}

int dispatch_score(const McScoreParams& p, int B, int v4, int f, cudaStream_t st) {
    const int C = p.acc.C;
    if (C <= 9) return dispatch_score_2_9(p, B, v4, f, st);
    if (C <= 16) return dispatch_score_10_16(p, B, v4, f, st);
    if (C <= 20) return dispatch_score_17_20(p, B, v4, f, st);
    if (C <= 24) return dispatch_score_21_24(p, B, v4, f, st);
    if (C <= 28) return dispatch_score_25_28(p, B, v4, f, st);
    return dispatch_score_29_32(p, B, v4, f, st);
}

int dispatch_score_tma(const McTmaParams& p, int f, int ctas, cudaStream_t st) {
    const int C = p.fin.C;
    if (C <= 9) return dispatch_score_tma_2_9(p, f, ctas, st);
    if (C <= 16) return dispatch_score_tma_10_16(p, f, ctas, st);
    if (C <= 20) return dispatch_score_tma_17_20(p, f, ctas, st);
    if (C <= 24) return dispatch_score_tma_21_24(p, f, ctas, st);
    if (C <= 28) return dispatch_score_tma_25_28(p, f, ctas, st);
    return dispatch_score_tma_29_32(p, f, ctas, st);
}

int dispatch_score_up(const McUpParams& p, int f, int ctas, cudaStream_t st) {
    const int C = p.fin.C;
    if (C <= 9) return dispatch_score_up_2_9(p, f, ctas, st);
    if (C <= 16) return dispatch_score_up_10_16(p, f, ctas, st);
    if (C <= 20) return dispatch_score_up_17_20(p, f, ctas, st);
    if (C <= 24) return dispatch_score_up_21_24(p, f, ctas, st);
    if (C <= 28) return dispatch_score_up_25_28(p, f, ctas, st);
    return dispatch_score_up_29_32(p, f, ctas, st);
}

int mc_validate(const das_mc_desc* d) {
    if (d == nullptr) return DAS_ERR_INVALID_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->T_cap <= 0) return DAS_ERR_INVALID_ARG;
    if ((d->flags & (DAS_MC_VOTES | DAS_MC_PROBS)) == 0 ||
        (d->flags & ~(DAS_MC_VOTES | DAS_MC_PROBS | DAS_MC_SINGLE_SHOT)))
        return DAS_ERR_INVALID_ARG;
    if ((d->flags & DAS_MC_SINGLE_SHOT) && d->T_cap > DAS_MAX_PASS_GROUP) return DAS_ERR_UNSUPPORTED;
    if (d->C < 2) return DAS_ERR_INVALID_ARG;
    if (d->C > DAS_MAX_CLASSES || d->T_cap > DAS_MAX_PASSES || d->B > 65535) return DAS_ERR_UNSUPPORTED;
    // plane offsets inside one image are formed in 32 bits (c * H*W*4 bytes)
    if ((unsigned long long)d->C * d->H * d->W * sizeof(float) >= (1ull << 32)) return DAS_ERR_UNSUPPORTED;
    return DAS_OK;
}

// pixels per thread: K2 and the vote-only kernels use 128-bit accesses when H*W % 4 == 0; the kernels that
// carry softmax accumulators use 64-bit accesses when H*W is even (see mc_inst.cu); otherwise scalar
static int fin_vec(const das_mc_desc& d) { return ((long long)d.H * d.W) % 4 == 0 ? 4 : 1; }
static int acc_vec(const das_mc_desc& d) {
    const long long HW = (long long)d.H * d.W;
    if (d.flags & DAS_MC_PROBS) return HW % 2 == 0 ? 2 : 1;
    return HW % 4 == 0 ? 4 : 1;
}
// pointer alignment every float tensor of this batch must have (bytes): 16 / 8 / 4 by H*W % 4, % 2
static uintptr_t float_align(const das_mc_desc& d) {
    const long long HW = (long long)d.H * d.W;
    return HW % 4 == 0 ? 16 : (HW % 2 == 0 ? 8 : 4);
}
static bool misaligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

// bytes of the fp32 accumulators at the head of the state (sum_p | sum_ent): what the streaming form round-trips per pass
static size_t state_acc_bytes(const das_mc_desc& d) {
    if (!(d.flags & DAS_MC_PROBS) || (d.flags & DAS_MC_SINGLE_SHOT)) return 0;
    const size_t HW = (size_t)d.H * d.W;
    return align_up((size_t)d.B * d.C * HW * sizeof(float), 256) + align_up((size_t)d.B * HW * sizeof(float), 256);
}

McLayout mc_layout(const das_mc_desc& d) {
    McLayout L;
    const size_t HW = (size_t)d.H * d.W;
    const int vec = fin_vec(d), avec = acc_vec(d);
    L.blocks_per_image = (int)((HW + (size_t)kFinalizeThreads * vec - 1) / ((size_t)kFinalizeThreads * vec));
    L.blocks_fused = (int)((HW + (size_t)kAccThreads * avec - 1) / ((size_t)kAccThreads * avec));
    const bool keep = !(d.flags & DAS_MC_SINGLE_SHOT);
    size_t off = 0;
    L.sum_p = off;
    if (keep && (d.flags & DAS_MC_PROBS)) off += align_up((size_t)d.B * d.C * HW * sizeof(float), 256);
    L.sum_ent = off;
    if (keep && (d.flags & DAS_MC_PROBS)) off += align_up((size_t)d.B * HW * sizeof(float), 256);
    L.votes = off;
    if (keep && (d.flags & DAS_MC_VOTES)) off += align_up((size_t)d.B * d.T_cap * HW, 256);
    L.partials = off;
    // sized for the finest block partition any kernel uses (the TMA kernel: 256-pixel tiles)
    const int blocks_tma = (int)((HW + kTmaFlatPix - 1) / kTmaFlatPix);
    // ... and the fused-upsample kernel: 16 x 16 tiles
    const int blocks_up = ((d.H + kUpTileH - 1) / kUpTileH) * ((d.W + up_tile_w(4) - 1) / up_tile_w(4));
    int blocks_max = L.blocks_fused > blocks_tma ? L.blocks_fused : blocks_tma;
    if (blocks_up > blocks_max) blocks_max = blocks_up;
    off += align_up((size_t)d.B * blocks_max * DAS_N_SCORES * sizeof(float), 256);
    L.total = off;
    return L;
}

// one warp per (image, score): fixed-order fp64 sum of the block partials -> mean over H*W
__global__ void mc_reduce_partials_kernel(const float* partials, int blocks_per_image, long long HW, int flags,
                                          float* image_scores) {
    const int b = blockIdx.x;
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= DAS_N_SCORES) return;
    double s = 0.0;
    for (int i = lane; i < blocks_per_image; i += 32)
        s += (double)partials[((size_t)b * blocks_per_image + i) * DAS_N_SCORES + k];
    s = warp_sum(s);
    if (lane == 0) {
        const bool have = (k == DAS_SCORE_VOTE_ENTROPY) ? (flags & DAS_MC_VOTES) : (flags & DAS_MC_PROBS);
        image_scores[(size_t)b * DAS_N_SCORES + k] = have ? (float)(s / (double)HW) : __int_as_float(0x7fc00000);
    }
}

}  // namespace das

using namespace das;

static int fill_acc_params(const das_mc_desc* desc, void* state, const float* const* pass_logits, int n_passes,
                           int pass_begin, McAccParams* out) {
    if (state == nullptr || pass_logits == nullptr) return DAS_ERR_INVALID_ARG;
    if (n_passes < 1 || pass_begin < 0 || pass_begin + n_passes > desc->T_cap) return DAS_ERR_INVALID_ARG;
    if (n_passes > DAS_MAX_PASS_GROUP) return DAS_ERR_UNSUPPORTED;
    const uintptr_t al = float_align(*desc);
    if (!aligned16(state)) return DAS_ERR_MISALIGNED;
    McAccParams& p = *out;
    for (int g = 0; g < n_passes; ++g) {
        if (pass_logits[g] == nullptr) return DAS_ERR_INVALID_ARG;
        if (misaligned(pass_logits[g], al)) return DAS_ERR_MISALIGNED;
        p.logits[g] = pass_logits[g];
    }
    for (int g = n_passes; g < DAS_MAX_PASS_GROUP; ++g) p.logits[g] = nullptr;
    const McLayout L = mc_layout(*desc);
    char* base = static_cast<char*>(state);
    p.sum_p = reinterpret_cast<float*>(base + L.sum_p);
    p.sum_ent = reinterpret_cast<float*>(base + L.sum_ent);
    p.votes = (desc->flags & DAS_MC_SINGLE_SHOT) || !(desc->flags & DAS_MC_VOTES)
                  ? nullptr
                  : reinterpret_cast<uint8_t*>(base + L.votes);
    p.HW = (long long)desc->H * desc->W;
    p.C = desc->C;
    p.T_cap = desc->T_cap;
    p.n_passes = n_passes;
    p.pass_begin = pass_begin;
    return DAS_OK;
}

static int fill_fin_params(const das_mc_desc* desc, void* state, const float* labels, int T, float* vote_entropy,
                           float* pred_entropy, float* bald, float* confidence, float* margin, uint8_t* weak_labels,
                           int blocks, McFinParams* out) {
    if (state == nullptr || T < 1 || T > desc->T_cap) return DAS_ERR_INVALID_ARG;
    const bool probs = desc->flags & DAS_MC_PROBS, votes = desc->flags & DAS_MC_VOTES;
    if (!votes && (vote_entropy || weak_labels)) return DAS_ERR_INVALID_ARG;
    if (!probs && (pred_entropy || bald || confidence || margin)) return DAS_ERR_INVALID_ARG;
    {
        const uintptr_t al = float_align(*desc);
        const void* ptrs[] = {labels, vote_entropy, pred_entropy, bald, confidence, margin};
        for (const void* q : ptrs)
            if (q != nullptr && misaligned(q, al)) return DAS_ERR_MISALIGNED;
        if (!aligned16(state) || (weak_labels != nullptr && misaligned(weak_labels, al / 4))) return DAS_ERR_MISALIGNED;
    }
    const McLayout L = mc_layout(*desc);
    char* base = static_cast<char*>(state);
    McFinParams& p = *out;
    p.sum_p = reinterpret_cast<const float*>(base + L.sum_p);
    p.sum_ent = reinterpret_cast<const float*>(base + L.sum_ent);
    p.votes = reinterpret_cast<const uint8_t*>(base + L.votes);
    p.labels = labels;
    p.vote_entropy = vote_entropy;
    p.pred_entropy = pred_entropy;
    p.bald = bald;
    p.confidence = confidence;
    p.margin = margin;
    p.weak_labels = weak_labels;
    p.partials = reinterpret_cast<float*>(base + L.partials);
    p.HW = (long long)desc->H * desc->W;
    p.C = desc->C;
    p.T_cap = desc->T_cap;
    p.T = T;
    p.blocks_per_image = blocks;
    return DAS_OK;
}

static int reduce_partials(const das_mc_desc* desc, const McFinParams& p, float* image_scores, cudaStream_t st) {
    if (image_scores == nullptr) return DAS_OK;
    DAS_LAUNCH(mc_reduce_partials_kernel, desc->B, 32 * DAS_N_SCORES, 0, st, p.partials, p.blocks_per_image, p.HW,
               desc->flags, image_scores);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

// ---- TMA-staged single-shot path -------------------------------------------------------------
// DAS_OPT_MC_TMA = 0 forces the LDG kernel (A/B measurements, tests of both paths); DAS_OPT_MC_TMA_CTAS = CTAs per SM
static bool tma_eligible(const das_handle* h, const das_mc_desc* desc, const McScoreParams& q) {
    if (!h->opt[DAS_OPT_MC_TMA] || q.acc.pass_begin != 0) return false;
    // vote-only scoring has no softmax work to overlap: its LDG kernel already runs at the copy roofline
    // (measured 1.00 vs 0.98 for the ring on 513 x 513 planes)
    if (!(desc->flags & DAS_MC_PROBS)) return false;
    const long long HW = (long long)desc->H * desc->W;
    if (HW < kTmaPix) return false;
    // odd planes go through the flat 1-D maps: element coordinates are 32 bit
    if (HW % 4 != 0 && (long long)desc->B * desc->C * HW >= (1ll << 31) - kTmaPix) return false;
    for (int g = 0; g < q.acc.n_passes; ++g)
        if (!aligned16(q.acc.logits[g])) return false;
    return true;
}
static int fill_tma_params(das_handle* h, const das_mc_desc* desc, const McScoreParams& q, McTmaParams* out) {
    const unsigned long long HW = (unsigned long long)desc->H * desc->W;
    const bool flat = HW % 4 != 0;  // plane strides are not multiples of 16 bytes: address by element instead
    for (int g = 0; g < q.acc.n_passes; ++g) {
        int rc = DAS_OK;
        const CUtensorMap* m;   // descriptors are cached per (buffer, shape) in the handle
        if (!flat) {
            const cuuint64_t dims[3] = {HW, (cuuint64_t)desc->C, (cuuint64_t)desc->B};
            const cuuint64_t strides[2] = {HW * sizeof(float), HW * desc->C * sizeof(float)};
            const cuuint32_t box[3] = {(cuuint32_t)kTmaPix, (cuuint32_t)desc->C, 1};
            m = cached_tensor_map(h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, q.acc.logits[g], dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, &rc);
        } else {
            const cuuint64_t dims[1] = {HW * desc->C * desc->B};
            const cuuint64_t strides[1] = {0};
            const cuuint32_t box[1] = {(cuuint32_t)kTmaPix};
            m = cached_tensor_map(h, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, q.acc.logits[g], dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, &rc);
        }
        if (m == nullptr) return rc != DAS_OK ? rc : DAS_ERR_CUDA;
        out->maps[g] = *m;
    }
    out->fin = q.fin;
    out->HW = (long long)HW;
    out->B = desc->B;
    out->n_passes = q.acc.n_passes;
    out->tiles_per_image = q.fin.blocks_per_image;  // 256-pixel tiles (the VEC=2 LDG kernel's partition) or 252 (flat)
    out->stages = 0;
    out->flat = flat ? 1 : 0;
    out->num_sms = h->num_sms;
    return DAS_OK;
}

// ---- fused bilinear upsample (low-resolution logits) --------------------------------------------

// ATen's align_corners scale (area_pixel_compute_scale<float>)
static float up_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f; }
// does every `tile`-pixel tile of the output axis read at most `window` consecutive source samples, and does a
// 4-pixel strip (tile % 4 == 0) start at most one source sample after its first pixel's?  (the kernel's arithmetic,
// replayed on the host: IEEE float multiply, truncation)
static bool up_window_fits(int n_in, int n_out, float scale, int tile, int window, int strip) {
    auto src = [&](int d) {
        int i = (int)(scale * (float)d);
        return i > n_in - 1 ? n_in - 1 : i;
    };
    for (int t0 = 0; t0 < n_out; t0 += tile) {
        const int last = (t0 + tile - 1 < n_out ? t0 + tile - 1 : n_out - 1);
        const int i0 = src(last);
        const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
        if (i1 - src(t0) > window - 1) return false;
    }
    for (int s0 = 0; s0 < n_out; s0 += strip) {
        const int last = (s0 + strip - 1 < n_out ? s0 + strip - 1 : n_out - 1);
        if (src(last) - src(s0) > 1) return false;
    }
    return true;
}
// kernel variant code `nw`: 4 | 15 = pixel-pair kernel with that many consumer warps (mc_up.cuh, strips of 4 columns);
// 200 + NW = one-pixel-per-lane kernel (mc_up1.cuh, strips of 2 columns, NW consumer warps; 220 is built)
static bool up_is_v1(int nw) { return nw >= 100; }
static int up_variant_tile_w(int nw) { return up_is_v1(nw) ? up1_tile_w(nw % 100) : up_tile_w(nw); }
static int up_variant_win_cols(int nw) { return up_is_v1(nw) ? up1_win_cols(nw % 100) : up_win_cols(nw); }
static bool up_variant_known(int nw) { return nw == 4 || nw == 15 || nw == 220 || nw == 216; }
static bool up_supported(int h, int w, int H, int W, int nw) {
    if (h < 1 || w < 1 || H < 1 || W < 1) return false;
    return up_window_fits(h, H, up_scale(h, H), kUpTileH, kUpRows, kUpStrip) &&
           up_window_fits(w, W, up_scale(w, W), up_variant_tile_w(nw), up_variant_win_cols(nw),
                          up_is_v1(nw) ? kUp1Strip : kUpStrip);
}
// Kernel variant for a shape (0 = not supported).  Measured on the B200 (profiles/r2_upsample_notes.md, 512 x 1024 and
// 513 x 513 outputs, B = 8, T = 20):
//   C <= 20  one pixel per lane, 20 consumer + 4 producer warps at 80 registers (220): 0.934 ms against 0.974 ms for the
//            pixel-pair kernel at C = 19 (0.490 / 0.520 ms on the Pascal shape)
//   C == 21  pixel pairs, 15 + 1 warps at 128 registers (15): 0.529 against 0.559 ms (Pascal shape), 1.026 / 1.072 ms
//   C >= 22  one pixel per lane, 16 + 4 warps at 96 registers (216): the pixel-pair kernel spills from here on
//            (C = 24: 1.17 against 1.22 ms, C = 28: 1.32 / 1.66, C = 32: 1.50 / 2.14)
// Narrow outputs (less than two of the wide tiles per row) keep the pixel-pair kernel with three 160-thread CTAs per SM
// and 16 x 16 tiles (4).  B and C are 0 when only the shape is asked about (das_mc_upsample_supported).
// Vote-only scoring (no softmax, no accumulators: nothing to gain from fewer registers) stays with the pixel-pair kernel:
// 0.788 against 0.838 ms at C = 19, 0.996 / 1.023 at C = 24 (its producer, fully unrolled, is no longer the critical path:
// rolled, vote-only scoring took 0.976 ms - as long as the full pass).
// DAS_OPT_MC_UP_WARPS overrides the choice.
static int up_warps(const das_handle* hd, int h, int w, int H, int W, int B = 0, int C = 0, int flags = DAS_MC_PROBS) {
    const int v = hd != nullptr ? hd->opt[DAS_OPT_MC_UP_WARPS] : 0;
    if (v != 0 && up_variant_known(v)) return up_supported(h, w, H, W, v) ? v : 0;
    const int v1 = C == 0 || C == 21 || !(flags & DAS_MC_PROBS) ? 0 : (C <= 20 ? 220 : 216);
    // the one-pixel-per-lane kernel forms source offsets inside the whole BATCH in 32 bits
    if (v1 != 0 && W >= 2 * up_variant_tile_w(v1) && (unsigned long long)B * C * h * w < (1ull << 32) &&
        up_supported(h, w, H, W, v1))
        return v1;
    if (W >= 2 * up_tile_w(15) && up_supported(h, w, H, W, 15)) return 15;
    return up_supported(h, w, H, W, 4) ? 4 : 0;
}

extern "C" {

int das_mc_state_bytes(const das_mc_desc* desc, size_t* bytes) {
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (bytes == nullptr) return DAS_ERR_INVALID_ARG;
    *bytes = mc_layout(*desc).total;
    return DAS_OK;
}

int das_mc_reset(das_handle* h, const das_mc_desc* desc, void* state, void* stream) {
    DAS_ENTER(h);
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (state == nullptr) return DAS_ERR_INVALID_ARG;
    DAS_CUDA(cudaMemsetAsync(state, 0, mc_layout(*desc).total, (cudaStream_t)stream));
    return DAS_OK;
}

int das_mc_accumulate(das_handle* h, const das_mc_desc* desc, void* state, const float* const* pass_logits,
                      int n_passes, int pass_begin, void* stream) {
    DAS_ENTER(h);
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (desc->flags & DAS_MC_SINGLE_SHOT) return DAS_ERR_INVALID_ARG;
    McAccParams p;
    rc = fill_acc_params(desc, state, pass_logits, n_passes, pass_begin, &p);
    if (rc != DAS_OK) return rc;
    const L2Window keep(h, (cudaStream_t)stream, (desc->flags & DAS_MC_PROBS) ? state : nullptr, state_acc_bytes(*desc));
    return dispatch_accumulate(p, desc->B, acc_vec(*desc), desc->flags & (DAS_MC_VOTES | DAS_MC_PROBS),
                               (cudaStream_t)stream);
}

int das_mc_finalize(das_handle* h, const das_mc_desc* desc, void* state, const float* labels, int T,
                    float* vote_entropy, float* pred_entropy, float* bald, float* confidence, float* margin,
                    uint8_t* weak_labels, float* image_scores, void* stream) {
    DAS_ENTER(h);
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (desc->flags & DAS_MC_SINGLE_SHOT) return DAS_ERR_INVALID_ARG;
    McFinParams p;
    rc = fill_fin_params(desc, state, labels, T, vote_entropy, pred_entropy, bald, confidence, margin, weak_labels,
                         mc_layout(*desc).blocks_per_image, &p);
    if (rc != DAS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const L2Window keep(h, st, (desc->flags & DAS_MC_PROBS) ? state : nullptr, state_acc_bytes(*desc));
    rc = dispatch_finalize(p, desc->B, fin_vec(*desc), desc->flags & (DAS_MC_VOTES | DAS_MC_PROBS), st);
    if (rc != DAS_OK) return rc;
    return reduce_partials(desc, p, image_scores, st);
}

int das_mc_accumulate_finalize(das_handle* h, const das_mc_desc* desc, void* state, const float* const* pass_logits,
                               int n_passes, int pass_begin, const float* labels, float* vote_entropy,
                               float* pred_entropy, float* bald, float* confidence, float* margin,
                               uint8_t* weak_labels, float* image_scores, void* stream) {
    DAS_ENTER(h);
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if ((desc->flags & DAS_MC_SINGLE_SHOT) && pass_begin != 0) return DAS_ERR_INVALID_ARG;
    McScoreParams q;
    rc = fill_acc_params(desc, state, pass_logits, n_passes, pass_begin, &q.acc);
    if (rc != DAS_OK) return rc;
    rc = fill_fin_params(desc, state, labels, pass_begin + n_passes, vote_entropy, pred_entropy, bald, confidence,
                         margin, weak_labels, mc_layout(*desc).blocks_fused, &q.fin);
    if (rc != DAS_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int flags = desc->flags & (DAS_MC_VOTES | DAS_MC_PROBS);
    if (tma_eligible(h, desc, q)) {
        // whole Monte-Carlo stack in one launch and 16-byte aligned planes: TMA-staged persistent kernel
        {
            const long long HW = (long long)desc->H * desc->W;
            const int tile = HW % 4 == 0 ? kTmaPix : kTmaFlatPix;
            q.fin.blocks_per_image = (int)((HW + tile - 1) / tile);
        }
        McTmaParams tp;
        rc = fill_tma_params(h, desc, q, &tp);
        if (rc != DAS_OK) return rc;
        rc = dispatch_score_tma(tp, flags, h->opt[DAS_OPT_MC_TMA_CTAS], st);
    } else {
        // the last group of a streamed batch reads the running state: keep it in L2 like das_mc_accumulate does
        const L2Window keep(h, st, (pass_begin > 0 && (desc->flags & DAS_MC_PROBS)) ? state : nullptr, state_acc_bytes(*desc));
        rc = dispatch_score(q, desc->B, acc_vec(*desc), flags, st);
    }
    if (rc != DAS_OK) return rc;
    return reduce_partials(desc, q.fin, image_scores, st);
}

int das_mc_upsample_supported(const das_handle* hd, int h, int w, int H, int W) {
    if (hd != nullptr && hd->magic != kDasHandleMagic) return 0;
    return up_warps(hd, h, w, H, W) != 0 ? 1 : 0;  // hd == NULL: the default options (host-only question)
}

int das_mc_upsample_variant(const das_handle* hd, const das_mc_desc* desc, int h, int w) {
    if ((hd != nullptr && hd->magic != kDasHandleMagic) || desc == nullptr) return 0;
    if (mc_validate(desc) != DAS_OK || (unsigned long long)desc->C * h * w >= (1ull << 31)) return 0;
    const int nw = up_warps(hd, h, w, desc->H, desc->W, desc->B, desc->C, desc->flags);
    if (up_is_v1(nw) && (unsigned long long)desc->B * desc->C * h * w >= (1ull << 32)) return 0;  // forced by the option
    return nw;
}

int das_mc_upsample_accumulate_finalize(das_handle* hd, const das_mc_desc* desc, void* state,
                                        const float* const* pass_lowres_logits, int n_passes, int h, int w,
                                        const float* labels, float* vote_entropy, float* pred_entropy, float* bald,
                                        float* confidence, float* margin, uint8_t* weak_labels, float* image_scores,
                                        void* stream) {
    DAS_ENTER(hd);
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (pass_lowres_logits == nullptr || h < 1 || w < 1) return DAS_ERR_INVALID_ARG;
    if (n_passes < 1 || n_passes > desc->T_cap) return DAS_ERR_INVALID_ARG;
    if (n_passes > DAS_MAX_PASS_GROUP) return DAS_ERR_UNSUPPORTED;
    const int nw = up_warps(hd, h, w, desc->H, desc->W, desc->B, desc->C, desc->flags);
    if (nw == 0) return DAS_ERR_UNSUPPORTED;
    // the one-pixel-per-lane kernel forms source offsets inside the whole BATCH in 32 bits
    if (up_is_v1(nw) && (unsigned long long)desc->B * desc->C * h * w >= (1ull << 32)) return DAS_ERR_UNSUPPORTED;
    // source offsets inside one image are formed in 32 bits
    if ((unsigned long long)desc->C * h * w >= (1ull << 31)) return DAS_ERR_UNSUPPORTED;
    McUpParams q;
    const int tiles_x = (desc->W + up_variant_tile_w(nw) - 1) / up_variant_tile_w(nw), tiles_y = (desc->H + kUpTileH - 1) / kUpTileH;
    rc = fill_fin_params(desc, state, labels, n_passes, vote_entropy, pred_entropy, bald, confidence, margin,
                         weak_labels, tiles_x * tiles_y, &q.fin);
    if (rc != DAS_OK) return rc;
    for (int g = 0; g < DAS_MAX_PASS_GROUP; ++g) {
        q.lowres[g] = g < n_passes ? pass_lowres_logits[g] : nullptr;
        if (g < n_passes && (q.lowres[g] == nullptr)) return DAS_ERR_INVALID_ARG;
        if (g < n_passes && misaligned(q.lowres[g], 4)) return DAS_ERR_MISALIGNED;
    }
    q.B = desc->B;
    q.n_passes = n_passes;
    q.stages = 0;
    q.h = h, q.w = w, q.H = desc->H, q.W = desc->W;
    q.tiles_x = tiles_x, q.tiles_y = tiles_y;
    q.rh = up_scale(h, desc->H), q.rw = up_scale(w, desc->W);
    q.num_sms = hd->num_sms;
    cudaStream_t st = (cudaStream_t)stream;
    rc = dispatch_score_up(q, desc->flags & (DAS_MC_VOTES | DAS_MC_PROBS), nw, st);
    if (rc != DAS_OK) return rc;
    return reduce_partials(desc, q.fin, image_scores, st);
}

int das_mc_votes_ptr(const das_mc_desc* desc, void* state, uint8_t** votes) {
    int rc = mc_validate(desc);
    if (rc != DAS_OK) return rc;
    if (state == nullptr || votes == nullptr || !(desc->flags & DAS_MC_VOTES) || (desc->flags & DAS_MC_SINGLE_SHOT))
        return DAS_ERR_INVALID_ARG;
    *votes = reinterpret_cast<uint8_t*>(static_cast<char*>(state) + mc_layout(*desc).votes);
    return DAS_OK;
}

}  // extern "C"
