// Instantiates the K1/K2 kernels for class counts DAS_C_LO..DAS_C_HI (compiled several times with
// different ranges so the 31 class counts build in parallel).
#include "mc_kernels.cuh"

#ifndef DAS_C_LO
#error "compile with -DDAS_C_LO=.. -DDAS_C_HI=.."
#endif

namespace das {

// Kernel configuration (measured with tools/k1_bench.cu on B200, profiles/r1_k1_variants.txt):
//   PROBS kernels, even H*W : VEC = 2 (64-bit loads).  C <= 20: accumulators in registers, 128 regs, 4 blocks/SM
//                             (best under the sustained power cap: 0.83 of the measured HBM peak, 0.92 in a
//                             burst); C > 20: accumulators in shared memory, 4-6 blocks/SM (0.80-0.82 / 0.95).
//                             VEC = 4 needs 160-255 registers and stays at 0.79 / 0.80-0.90.
//   PROBS kernels, odd  H*W : VEC = 1 (class planes are mutually misaligned), accumulators in registers
//   vote-only kernels       : VEC = 4 / 1, everything in registers (already at the copy roofline)
// static + dynamic shared memory can exceed the 48 KB default limit -> always opt in.
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    return DAS_OK;
}
// blocks/SM asked from ptxas: registers/thread ~ VEC*C + 42, 512 threads*regs <= 64K
constexpr int minb_vec2(int C) { return C <= 12 ? 8 : (C <= 21 ? 6 : (C <= 28 ? 5 : 4)); }
constexpr int minb_vec1(int C) { return C <= 16 ? 8 : (C <= 24 ? 6 : 4); }

template <int C>
int launch_accumulate(const McAccParams& p, int B, int vec, int flags, cudaStream_t st) {
    const long long per_block = (long long)kAccThreads * vec;
    dim3 grid((unsigned)((p.HW + per_block - 1) / per_block), (unsigned)B);
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
    const size_t smem2 = acc_smem_bytes(C, 2);
#define DAS_ACC(V, P, Q, S, MB, BYTES)                                                              \
    do {                                                                                            \
        int rc__ = set_smem(mc_accumulate_kernel<C, V, P, Q, S, MB>, BYTES);                        \
        if (rc__ != DAS_OK) return rc__;                                                            \
        DAS_LAUNCH((mc_accumulate_kernel<C, V, P, Q, S, MB>), grid, kAccThreads, BYTES, st, p);     \
    } while (0)
    if (probs) {
        if (vec == 2) {
            if constexpr (C <= 20) {
                if (votes) DAS_ACC(2, true, true, false, 4, 0);
                else DAS_ACC(2, true, false, false, 4, 0);
            } else {
                if (votes) DAS_ACC(2, true, true, true, minb_vec2(C), smem2);
                else DAS_ACC(2, true, false, true, minb_vec2(C), smem2);
            }
        } else if (vec == 1) {
            if (votes) DAS_ACC(1, true, true, false, minb_vec1(C), 0);
            else DAS_ACC(1, true, false, false, minb_vec1(C), 0);
        } else return DAS_ERR_INVALID_ARG;
    } else {
        if (vec == 4) DAS_ACC(4, false, true, false, 1, 0);
        else if (vec == 1) DAS_ACC(1, false, true, false, 1, 0);
        else return DAS_ERR_INVALID_ARG;
    }
#undef DAS_ACC
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

template <int C>
int launch_finalize(const McFinParams& p, int B, int vec, int flags, cudaStream_t st) {
    const bool vec4 = vec == 4;
    dim3 grid((unsigned)p.blocks_per_image, (unsigned)B);
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
#define DAS_FIN(V, P, Q) DAS_LAUNCH((mc_finalize_kernel<C, V, P, Q>), grid, kFinalizeThreads, 0, st, p)
    if (vec4) {
        if (probs && votes) DAS_FIN(4, true, true);
        else if (probs) DAS_FIN(4, true, false);
        else DAS_FIN(4, false, true);
    } else {
        if (probs && votes) DAS_FIN(1, true, true);
        else if (probs) DAS_FIN(1, true, false);
        else DAS_FIN(1, false, true);
    }
#undef DAS_FIN
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

template <int C>
int launch_score(const McScoreParams& p, int B, int vec, int flags, cudaStream_t st) {
    dim3 grid((unsigned)p.fin.blocks_per_image, (unsigned)B);
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
    const size_t smem2 = acc_smem_bytes(C, 2);
#define DAS_SCORE(V, P, Q, S, MB, BYTES)                                                            \
    do {                                                                                            \
        int rc__ = set_smem(mc_score_kernel<C, V, P, Q, S, MB>, BYTES);                             \
        if (rc__ != DAS_OK) return rc__;                                                            \
        DAS_LAUNCH((mc_score_kernel<C, V, P, Q, S, MB>), grid, kAccThreads, BYTES, st, p);          \
    } while (0)
    if (probs) {
        if (vec == 2) {
            if constexpr (C <= 20) {
                if (votes) DAS_SCORE(2, true, true, false, 4, 0);
                else DAS_SCORE(2, true, false, false, 4, 0);
            } else {
                if (votes) DAS_SCORE(2, true, true, true, minb_vec2(C), smem2);
                else DAS_SCORE(2, true, false, true, minb_vec2(C), smem2);
            }
        } else if (vec == 1) {
            if (votes) DAS_SCORE(1, true, true, false, minb_vec1(C), 0);
            else DAS_SCORE(1, true, false, false, minb_vec1(C), 0);
        } else return DAS_ERR_INVALID_ARG;
    } else {
        if (vec == 4) DAS_SCORE(4, false, true, false, 1, 0);
        else if (vec == 1) DAS_SCORE(1, false, true, false, 1, 0);
        else return DAS_ERR_INVALID_ARG;
    }
#undef DAS_SCORE
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

template <int C>
int launch_score_tma(const McTmaParams& p, int flags, int ctas_per_sm, cudaStream_t st) {
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
    // ring depth from the shared memory left per CTA: 227 KB per SM, 1 KB reserved per CTA, ~C/4 + 2 KB static
    const size_t stage = (size_t)C * kTmaPix * sizeof(float);
    int stages = 0;
    if (ctas_per_sm <= 0) ctas_per_sm = tma_ctas_per_sm(C);
    for (; ctas_per_sm >= 1; --ctas_per_sm) {
        const size_t per_cta = (size_t)227 * 1024 / ctas_per_sm - 1024 - ((size_t)C * 256 + 2048);
        stages = (int)(per_cta / stage);
        if (stages > kTmaMaxStages) stages = kTmaMaxStages;
        if (stages >= 2) break;
    }
    if (stages < 2) return DAS_ERR_UNSUPPORTED;
    McTmaParams q = p;
    q.stages = stages;
    const size_t smem = stage * stages;
    const int tiles = p.B * p.tiles_per_image;
    // persistent grid = what is actually co-resident (registers can allow fewer CTAs than shared memory does)
#define DAS_TMA(P, Q, A)                                                                           \
    do {                                                                                           \
        int rc__ = set_smem(mc_score_tma_kernel<C, P, Q, A>, smem);                                \
        if (rc__ != DAS_OK) return rc__;                                                           \
        int occ__ = 0;                                                                             \
        cudaError_t e__ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(                           \
            &occ__, mc_score_tma_kernel<C, P, Q, A>, kTmaThreads, smem);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__);                                             \
        if (occ__ < 1) return DAS_ERR_UNSUPPORTED;                                                 \
        if (occ__ > ctas_per_sm) occ__ = ctas_per_sm;                                              \
        const int grid__ = tiles < q.num_sms * occ__ ? tiles : q.num_sms * occ__;                      \
        DAS_LAUNCH((mc_score_tma_kernel<C, P, Q, A>), grid__, kTmaThreads, smem, st, q);           \
    } while (0)
    if (!p.flat) {
        if (probs && votes) DAS_TMA(true, true, false);
        else if (probs) DAS_TMA(true, false, false);
        else DAS_TMA(false, true, false);
    } else {
        if (probs && votes) DAS_TMA(true, true, true);
        else if (probs) DAS_TMA(true, false, true);
        else DAS_TMA(false, true, true);
    }
#undef DAS_TMA
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

// NW consumer warps per CTA (tile = 16 rows x 4 NW columns): 4 -> 3 CTAs/SM x 160 threads (2 above 24 classes),
// 15 -> 1 CTA x 512 threads; both leave 4 warps per scheduler, i.e. 128 registers per thread (a 17th warp would
// put 5 warps on one scheduler and cap the kernel at 96 registers).
template <int C, int NW, int MINB>
int launch_score_up_nw(const McUpParams& p, int flags, cudaStream_t st) {
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
    // ring depth: the windows are small, 4 stages keep the producer one round trip ahead of the consumers
    McUpParams q = p;
    q.stages = 4;
    const size_t smem = up_rows_bytes(C, NW) + up_wts_bytes(NW) + (size_t)q.stages * up_stage_bytes(C, NW);
    const int tiles = p.B * p.tiles_x * p.tiles_y;
#define DAS_UP(P, Q)                                                                               \
    do {                                                                                           \
        int rc__ = set_smem(mc_score_up_kernel<C, P, Q, NW, MINB>, smem);                          \
        if (rc__ != DAS_OK) return rc__;                                                           \
        int occ__ = 0;                                                                             \
        cudaError_t e__ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(                           \
            &occ__, mc_score_up_kernel<C, P, Q, NW, MINB>, up_threads(NW), smem);                  \
        if (e__ != cudaSuccess) return cuda_fail(e__);                                             \
        if (occ__ < 1) return DAS_ERR_UNSUPPORTED;                                                 \
        if (occ__ > MINB) occ__ = MINB;                                                            \
        const int grid__ = tiles < q.num_sms * occ__ ? tiles : q.num_sms * occ__;                      \
        DAS_LAUNCH((mc_score_up_kernel<C, P, Q, NW, MINB>), grid__, up_threads(NW), smem, st, q);  \
    } while (0)
    if (probs && votes) DAS_UP(true, true);
    else if (probs) DAS_UP(true, false);
    else DAS_UP(false, true);
#undef DAS_UP
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

// One pixel per lane, class pairs in the packed pipe (mc_up1.cuh): NW consumer warps per CTA, tile = 16 rows x 2 NW columns.
template <int C, int NW, int NP, int MINB>
int launch_score_up1_nw(const McUpParams& p, int flags, cudaStream_t st) {
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
    McUpParams q = p;
    q.stages = 4;
    const size_t smem = up1_rows_bytes(C, NW) + up1_wts_bytes(NW) + (size_t)q.stages * up1_stage_bytes(C, NW);
    const int tiles = p.B * p.tiles_x * p.tiles_y;
#define DAS_UP1(P, Q)                                                                              \
    do {                                                                                           \
        int rc__ = set_smem(mc_score_up1_kernel<C, P, Q, NW, NP, MINB>, smem);                         \
        if (rc__ != DAS_OK) return rc__;                                                           \
        int occ__ = 0;                                                                             \
        cudaError_t e__ = cudaOccupancyMaxActiveBlocksPerMultiprocessor(                           \
            &occ__, mc_score_up1_kernel<C, P, Q, NW, NP, MINB>, 32 * (NW + NP), smem);                 \
        if (e__ != cudaSuccess) return cuda_fail(e__);                                             \
        if (occ__ < 1) return DAS_ERR_UNSUPPORTED;                                                 \
        if (occ__ > MINB) occ__ = MINB;                                                            \
        const int grid__ = tiles < q.num_sms * occ__ ? tiles : q.num_sms * occ__;                  \
        DAS_LAUNCH((mc_score_up1_kernel<C, P, Q, NW, NP, MINB>), grid__, 32 * (NW + NP), smem, st, q); \
    } while (0)
    if (probs && votes) DAS_UP1(true, true);
    else if (probs) DAS_UP1(true, false);
    else DAS_UP1(false, true);
#undef DAS_UP1
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

// nw: 4 | 15 = pixel-pair kernel (mc_up.cuh); 220 | 216 = one-pixel-per-lane kernel (mc_up1.cuh) with 20 | 16 consumer + 4
// producer warps - the best of the splits measured in round 2 (profiles/r2_upsample_notes.md).  mc_api.cu: up_warps()
// picks per class count; DAS_OPT_MC_UP_WARPS forces one
template <int C>
int launch_score_up(const McUpParams& p, int flags, int nw, cudaStream_t st) {
    switch (nw) {
        case 4: return launch_score_up_nw<C, 4, (C <= 24 ? 3 : 2)>(p, flags, st);
        case 15: return launch_score_up_nw<C, 15, 1>(p, flags, st);
        case 220: return launch_score_up1_nw<C, 20, 4, 1>(p, flags, st);   // 20 + 4 warps, 80 registers
        case 216: return launch_score_up1_nw<C, 16, 4, 1>(p, flags, st);   // 16 + 4 warps, 96 registers
        default: return DAS_ERR_INVALID_ARG;
    }
}

template <int C>
struct Range {
    static int score_up(const McUpParams& p, int f, int ctas, cudaStream_t st) {
        if (p.fin.C == C) return launch_score_up<C>(p, f, ctas, st);
        return Range<C + 1>::score_up(p, f, ctas, st);
    }
    static int score_tma(const McTmaParams& p, int f, int ctas, cudaStream_t st) {
        if (p.fin.C == C) return launch_score_tma<C>(p, f, ctas, st);
        return Range<C + 1>::score_tma(p, f, ctas, st);
    }
    static int score(const McScoreParams& p, int B, int v4, int f, cudaStream_t st) {
        if (p.acc.C == C) return launch_score<C>(p, B, v4, f, st);
        return Range<C + 1>::score(p, B, v4, f, st);
    }
    static int acc(const McAccParams& p, int B, int v4, int f, cudaStream_t st) {
        if (p.C == C) return launch_accumulate<C>(p, B, v4, f, st);
        return Range<C + 1>::acc(p, B, v4, f, st);
    }
    static int fin(const McFinParams& p, int B, int v4, int f, cudaStream_t st) {
        if (p.C == C) return launch_finalize<C>(p, B, v4, f, st);
        return Range<C + 1>::fin(p, B, v4, f, st);
    }
};
template <>
struct Range<DAS_C_HI + 1> {
    static int acc(const McAccParams&, int, int, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
    static int fin(const McFinParams&, int, int, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
    static int score(const McScoreParams&, int, int, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
    static int score_tma(const McTmaParams&, int, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
    static int score_up(const McUpParams&, int, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
};

#define DAS_CAT_(a, b, c) a##b##_##c
#define DAS_CAT(a, b, c) DAS_CAT_(a, b, c)
int DAS_CAT(dispatch_accumulate_, DAS_C_LO, DAS_C_HI)(const McAccParams& p, int B, int v4, int f, cudaStream_t st) {
    return Range<DAS_C_LO>::acc(p, B, v4, f, st);
}
int DAS_CAT(dispatch_finalize_, DAS_C_LO, DAS_C_HI)(const McFinParams& p, int B, int v4, int f, cudaStream_t st) {
    return Range<DAS_C_LO>::fin(p, B, v4, f, st);
}

int DAS_CAT(dispatch_score_, DAS_C_LO, DAS_C_HI)(const McScoreParams& p, int B, int v4, int f, cudaStream_t st) {
    return Range<DAS_C_LO>::score(p, B, v4, f, st);
}
int DAS_CAT(dispatch_score_tma_, DAS_C_LO, DAS_C_HI)(const McTmaParams& p, int f, int ctas, cudaStream_t st) {
    return Range<DAS_C_LO>::score_tma(p, f, ctas, st);
}
int DAS_CAT(dispatch_score_up_, DAS_C_LO, DAS_C_HI)(const McUpParams& p, int f, int ctas, cudaStream_t st) {
    return Range<DAS_C_LO>::score_up(p, f, ctas, st);
}

}  // namespace das
