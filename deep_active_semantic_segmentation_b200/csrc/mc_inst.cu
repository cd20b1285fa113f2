// Instantiates the K1/K2 kernels for class counts DAS_C_LO..DAS_C_HI (compiled several times with
// different ranges so the 31 class counts build in parallel).
#include "mc_kernels.cuh"

#ifndef DAS_C_LO
#error "compile with -DDAS_C_LO=.. -DDAS_C_HI=.."
#endif

namespace das {

template <int C>
int launch_accumulate(const McAccParams& p, int B, bool vec4, int flags, cudaStream_t st) {
    const int vec = vec4 ? 4 : 1;
    const long long per_block = (long long)kAccThreads * vec;
    dim3 grid((unsigned)((p.HW + per_block - 1) / per_block), (unsigned)B);
    const bool probs = flags & DAS_MC_PROBS, votes = flags & DAS_MC_VOTES;
#define DAS_ACC(V, P, Q) DAS_LAUNCH((mc_accumulate_kernel<C, V, P, Q>), grid, kAccThreads, 0, st, p)
    if (vec4) {
        if (probs && votes) DAS_ACC(4, true, true);
        else if (probs) DAS_ACC(4, true, false);
        else DAS_ACC(4, false, true);
    } else {
        if (probs && votes) DAS_ACC(1, true, true);
        else if (probs) DAS_ACC(1, true, false);
        else DAS_ACC(1, false, true);
    }
#undef DAS_ACC
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

template <int C>
int launch_finalize(const McFinParams& p, int B, bool vec4, int flags, cudaStream_t st) {
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
struct Range {
    static int acc(const McAccParams& p, int B, bool v4, int f, cudaStream_t st) {
        if (p.C == C) return launch_accumulate<C>(p, B, v4, f, st);
        return Range<C + 1>::acc(p, B, v4, f, st);
    }
    static int fin(const McFinParams& p, int B, bool v4, int f, cudaStream_t st) {
        if (p.C == C) return launch_finalize<C>(p, B, v4, f, st);
        return Range<C + 1>::fin(p, B, v4, f, st);
    }
};
template <>
struct Range<DAS_C_HI + 1> {
    static int acc(const McAccParams&, int, bool, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
    static int fin(const McFinParams&, int, bool, int, cudaStream_t) { return DAS_ERR_UNSUPPORTED; }
};

#define DAS_CAT_(a, b, c) a##b##_##c
#define DAS_CAT(a, b, c) DAS_CAT_(a, b, c)
int DAS_CAT(dispatch_accumulate_, DAS_C_LO, DAS_C_HI)(const McAccParams& p, int B, bool v4, int f, cudaStream_t st) {
    return Range<DAS_C_LO>::acc(p, B, v4, f, st);
}
int DAS_CAT(dispatch_finalize_, DAS_C_LO, DAS_C_HI)(const McFinParams& p, int B, bool v4, int f, cudaStream_t st) {
    return Range<DAS_C_LO>::fin(p, B, v4, f, st);
}

}  // namespace das
