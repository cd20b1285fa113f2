// K1 (streaming softmax/argmax -> accumulate), K2 (finalize + image pooling) and the fused K1+K2 kernel.
//
// Data layout in HBM (all per batch of B images, HW = H*W pixels):
//   logits   f32 [B,C,HW]      one buffer per Monte-Carlo pass, read exactly once (no L1 allocation)
//   sum_p    f32 [B,C,HW]      running sum_t softmax(x_t)            (same layout as the logits)
//   sum_ent  f32 [B,HW]        running sum_t entropy(softmax(x_t))
//   votes    u8  [B,T_cap,HW]  argmax of every pass (the reference's outputs[B,T,H,W], as bytes)
// A thread owns VEC consecutive pixels (VEC = 4 / 2: one 128 / 64-bit load per class plane, needs
// HW % VEC == 0; VEC = 1: 513x513-style planes whose class planes are mutually misaligned)
// and keeps the whole class vector of those pixels in registers, so each logit is touched once.
#pragma once

#include "das_common.cuh"

namespace das {

struct McAccParams {
    const float* logits[DAS_MAX_PASS_GROUP];
    float* sum_p;
    float* sum_ent;
    uint8_t* votes;
    long long HW;
    int C, T_cap, n_passes, pass_begin;
};

struct McFinParams {
    const float* sum_p;
    const float* sum_ent;
    const uint8_t* votes;
    const float* labels;
    float* vote_entropy;
    float* pred_entropy;
    float* bald;
    float* confidence;
    float* margin;
    uint8_t* weak_labels;
    float* partials;  // [B, blocks_per_image, DAS_N_SCORES]
    long long HW;
    int C, T_cap, T, blocks_per_image;
};

// fused "last group": accumulate + finalize without writing the state back
struct McScoreParams {
    McAccParams acc;
    McFinParams fin;
};

// ---- VEC-wide float / byte vectors ----------------------------------------------------------
template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using F = float4;
    using U = uint32_t;
};
template <>
struct VecT<2> {
    using F = float2;
    using U = uint16_t;
};
template <>
struct VecT<1> {
    using F = float;
    using U = uint8_t;
};

template <int VEC>
__device__ __forceinline__ void unpack(const typename VecT<VEC>::F& v, float* out);
template <>
__device__ __forceinline__ void unpack<4>(const float4& v, float* out) {
    out[0] = v.x, out[1] = v.y, out[2] = v.z, out[3] = v.w;
}
template <>
__device__ __forceinline__ void unpack<2>(const float2& v, float* out) {
    out[0] = v.x, out[1] = v.y;
}
template <>
__device__ __forceinline__ void unpack<1>(const float& v, float* out) {
    out[0] = v;
}
template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::F pack(const float* in);
template <>
__device__ __forceinline__ float4 pack<4>(const float* in) {
    return make_float4(in[0], in[1], in[2], in[3]);
}
template <>
__device__ __forceinline__ float2 pack<2>(const float* in) {
    return make_float2(in[0], in[1]);
}
template <>
__device__ __forceinline__ float pack<1>(const float* in) {
    return in[0];
}

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::F ldg_stream_v(const float* p) {
    return ldg_stream(reinterpret_cast<const typename VecT<VEC>::F*>(p));
}

// VEC bytes (one per pixel) <-> memory
template <int VEC>
__device__ __forceinline__ uint32_t load_bytes(const uint8_t* p) {
    return (uint32_t)*reinterpret_cast<const typename VecT<VEC>::U*>(p);
}
template <int VEC>
__device__ __forceinline__ void store_bytes(uint8_t* p, uint32_t w) {
    *reinterpret_cast<typename VecT<VEC>::U*>(p) = (typename VecT<VEC>::U)w;
}

constexpr int kAccThreads = 128;

// Running accumulators of a thread: C x VEC float32.  Two homes:
//   registers      (SMEM = false)  - no extra instructions, but C*VEC live registers next to the C*VEC
//                                    registers that hold the in-flight logits (VEC=4, C=19: 224 regs, 8 warps/SM)
//   shared memory  (SMEM = true)   - one conflict-free LDS/STS per class and pass ([c][thread] vectors); the
//                                    register file only holds in-flight logits, more warps fit per SM and more
//                                    bytes are in flight (the kernel is latency bound, not issue bound;
//                                    profiles/r1_k1_notes.md)
template <int C, int VEC, bool SMEM>
struct Acc {
    float r[SMEM ? 1 : C][VEC];
    typename VecT<VEC>::F* s;  // SMEM: &smem[tid], stride kAccThreads between classes
    __device__ __forceinline__ void get(int c, float* out) const {
        if (SMEM)
            unpack<VEC>(s[c * kAccThreads], out);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) out[j] = r[SMEM ? 0 : c][j];
        }
    }
    __device__ __forceinline__ void set(int c, const float* in) {
        if (SMEM)
            s[c * kAccThreads] = pack<VEC>(in);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) r[SMEM ? 0 : c][j] = in[j];
        }
    }
};

constexpr size_t acc_smem_bytes(int C, int VEC) { return (size_t)C * kAccThreads * VEC * sizeof(float); }

// Maximum of N floats as a 3-ary tree (depth 3 for N <= 27 instead of an N-long chain; max is exact, so the shape does
// not change the result).  One template level per tree level: every loop has a constant trip count.  (The same tree
// written as a loop over levels - `for (n = C; n > 1; n = (n + 2) / 3)` - is not unrolled by nvcc 12.9 for C >= 22: the
// array then lives in LOCAL memory, 88 - 128 bytes of stack traffic per pass in every kernel.)
template <int N>
struct MaxTree {
    static __device__ __forceinline__ float run(const float (&t)[N]) {
        constexpr int M = (N + 2) / 3;
        float u[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            float v = t[3 * i];
            if (3 * i + 1 < N) v = fmaxf(v, t[3 * i + 1]);
            if (3 * i + 2 < N) v = fmaxf(v, t[3 * i + 2]);
            u[i] = v;
        }
        return MaxTree<M>::run(u);
    }
};
template <>
struct MaxTree<1> {
    static __device__ __forceinline__ float run(const float (&t)[1]) { return t[0]; }
};

// One Monte-Carlo pass for the VEC pixels of a thread: loads the C logits of each pixel once (streaming, evict
// first), returns the votes packed one byte per pixel and updates the running accumulators.
//   vote   v   = first argmax_c x_c                                     (mc_dropout.py:40)
//   softmax p_c = 2^((x_c - m)*log2e) / s, s = sum_c 2^(...)             (nn.Softmax2d, ceal.py:111)
//   entropy of the pass: -sum p_c log2 p_c = log2 s - (sum_c e_c y_c)/s, y_c = (x_c - m) log2e <= 0
//     (log-sum-exp form of ceal.py:118: one log2 per pixel instead of one per logit; it differs from
//      the reference's "+1e-12" form by < 2e-12 per class and has no cancellation, both terms >= 0)
template <int C, int VEC, bool PROBS, bool VOTES, bool SMEM>
__device__ __forceinline__ uint32_t mc_pass_math(float (&x)[C][VEC], Acc<C, VEC, SMEM>& acc, float* ent);

template <int C, int VEC, bool PROBS, bool VOTES, bool SMEM>
__device__ __forceinline__ uint32_t mc_pass(const float* __restrict__ xp, uint32_t plane_bytes,
                                            Acc<C, VEC, SMEM>& acc, float* ent) {
    float x[C][VEC];
    // plane c of this image lives plane_bytes*c further on
    const char* xb = reinterpret_cast<const char*>(xp);
#pragma unroll
    for (int c = 0; c < C; ++c)
        unpack<VEC>(ldg_stream_v<VEC>(reinterpret_cast<const float*>(xb + (size_t)((uint32_t)c * plane_bytes))), x[c]);
    return mc_pass_math<C, VEC, PROBS, VOTES, SMEM>(x, acc, ent);
}

// the arithmetic of one pass on the C x VEC logits already in registers (shared by the LDG and the TMA kernels)
template <int C, int VEC, bool PROBS, bool VOTES, bool SMEM>
__device__ __forceinline__ uint32_t mc_pass_math(float (&x)[C][VEC], Acc<C, VEC, SMEM>& acc, float* ent) {
    // Latency matters as much as instruction count here (3-4 warps per scheduler): the maximum is a 3-ary tree
    // (depth 3 instead of a 18-long chain; max is exact, so the result is unchanged), the vote flags are collected
    // by two half chains, and the softmax denominator / entropy sums run as 4 / 2 interleaved partial sums.
    uint32_t vote_word = 0;
    float inv[VEC];
    float m[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) t[c] = x[c][j];
        m[j] = MaxTree<C>::run(t);
    }
    // d_c = x_c + (0 - m): exactly +0 for the maxima, negative otherwise (IEEE subtraction of distinct floats is never
    // 0), so the sign bits ARE the "not a maximum" flags: one funnel shift per class collects them (instead of a
    // compare + select), the first maximum is the highest clear bit.  The softmax reuses d: y = d * log2(e).
    // `0 - m` (not `-m`): for m = +-0 it is +0, so a logit of -0 next to a maximum of +0 gives (-0) + (+0) = +0 - torch's
    // argmax compares -0 == +0 and takes the first of them (mc_dropout.py:40); `-m` would give (-0) + (-0) = -0.
    if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int h = 0; h < VEC / 2; ++h) {
            const int j0 = 2 * h, j1 = 2 * h + 1;
            const f32x2 nm = {0.f - m[j0], 0.f - m[j1]};
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const f32x2 d = add2(f32x2{x[c][j0], x[c][j1]}, nm);
                x[c][j0] = d.x, x[c][j1] = d.y;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float nm = 0.f - m[j];
#pragma unroll
            for (int c = 0; c < C; ++c) x[c][j] = x[c][j] + nm;
        }
    }
    if (VOTES) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            uint32_t lo = 0, hi = 0;  // two half chains; bit (n - 1 - i) of a chain = flag of its i-th class
            constexpr int HALF = (C + 1) / 2;
#pragma unroll
            for (int c = 0; c < HALF; ++c) lo = __funnelshift_l(__float_as_uint(x[c][j]), lo, 1);
#pragma unroll
            for (int c = HALF; c < C; ++c) hi = __funnelshift_l(__float_as_uint(x[c][j]), hi, 1);
            const uint32_t notmax = (lo << (C - HALF)) | hi;                   // bit (C - 1 - c) = flag of class c
            const uint32_t ismax = ~notmax & (0xffffffffu >> (32 - C));
            const int v = __clz(ismax) - (32 - C);                            // first maximum = highest set bit
            vote_word |= (uint32_t)v << (8 * j);
        }
    }
    if (PROBS) {
        if constexpr (VEC % 2 == 0) {
            // pixel pairs through the packed fp32 pipe (FFMA2 / FADD2): same per-lane IEEE arithmetic as below
#pragma unroll
            for (int h = 0; h < VEC / 2; ++h) {
                const int j0 = 2 * h, j1 = 2 * h + 1;
                const f32x2 L2 = {kLog2e, kLog2e};
                f32x2 sp[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}}, ap[2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const f32x2 y = mul2(f32x2{x[c][j0], x[c][j1]}, L2);
                    const f32x2 e = {ex2_approx(y.x), ex2_approx(y.y)};
                    sp[c & 3] = add2(sp[c & 3], e);
                    ap[c & 1] = fma2(e, y, ap[c & 1]);
                    x[c][j0] = e.x;
                    x[c][j1] = e.y;
                }
                const f32x2 s = add2(add2(sp[0], sp[1]), add2(sp[2], sp[3])), a = add2(ap[0], ap[1]);
                inv[j0] = rcp_approx(s.x);  // s in [1, C]: 1 ulp, no denormal slow path
                inv[j1] = rcp_approx(s.y);
                ent[j0] += lg2_approx(s.x) - a.x * inv[j0];
                ent[j1] += lg2_approx(s.y) - a.y * inv[j1];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float av[VEC];
                acc.get(c, av);
#pragma unroll
                for (int h = 0; h < VEC / 2; ++h) {
                    const f32x2 r = fma2(f32x2{x[c][2 * h], x[c][2 * h + 1]}, f32x2{inv[2 * h], inv[2 * h + 1]},
                                         f32x2{av[2 * h], av[2 * h + 1]});
                    av[2 * h] = r.x;
                    av[2 * h + 1] = r.y;
                }
                acc.set(c, av);
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float sp[4] = {0.f, 0.f, 0.f, 0.f}, ap[2] = {0.f, 0.f};  // same association as the packed path
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float y = x[c][j] * kLog2e;
                    const float e = ex2_approx(y);
                    sp[c & 3] += e;
                    ap[c & 1] = fmaf(e, y, ap[c & 1]);
                    x[c][j] = e;
                }
                const float s = (sp[0] + sp[1]) + (sp[2] + sp[3]), a = ap[0] + ap[1];
                inv[j] = rcp_approx(s);
                ent[j] += lg2_approx(s) - a * inv[j];
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float av[VEC];
                acc.get(c, av);
#pragma unroll
                for (int j = 0; j < VEC; ++j) av[j] = fmaf(x[c][j], inv[j], av[j]);
                acc.set(c, av);
            }
        }
    }
    return vote_word;
}

// initialise the accumulators: zeros for the first group, otherwise the state written by earlier groups
template <int C, int VEC, bool SMEM>
__device__ __forceinline__ void mc_acc_init(Acc<C, VEC, SMEM>& acc, float* ent, bool first, const float* sum_p,
                                            const float* sum_ent, size_t img_off, size_t map_off, long long HW) {
    using F = typename VecT<VEC>::F;
    float z[VEC];
    if (first) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) z[j] = 0.f, ent[j] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.set(c, z);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            unpack<VEC>(*reinterpret_cast<const F*>(sum_p + img_off + (size_t)c * HW), z);
            acc.set(c, z);
        }
        unpack<VEC>(*reinterpret_cast<const F*>(sum_ent + map_off), ent);
    }
}

// ---------------------------------------------------------------------------------------------
// K1: consume n_passes passes of logits, add to the running state in HBM (streaming form).
// ---------------------------------------------------------------------------------------------
template <int C, int VEC, bool PROBS, bool VOTES, bool SMEM, int MINB>
__global__ void __launch_bounds__(kAccThreads, MINB) mc_accumulate_kernel(const McAccParams p) {
    using F = typename VecT<VEC>::F;
    extern __shared__ float4 acc_smem[];
    const long long pix = ((long long)blockIdx.x * kAccThreads + threadIdx.x) * VEC;
    if (pix >= p.HW) return;
    const int b = blockIdx.y;
    const size_t img_off = (size_t)b * C * p.HW + pix;
    const size_t map_off = (size_t)b * p.HW + pix;

    const uint32_t plane_bytes = (uint32_t)(p.HW * sizeof(float));  // C * plane_bytes < 4 GB is validated on the host
    Acc<C, VEC, SMEM> acc;
    acc.s = reinterpret_cast<F*>(acc_smem) + threadIdx.x;
    float ent[VEC];
    if (PROBS) mc_acc_init<C, VEC, SMEM>(acc, ent, p.pass_begin == 0, p.sum_p, p.sum_ent, img_off, map_off, p.HW);

    for (int g = 0; g < p.n_passes; ++g) {
        const uint32_t vote_word = mc_pass<C, VEC, PROBS, VOTES, SMEM>(p.logits[g] + img_off, plane_bytes, acc, ent);
        if (VOTES) store_bytes<VEC>(p.votes + ((size_t)b * p.T_cap + (p.pass_begin + g)) * p.HW + pix, vote_word);
    }

    if (PROBS) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float a[VEC];
            acc.get(c, a);
            *reinterpret_cast<F*>(p.sum_p + img_off + (size_t)c * p.HW) = pack<VEC>(a);
        }
        *reinterpret_cast<F*>(p.sum_ent + map_off) = pack<VEC>(ent);
    }
}

// ---- pieces shared by K2 and the fused kernel ---------------------------------------------------

// per-thread byte histogram of votes in shared memory: hist8[(class * NT + tid) * VEC + pixel]
// (a thread's counters of one class form one VEC-byte word; the bank is a function of tid only -> no conflicts)
template <int C, int VEC, int NT>
__device__ __forceinline__ void hist_setup(uint8_t* hist8, float* lut, int T, int tid) {
    const float Tf = (float)T;
    for (int n = tid; n <= T; n += NT) {
        const float pr = (float)n / Tf;
        lut[n] = pr * log2f(pr + kEps);  // p * log2(p + 1e-12), p = n / T in float32 (mc_dropout.py:47-48)
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_bytes<VEC>(hist8 + (size_t)(c * NT + tid) * VEC, 0u);
    __syncthreads();
}
template <int VEC, int NT>
__device__ __forceinline__ void hist_add(uint8_t* hist8, uint32_t vote_word, int tid) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) hist8[(((vote_word >> (8 * j)) & 0xff) * NT + tid) * VEC + j] += 1;
}
// VE = sum over classes ascending of -(p log2(p + 1e-12)) (mc_dropout.py:46-48)
template <int C, int VEC, int NT>
__device__ __forceinline__ void hist_vote_entropy(const uint8_t* hist8, const float* lut, int tid, float* ve) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) ve[j] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint32_t w = load_bytes<VEC>(hist8 + (size_t)(c * NT + tid) * VEC);
#pragma unroll
        for (int j = 0; j < VEC; ++j) ve[j] = ve[j] - lut[(w >> (8 * j)) & 0xff];
    }
}

template <int C, int VEC>
__device__ __forceinline__ void load_valid(const float* labels, size_t map_off, bool* valid) {
    if (labels != nullptr) {
        float lab[VEC];
        unpack<VEC>(*reinterpret_cast<const typename VecT<VEC>::F*>(labels + map_off), lab);
#pragma unroll
        for (int j = 0; j < VEC; ++j) valid[j] = !((lab[j] < 0.f) || (lab[j] >= (float)C));  // mc_dropout.py:45
    } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) valid[j] = true;
    }
}

// predictive entropy / confidence / margin of p_bar = sum_p / T with the reference's formulas
// (ceal.py:36,84-90,116-118); BALD = pred_entropy - sum_ent / T; masking as mc_dropout.py:49, ceal.py:39,91
template <int C, int VEC, typename GetAcc>
__device__ __forceinline__ void probs_scores(GetAcc get_acc, const float* ent, float Tf, const bool* valid,
                                             float (*sc)[VEC]) {
    // mean over passes as a multiplication by fl(1/T): exact for T = 1 (the reference-pinned CEAL case) and for
    // power-of-two T, otherwise within one float32 ulp of the division the composed oracle uses
    const float invT = __frcp_rn(Tf);
    float pe[VEC], top1[VEC], top2[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) pe[j] = 0.f, top1[j] = -1.f, top2[j] = -1.f;
    // log2 through MUFU.LG2 (relative error 2^-22 for arguments below 0.5) for every class, then the one term whose
    // argument can lie in (0.5, 1] - the largest probability, where MUFU.LG2 only bounds the ABSOLUTE error and the
    // term itself is tiny - is replaced by its log2f value: 1 accurate logarithm per pixel instead of C
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float a[VEC];
        get_acc(c, a);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const float pb = a[j] * invT;
            pe[j] = pe[j] - pb * lg2_approx(pb + kEps);
            top2[j] = fmaxf(top2[j], fminf(top1[j], pb));
            top1[j] = fmaxf(top1[j], pb);
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        pe[j] = (pe[j] + top1[j] * lg2_approx(top1[j] + kEps)) - top1[j] * log2f(top1[j] + kEps);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const float ee = ent[j] * invT;
        sc[DAS_SCORE_PRED_ENTROPY][j] = valid[j] ? pe[j] : 0.f;
        sc[DAS_SCORE_EXPECTED_ENTROPY][j] = valid[j] ? ee : 0.f;
        sc[DAS_SCORE_BALD][j] = valid[j] ? pe[j] - ee : 0.f;
        sc[DAS_SCORE_CONFIDENCE][j] = valid[j] ? top1[j] : 1.f;
        sc[DAS_SCORE_MARGIN][j] = valid[j] ? top1[j] - top2[j] : 1.f;
    }
}

template <int VEC>
__device__ __forceinline__ void store_maps(const McFinParams& f, size_t map_off, float (*sc)[VEC], bool probs,
                                           bool votes) {
    using F = typename VecT<VEC>::F;
    if (probs) {
        if (f.pred_entropy) *reinterpret_cast<F*>(f.pred_entropy + map_off) = pack<VEC>(sc[DAS_SCORE_PRED_ENTROPY]);
        if (f.bald) *reinterpret_cast<F*>(f.bald + map_off) = pack<VEC>(sc[DAS_SCORE_BALD]);
        if (f.confidence) *reinterpret_cast<F*>(f.confidence + map_off) = pack<VEC>(sc[DAS_SCORE_CONFIDENCE]);
        if (f.margin) *reinterpret_cast<F*>(f.margin + map_off) = pack<VEC>(sc[DAS_SCORE_MARGIN]);
    }
    if (votes && f.vote_entropy)
        *reinterpret_cast<F*>(f.vote_entropy + map_off) = pack<VEC>(sc[DAS_SCORE_VOTE_ENTROPY]);
}

// vote of pass 0, 255 where invalid (ceal.py:157-163)
template <int VEC>
__device__ __forceinline__ void store_weak_labels(uint8_t* weak_labels, size_t map_off, uint32_t first_vote,
                                                  const bool* valid) {
    uint32_t wl = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) wl |= (valid[j] ? ((first_vote >> (8 * j)) & 0xffu) : 255u) << (8 * j);
    store_bytes<VEC>(weak_labels + map_off, wl);
}

// image pooling: thread -> warp shuffle -> shared memory -> one partial row per block (fixed order)
struct SyncBlock {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// barrier among the first N threads of the block only (named barrier 1): the TMA kernel's consumer warps
template <int N>
struct SyncNamed {
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }
};

template <int VEC, int NT, typename Sync = SyncBlock>
__device__ __forceinline__ void block_partials(float (*sc)[VEC], float (*red)[NT / 32], float* partials_row, int tid,
                                               Sync sync = Sync()) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < DAS_N_SCORES; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) s += sc[k][j];
        s = warp_sum(s);
        if (lane == 0) red[k][wid] = s;
    }
    sync();
    if (tid < DAS_N_SCORES) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[tid][w];
        partials_row[tid] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// K2: state -> per-pixel maps + per-block partial sums of every score.
// ---------------------------------------------------------------------------------------------
template <int C, int VEC, bool PROBS, bool VOTES>
// 3 blocks per SM (<= 85 registers): measured 133 us against 154 us at 2 blocks (100 registers) and 153 us at 4 (64, spills)
// for 8 images of 512 x 1024, C = 19, T = 20
__global__ void __launch_bounds__(kFinalizeThreads, 3) mc_finalize_kernel(const McFinParams p) {
    using F = typename VecT<VEC>::F;
    constexpr int NT = kFinalizeThreads;
    __shared__ uint32_t hist32[VOTES ? C * NT * VEC / 4 : 1];
    __shared__ float lut[VOTES ? 256 : 1];
    __shared__ float red[DAS_N_SCORES][NT / 32];
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist32);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const long long pix = ((long long)blockIdx.x * NT + tid) * VEC;
    const bool active = pix < p.HW;
    if (VOTES) hist_setup<C, VEC, NT>(hist8, lut, p.T, tid);

    float sc[DAS_N_SCORES][VEC];
#pragma unroll
    for (int k = 0; k < DAS_N_SCORES; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) sc[k][j] = 0.f;

    if (active) {
        const size_t map_off = (size_t)b * p.HW + pix;
        bool valid[VEC];
        load_valid<C, VEC>(p.labels, map_off, valid);
        if (PROBS) {
            const float* sp = p.sum_p + (size_t)b * C * p.HW + pix;
            float se[VEC];
            unpack<VEC>(*reinterpret_cast<const F*>(p.sum_ent + map_off), se);
            probs_scores<C, VEC>([&](int c, float* a) { unpack<VEC>(*reinterpret_cast<const F*>(sp + (size_t)c * p.HW), a); },
                                 se, (float)p.T, valid, sc);
        }
        if (VOTES) {
            const uint8_t* vp = p.votes + (size_t)b * p.T_cap * p.HW + pix;
            uint32_t first_vote = 0;
            constexpr int U = 4;
            for (int t0 = 0; t0 < p.T; t0 += U) {
                uint32_t w[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t0 + u < p.T) w[u] = load_bytes<VEC>(vp + (size_t)(t0 + u) * p.HW);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t0 + u < p.T) hist_add<VEC, NT>(hist8, w[u], tid);
                if (t0 == 0) first_vote = w[0];
            }
            float ve[VEC];
            hist_vote_entropy<C, VEC, NT>(hist8, lut, tid, ve);
#pragma unroll
            for (int j = 0; j < VEC; ++j) sc[DAS_SCORE_VOTE_ENTROPY][j] = valid[j] ? ve[j] : 0.f;
            if (p.weak_labels) store_weak_labels<VEC>(p.weak_labels, map_off, first_vote, valid);
        }
        store_maps<VEC>(p, map_off, sc, PROBS, VOTES);
    }
    block_partials<VEC, NT>(sc, red, p.partials + ((size_t)b * p.blocks_per_image + blockIdx.x) * DAS_N_SCORES, tid);
}

// ---------------------------------------------------------------------------------------------
// K1+K2 fused, for the LAST pass group of a batch: the group's logits are consumed exactly as in K1, but
// the accumulators never go back to HBM - the finalize arithmetic of K2 runs on them directly and only
// maps / block partials are written.  With pass_begin == 0 (all T passes in one group, T <= 32) no state is
// read or written at all: HBM traffic == the logits, once.  Votes go straight into the per-thread
// shared-memory histogram (votes of earlier groups are re-read from the state).
// ---------------------------------------------------------------------------------------------
template <int C, int VEC, bool PROBS, bool VOTES, bool SMEM, int MINB>
__global__ void __launch_bounds__(kAccThreads, MINB) mc_score_kernel(const McScoreParams q) {
    using F = typename VecT<VEC>::F;
    constexpr int NT = kAccThreads;
    const McAccParams& p = q.acc;
    const McFinParams& f = q.fin;
    extern __shared__ float4 acc_smem[];
    __shared__ uint32_t hist32[VOTES ? C * NT * VEC / 4 + 1 : 1];
    __shared__ float lut[VOTES ? 256 : 1];
    __shared__ float red[DAS_N_SCORES][NT / 32];
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist32);

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const long long pix = ((long long)blockIdx.x * NT + tid) * VEC;
    const bool active = pix < p.HW;
    if (VOTES) hist_setup<C, VEC, NT>(hist8, lut, f.T, tid);

    float sc[DAS_N_SCORES][VEC];
#pragma unroll
    for (int k = 0; k < DAS_N_SCORES; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) sc[k][j] = 0.f;

    if (active) {
        const size_t img_off = (size_t)b * C * p.HW + pix;
        const size_t map_off = (size_t)b * p.HW + pix;
    
        const uint32_t plane_bytes = (uint32_t)(p.HW * sizeof(float));
        Acc<C, VEC, SMEM> acc;
        acc.s = reinterpret_cast<F*>(acc_smem) + tid;
        float ent[VEC];
        uint32_t first_vote = 0;
        if (PROBS) mc_acc_init<C, VEC, SMEM>(acc, ent, p.pass_begin == 0, p.sum_p, p.sum_ent, img_off, map_off, p.HW);
        if (VOTES) {  // votes recorded by earlier groups of this batch
            const uint8_t* vp = p.votes + (size_t)b * p.T_cap * p.HW + pix;
            for (int t = 0; t < p.pass_begin; ++t) {
                const uint32_t w = load_bytes<VEC>(vp + (size_t)t * p.HW);
                if (t == 0) first_vote = w;
                hist_add<VEC, NT>(hist8, w, tid);
            }
        }

        for (int g = 0; g < p.n_passes; ++g) {
            const uint32_t vote_word = mc_pass<C, VEC, PROBS, VOTES, SMEM>(p.logits[g] + img_off, plane_bytes, acc, ent);
            if (VOTES) {
                hist_add<VEC, NT>(hist8, vote_word, tid);
                if (g == 0 && p.pass_begin == 0) first_vote = vote_word;
                if (p.votes != nullptr)  // keep the recorded votes complete (absent in single-shot states)
                    store_bytes<VEC>(p.votes + ((size_t)b * p.T_cap + (p.pass_begin + g)) * p.HW + pix, vote_word);
            }
        }

        // ---- finalize straight from the accumulators (same arithmetic as mc_finalize_kernel) ----
        bool valid[VEC];
        load_valid<C, VEC>(f.labels, map_off, valid);
        if (PROBS) probs_scores<C, VEC>([&](int c, float* a) { acc.get(c, a); }, ent, (float)f.T, valid, sc);
        if (VOTES) {
            float ve[VEC];
            hist_vote_entropy<C, VEC, NT>(hist8, lut, tid, ve);
#pragma unroll
            for (int j = 0; j < VEC; ++j) sc[DAS_SCORE_VOTE_ENTROPY][j] = valid[j] ? ve[j] : 0.f;
            if (f.weak_labels) store_weak_labels<VEC>(f.weak_labels, map_off, first_vote, valid);
        }
        store_maps<VEC>(f, map_off, sc, PROBS, VOTES);
    }
    block_partials<VEC, NT>(sc, red, f.partials + ((size_t)b * f.blocks_per_image + blockIdx.x) * DAS_N_SCORES, tid);
}

}  // namespace das
#include "mc_tma.cuh"
#include "mc_up.cuh"
#include "mc_up1.cuh"
namespace das {

// per-class-count launchers (instantiated in mc_inst.cu for a range of C)
template <int C>
int launch_accumulate(const McAccParams& p, int B, int vec, int flags, cudaStream_t st);
template <int C>
int launch_finalize(const McFinParams& p, int B, int vec, int flags, cudaStream_t st);
template <int C>
int launch_score(const McScoreParams& p, int B, int vec, int flags, cudaStream_t st);
template <int C>
int launch_score_tma(const McTmaParams& p, int flags, int ctas_per_sm, cudaStream_t st);

int dispatch_accumulate(const McAccParams& p, int B, int vec, int flags, cudaStream_t st);
int dispatch_finalize(const McFinParams& p, int B, int vec, int flags, cudaStream_t st);
int dispatch_score(const McScoreParams& p, int B, int vec, int flags, cudaStream_t st);
int dispatch_score_tma(const McTmaParams& p, int flags, int ctas_per_sm, cudaStream_t st);

}  // namespace das
