// K1 (streaming softmax/argmax -> accumulate) and K2 (finalize + image pooling) kernel templates.
//
// Data layout in HBM (all per batch of B images, HW = H*W pixels):
//   logits   f32 [B,C,HW]      one buffer per Monte-Carlo pass, read exactly once (evict-first)
//   sum_p    f32 [B,C,HW]      running sum_t softmax(x_t)            (same layout as the logits)
//   sum_ent  f32 [B,HW]        running sum_t entropy(softmax(x_t))
//   votes    u8  [B,T_cap,HW]  argmax of every pass (the reference's outputs[B,T,H,W], as bytes)
// A thread owns VEC consecutive pixels (VEC = 4: one 128-bit load per class plane, needs HW % 4 == 0;
// VEC = 1: 513x513-style planes whose class planes are mutually misaligned) and keeps
// the whole class vector of those pixels in registers, so each logit is touched once.
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

template <int VEC>
struct VecT;
template <>
struct VecT<4> {
    using F = float4;
};
template <>
struct VecT<1> {
    using F = float;
};

template <int VEC>
__device__ __forceinline__ void unpack(const typename VecT<VEC>::F& v, float* out);
template <>
__device__ __forceinline__ void unpack<4>(const float4& v, float* out) {
    out[0] = v.x, out[1] = v.y, out[2] = v.z, out[3] = v.w;
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
__device__ __forceinline__ float pack<1>(const float* in) {
    return in[0];
}

constexpr int kAccThreads = 128;

// ---------------------------------------------------------------------------------------------
// K1: consume n_passes passes of logits for VEC pixels per thread.
//   vote   v   = first argmax_c x_c                                     (mc_dropout.py:40)
//   softmax p_c = 2^(x_c*log2e - m*log2e) / s, s = sum_c 2^(...)        (nn.Softmax2d, ceal.py:111)
//   entropy of the pass: -sum p_c log2 p_c = log2 s - (sum_c e_c y_c)/s, y_c = (x_c - m) log2e <= 0
//     (log-sum-exp form of ceal.py:118: one log2 per pixel instead of one per logit; it differs from
//      the reference's "+1e-12" form by < 2e-12 per class and has no cancellation, both terms >= 0)
// ---------------------------------------------------------------------------------------------
template <int C, int VEC, bool PROBS, bool VOTES>
__global__ void __launch_bounds__(kAccThreads) mc_accumulate_kernel(const McAccParams p) {
    using F = typename VecT<VEC>::F;
    const long long pix = ((long long)blockIdx.x * kAccThreads + threadIdx.x) * VEC;
    if (pix >= p.HW) return;
    const int b = blockIdx.y;
    const size_t img_off = (size_t)b * C * p.HW + pix;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();

    float acc[PROBS ? C : 1][VEC];
    float ent[VEC];
    if (PROBS) {
        if (p.pass_begin == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[c][j] = 0.f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) ent[j] = 0.f;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                F v = ldg_hint(reinterpret_cast<const F*>(p.sum_p + img_off + (size_t)c * p.HW), pol_keep);
                unpack<VEC>(v, acc[c]);
            }
            F e = ldg_hint(reinterpret_cast<const F*>(p.sum_ent + (size_t)b * p.HW + pix), pol_keep);
            unpack<VEC>(e, ent);
        }
    }

    for (int g = 0; g < p.n_passes; ++g) {
        const float* xp = p.logits[g] + img_off;
        float x[C][VEC];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            F v = ldg_stream(reinterpret_cast<const F*>(xp + (size_t)c * p.HW), pol_stream);
            unpack<VEC>(v, x[c]);
        }
        uint32_t vote_word = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float m = x[0][j];
#pragma unroll
            for (int c = 1; c < C; ++c) m = fmaxf(m, x[c][j]);
            if (VOTES) {
                int v = 0;
#pragma unroll
                for (int c = C - 1; c >= 0; --c) v = (x[c][j] == m) ? c : v;  // first max wins
                vote_word |= (uint32_t)v << (8 * j);
            }
            if (PROBS) {
                const float mL = m * kLog2e;
                float s = 0.f, a = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float y = fmaf(x[c][j], kLog2e, -mL);
                    const float e = ex2_approx(y);
                    s += e;
                    a = fmaf(e, y, a);
                    x[c][j] = e;
                }
                const float inv = __frcp_rn(s);
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c][j] = fmaf(x[c][j], inv, acc[c][j]);
                ent[j] += log2f(s) - a * inv;
            }
        }
        if (VOTES) {
            uint8_t* vp = p.votes + ((size_t)b * p.T_cap + (p.pass_begin + g)) * p.HW + pix;
            if (VEC == 4)
                *reinterpret_cast<uint32_t*>(vp) = vote_word;
            else
                *vp = (uint8_t)vote_word;
        }
    }

    if (PROBS) {
#pragma unroll
        for (int c = 0; c < C; ++c)
            stg_hint(reinterpret_cast<F*>(p.sum_p + img_off + (size_t)c * p.HW), pack<VEC>(acc[c]), pol_keep);
        stg_hint(reinterpret_cast<F*>(p.sum_ent + (size_t)b * p.HW + pix), pack<VEC>(ent), pol_keep);
    }
}

// ---------------------------------------------------------------------------------------------
// K2: state -> per-pixel maps + per-block partial sums of every score.
//   vote entropy: per-thread byte histogram of the T votes in shared memory (bank = lane, so no
//   conflicts), then VE = sum_c ascending -(p log2(p + 1e-12)), p = n_c / T, through a T+1 entry table
//   evaluated with the reference's float32 formula (mc_dropout.py:46-48).
//   predictive entropy / confidence / margin on p_bar = sum_p / T with the reference's formulas
//   (ceal.py:36,84-90,116-118); BALD = pred_entropy - sum_ent / T.
// ---------------------------------------------------------------------------------------------
template <int C, int VEC, bool PROBS, bool VOTES>
__global__ void __launch_bounds__(kFinalizeThreads) mc_finalize_kernel(const McFinParams p) {
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
    const float Tf = (float)p.T;

    if (VOTES) {
        for (int n = tid; n <= p.T; n += NT) {
            const float pr = (float)n / Tf;
            lut[n] = pr * log2f(pr + kEps);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (VEC == 4)
                hist32[c * NT + tid] = 0u;
            else
                hist8[c * NT + tid] = 0;
        }
        __syncthreads();
    }

    float sc[DAS_N_SCORES][VEC];
#pragma unroll
    for (int k = 0; k < DAS_N_SCORES; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) sc[k][j] = 0.f;

    if (active) {
        bool valid[VEC];
        if (p.labels != nullptr) {
            float lab[VEC];
            unpack<VEC>(*reinterpret_cast<const F*>(p.labels + (size_t)b * p.HW + pix), lab);
#pragma unroll
            for (int j = 0; j < VEC; ++j) valid[j] = !((lab[j] < 0.f) || (lab[j] >= (float)C));  // mc_dropout.py:45
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) valid[j] = true;
        }

        if (PROBS) {
            float pe[VEC], top1[VEC], top2[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) pe[j] = 0.f, top1[j] = -1.f, top2[j] = -1.f;
            const float* sp = p.sum_p + (size_t)b * C * p.HW + pix;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float a[VEC];
                unpack<VEC>(*reinterpret_cast<const F*>(sp + (size_t)c * p.HW), a);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float pb = __fdiv_rn(a[j], Tf);
                    pe[j] = pe[j] - pb * log2f(pb + kEps);
                    top2[j] = fmaxf(top2[j], fminf(top1[j], pb));
                    top1[j] = fmaxf(top1[j], pb);
                }
            }
            float se[VEC];
            unpack<VEC>(*reinterpret_cast<const F*>(p.sum_ent + (size_t)b * p.HW + pix), se);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float ee = __fdiv_rn(se[j], Tf);
                sc[DAS_SCORE_PRED_ENTROPY][j] = valid[j] ? pe[j] : 0.f;
                sc[DAS_SCORE_EXPECTED_ENTROPY][j] = valid[j] ? ee : 0.f;
                sc[DAS_SCORE_BALD][j] = valid[j] ? pe[j] - ee : 0.f;
                sc[DAS_SCORE_CONFIDENCE][j] = valid[j] ? top1[j] : 1.f;
                sc[DAS_SCORE_MARGIN][j] = valid[j] ? top1[j] - top2[j] : 1.f;
            }
            const size_t o = (size_t)b * p.HW + pix;
            if (p.pred_entropy) *reinterpret_cast<F*>(p.pred_entropy + o) = pack<VEC>(sc[DAS_SCORE_PRED_ENTROPY]);
            if (p.bald) *reinterpret_cast<F*>(p.bald + o) = pack<VEC>(sc[DAS_SCORE_BALD]);
            if (p.confidence) *reinterpret_cast<F*>(p.confidence + o) = pack<VEC>(sc[DAS_SCORE_CONFIDENCE]);
            if (p.margin) *reinterpret_cast<F*>(p.margin + o) = pack<VEC>(sc[DAS_SCORE_MARGIN]);
        }

        if (VOTES) {
            const uint8_t* vp = p.votes + (size_t)b * p.T_cap * p.HW + pix;
            constexpr int U = 4;
            for (int t0 = 0; t0 < p.T; t0 += U) {
                uint32_t w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int t = t0 + u;
                    if (t < p.T) {
                        if (VEC == 4)
                            w[u] = *reinterpret_cast<const uint32_t*>(vp + (size_t)t * p.HW);
                        else
                            w[u] = vp[(size_t)t * p.HW];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (t0 + u < p.T) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const int v = (w[u] >> (8 * j)) & 0xff;
                            hist8[(v * NT + tid) * VEC + j] += 1;
                        }
                    }
                }
                if (t0 == 0 && p.weak_labels) {  // vote of pass 0, 255 where invalid (ceal.py:157-163)
                    uint32_t wl = 0;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) wl |= (valid[j] ? ((w[0] >> (8 * j)) & 0xffu) : 255u) << (8 * j);
                    uint8_t* wp = p.weak_labels + (size_t)b * p.HW + pix;
                    if (VEC == 4)
                        *reinterpret_cast<uint32_t*>(wp) = wl;
                    else
                        *wp = (uint8_t)wl;
                }
            }
            float ve[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) ve[j] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t w;
                if (VEC == 4)
                    w = hist32[c * NT + tid];
                else
                    w = hist8[c * NT + tid];
#pragma unroll
                for (int j = 0; j < VEC; ++j) ve[j] = ve[j] - lut[(w >> (8 * j)) & 0xff];
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) sc[DAS_SCORE_VOTE_ENTROPY][j] = valid[j] ? ve[j] : 0.f;
            if (p.vote_entropy)
                *reinterpret_cast<F*>(p.vote_entropy + (size_t)b * p.HW + pix) = pack<VEC>(sc[DAS_SCORE_VOTE_ENTROPY]);
        }
    }

    // image pooling: thread -> warp shuffle -> shared memory -> one partial row per block
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < DAS_N_SCORES; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) s += sc[k][j];
        s = warp_sum(s);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (tid < DAS_N_SCORES) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[tid][w];
        p.partials[((size_t)b * p.blocks_per_image + blockIdx.x) * DAS_N_SCORES + tid] = s;
    }
}

// per-class-count launchers (instantiated in mc_inst.cu for a range of C)
template <int C>
int launch_accumulate(const McAccParams& p, int B, bool vec4, int flags, cudaStream_t st);
template <int C>
int launch_finalize(const McFinParams& p, int B, bool vec4, int flags, cudaStream_t st);

int dispatch_accumulate(const McAccParams& p, int B, bool vec4, int flags, cudaStream_t st);
int dispatch_finalize(const McFinParams& p, int B, bool vec4, int flags, cudaStream_t st);

}  // namespace das
