// Fused K1+K2 with TMA staging: the single-shot Monte-Carlo reduction (all T passes of a batch in one launch)
// for planes whose size is a multiple of 4 pixels (Cityscapes-shaped pools).
//
// Why: the LDG kernel (mc_score_kernel) is memory-LATENCY bound - a warp issues the 19 loads of a pass, waits,
// computes, issues the next 19 (profiles/r1_k1_notes.md: long_scoreboard 3.3 of 6.1 stalled warps per issue), and
// every load pays a 64-bit address add (54 IADD3 per pass).  Here one elected producer thread streams
// [C x 256 pixel] boxes (one cp.async.bulk.tensor per pass and tile, 19 KB) through a shared-memory ring, so
// bytes in flight no longer depend on occupancy or on the compute phase, and the consumers read their two
// pixels of every class with immediate-offset LDS.64 (no address arithmetic).  CTAs are persistent: the ring
// keeps prefetching the next tile's passes while the consumers finalize the current tile.
//
// Same per-pixel arithmetic (mc_pass_math, probs_scores, the byte histogram) and the same 256-pixel block
// partials as mc_score_kernel<VEC=2>: results are bit-identical between the two kernels.
#pragma once

#include <cuda.h>

namespace das {

constexpr int kTmaPix = 256;      // pixels per tile: 128 consumer threads x 2 pixels (box width of every copy)
constexpr int kTmaFlatPix = 252;  // flat mode: pixels per tile; the 256-float box absorbs a start shift of 0..3
constexpr int kTmaThreads = 160;  // 4 consumer warps + 1 producer warp
constexpr int kTmaMaxStages = 8;

struct McTmaParams {
    // one map per pass buffer.
    // flat == 0: 3-D (HW, C, B), box (256, C, 1): one copy per (tile, pass).  Needs H*W % 4 == 0 (16-byte strides).
    // flat == 1: the buffer as ONE 1-D tensor of B*C*H*W floats, box 256, one copy per class plane.  A TMA box
    //   must START on a 16-byte boundary (tools/probes/tma1d_probe.cu: anything else is an illegal instruction), and
    //   the planes of a 513 x 513 image are mutually misaligned, so every copy starts at the 4-float boundary below
    //   its first pixel and the consumers skip `shift_c` = 0..3 floats of class plane c; tiles are 252 pixels.
    CUtensorMap maps[DAS_MAX_PASS_GROUP];
    McFinParams fin;
    long long HW;
    int B, n_passes, tiles_per_image, stages, flat;
    int num_sms;  // host only: persistent grid = min(tiles, num_sms * CTAs per SM)
};

__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// for warps that have nothing to do until the barrier flips: poll it with a pause in between, so that the polling does not
// take issue slots from the warps that work
__device__ __forceinline__ void tma_mbar_wait_idle(uint32_t bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}

// CTAs per SM: 3 x 160 threads at 128 registers for C <= 20, 2 x 160 at up to 200 registers for more classes.
// (__maxnreg__(136) removes the last 16-byte spill at C = 19 but the per-warp register allocation granularity
// then only fits 2 CTAs per SM: measured 0.85 instead of 0.96 of the HBM peak.)
constexpr int tma_ctas_per_sm(int C) { return C <= 24 ? 3 : 2; }

// FLAT = false: thread t owns pixels (2t, 2t+1) of a 256-pixel tile, 64-bit shared / global accesses.
// FLAT = true : thread t < 126 owns pixels t and t + 126 of a 252-pixel tile: unit-stride (conflict-free, always
//               aligned) 32-bit accesses at a per-class shift.
template <int C, bool PROBS, bool VOTES, bool FLAT>
__global__ void __launch_bounds__(kTmaThreads, tma_ctas_per_sm(C)) mc_score_tma_kernel(const __grid_constant__ McTmaParams q) {
    constexpr int NT = 128, VEC = 2;
    constexpr int TILE = FLAT ? kTmaFlatPix : kTmaPix;
    constexpr int HALF = kTmaFlatPix / 2;  // flat mode: distance between the two pixels of a thread
    constexpr uint32_t kStageBytes = (uint32_t)C * kTmaPix * sizeof(float);
    extern __shared__ __align__(1024) uint8_t ring[];
    __shared__ uint64_t bars[2 * kTmaMaxStages];
    __shared__ uint32_t hist32[VOTES ? C * NT * VEC / 4 + 1 : 1];
    __shared__ float lut[VOTES ? 256 : 1];
    __shared__ float red[DAS_N_SCORES][NT / 32];
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist32);

    const McFinParams& f = q.fin;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = q.stages, T = q.n_passes;
    const uint32_t ring0 = tma_smem_u32(ring), bar0 = tma_smem_u32(bars);
    const int total_tiles = q.B * q.tiles_per_image;
    const uint32_t hw = (uint32_t)q.HW;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(1) : "memory");       // full
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (S + s)), "r"(4) : "memory");  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        // ===== producer: one thread; one bulk tensor copy per (tile, pass) [per class plane in flat mode] =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int b = tile / q.tiles_per_image, px0 = (tile % q.tiles_per_image) * TILE;
                for (int g = 0; g < T; ++g) {
                    tma_mbar_wait(bar0 + 8u * (S + stage), phase ^ 1u);
                    const uint32_t full = bar0 + 8u * stage;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(kStageBytes)
                                 : "memory");
                    if (!FLAT) {
                        asm volatile(
                            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                            "%5}], [%2];" ::"r"(ring0 + stage * kStageBytes),
                            "l"(&q.maps[g]), "r"(full), "r"(px0), "r"(0), "r"(b)
                            : "memory");
                    } else {
                        const uint32_t e0 = (uint32_t)b * C * hw + (uint32_t)px0;  // element index, < 2^31 (host check)
#pragma unroll 1
                        for (int c = 0; c < C; ++c)
                            asm volatile(
                                "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], "
                                "[%2];" ::"r"(ring0 + stage * kStageBytes + c * (kTmaPix * 4)),
                                "l"(&q.maps[g]), "r"(full), "r"((int)((e0 + c * hw) & ~3u))
                                : "memory");
                    }
                    if (++stage == S) stage = 0, phase ^= 1u;
                }
            }
        }
        return;
    }

    // ===== consumers: 128 threads, 2 pixels each =====
    const SyncNamed<NT> sync;
    if (VOTES) {
        const float Tf = (float)f.T;
        for (int n = tid; n <= f.T; n += NT) {
            const float pr = (float)n / Tf;
            lut[n] = pr * log2f(pr + kEps);  // p * log2(p + 1e-12), p = n / T in float32 (mc_dropout.py:47-48)
        }
        sync();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / q.tiles_per_image, blk = tile % q.tiles_per_image;
        // pixel p[j] of this thread inside the plane, act[j]: it exists
        int p[VEC];  // < H*W < 2^30 (mc_validate)
        bool act[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            p[j] = blk * TILE + (FLAT ? tid + j * HALF : tid * VEC + j);
            act[j] = p[j] < (int)hw && (!FLAT || tid < HALF);
        }
        const uint32_t e0 = FLAT ? (uint32_t)b * C * hw + (uint32_t)blk * TILE : 0u;
        if (VOTES) {
#pragma unroll
            for (int c = 0; c < C; ++c) store_bytes<VEC>(hist8 + (size_t)(c * NT + tid) * VEC, 0u);  // thread-private
        }
        Acc<C, VEC, false> acc;
        acc.s = nullptr;
        float ent[VEC], z[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) z[j] = 0.f, ent[j] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.set(c, z);
        uint32_t first_vote = 0;

        for (int g = 0; g < T; ++g) {
            tma_mbar_wait(bar0 + 8u * stage, phase);
            float x[C][VEC];
            if (!FLAT) {
                const float2* sp = reinterpret_cast<const float2*>(ring + stage * kStageBytes) + tid;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float2 v = sp[c * (kTmaPix / 2)];
                    x[c][0] = v.x, x[c][1] = v.y;
                }
            } else {
                const float* sp = reinterpret_cast<const float*>(ring + stage * kStageBytes) + (tid < HALF ? tid : 0);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float* sc_ = sp + c * kTmaPix + ((e0 + c * hw) & 3u);  // skip the alignment shift of plane c
                    x[c][0] = sc_[0], x[c][1] = sc_[HALF];
                }
            }
            __syncwarp();
            if (lane == 0)  // this warp has its copy of the stage in registers: hand the slot back
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (S + stage)) : "memory");
            if (++stage == S) stage = 0, phase ^= 1u;
            const uint32_t vote_word = mc_pass_math<C, VEC, PROBS, VOTES, false>(x, acc, ent);
            if (VOTES) {
                hist_add<VEC, NT>(hist8, vote_word, tid);
                if (g == 0) first_vote = vote_word;
            }
        }

        float sc[DAS_N_SCORES][VEC];
#pragma unroll
        for (int k = 0; k < DAS_N_SCORES; ++k)
#pragma unroll
            for (int j = 0; j < VEC; ++j) sc[k][j] = 0.f;
        if (act[0] || act[1]) {
            const size_t map0 = (size_t)b * q.HW;
            bool valid[VEC];
            if (!FLAT) {
                load_valid<C, VEC>(f.labels, map0 + p[0], valid);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    valid[j] = true;
                    if (f.labels != nullptr && act[j]) {
                        const float lab = f.labels[map0 + p[j]];
                        valid[j] = !((lab < 0.f) || (lab >= (float)C));  // mc_dropout.py:45
                    }
                }
            }
            if (PROBS) probs_scores<C, VEC>([&](int c, float* a) { acc.get(c, a); }, ent, (float)f.T, valid, sc);
            if (VOTES) {
                float ve[VEC];
                hist_vote_entropy<C, VEC, NT>(hist8, lut, tid, ve);
#pragma unroll
                for (int j = 0; j < VEC; ++j) sc[DAS_SCORE_VOTE_ENTROPY][j] = valid[j] ? ve[j] : 0.f;
            }
            if (!FLAT) {
                if (VOTES && f.weak_labels) store_weak_labels<VEC>(f.weak_labels, map0 + p[0], first_vote, valid);
                store_maps<VEC>(f, map0 + p[0], sc, PROBS, VOTES);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (act[j]) {
                        const size_t o = map0 + p[j];
                        if (PROBS) {
                            if (f.pred_entropy) f.pred_entropy[o] = sc[DAS_SCORE_PRED_ENTROPY][j];
                            if (f.bald) f.bald[o] = sc[DAS_SCORE_BALD][j];
                            if (f.confidence) f.confidence[o] = sc[DAS_SCORE_CONFIDENCE][j];
                            if (f.margin) f.margin[o] = sc[DAS_SCORE_MARGIN][j];
                        }
                        if (VOTES && f.vote_entropy) f.vote_entropy[o] = sc[DAS_SCORE_VOTE_ENTROPY][j];
                        if (VOTES && f.weak_labels)
                            f.weak_labels[o] = valid[j] ? (uint8_t)((first_vote >> (8 * j)) & 0xffu) : (uint8_t)255;
                    } else {  // no such pixel: contributes nothing to the image means
#pragma unroll
                        for (int k = 0; k < DAS_N_SCORES; ++k) sc[k][j] = 0.f;
                    }
                }
            }
        }
        block_partials<VEC, NT>(sc, red, f.partials + ((size_t)b * f.blocks_per_image + blk) * DAS_N_SCORES, tid, sync);
        sync();  // `red` is reused by the next tile
    }
}

}  // namespace das
