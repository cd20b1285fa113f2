// Fused upsample + K1 + K2, ONE pixel per lane with the packed fp32 pipe working on CLASS pairs.
//
// mc_up.cuh gives a lane a horizontal pixel pair (FFMA2 / FADD2 / FMUL2 on the two pixels): 38 accumulators + 38 logits
// in flight = 128 registers, i.e. 4 warps per scheduler, and the kernel sits at 0.52 of its issue bound because three
// pipes (issue, FMA, XU) are all within 20 % of each other and four warps cannot keep them overlapped
// (profiles/r2_upsample_notes.md).  Here a lane owns ONE pixel and the packed instructions pair classes (2k, 2k+1) of that
// pixel instead: the same number of issue slots per pixel, half the registers per thread, more warps per scheduler.
// Moving the accumulators to shared memory (round 2, measured, +19 %) bought the same occupancy with 38 more LDS/STS per
// pass in an MIO queue that was already throttling; this form adds none.  Measured (profiles/r2_upsample_notes.md): issue
// slots 66 % busy instead of 56 % for 17 % more instructions - faster for C <= 20 (20 consumer warps, 80 registers) and for
// C >= 22, where the pixel-pair kernel spills (16 consumer warps, 96 registers); mc_api.cu: up_warps() chooses.
//
//   tile      16 output rows x 2 NW columns per CTA; consumer warp -> 16 x 2 strip, lane -> pixel (row lane / 2, column
//             lane % 2)
//   producer  NP warps (4: ONE producer warp starves 20+ consumers - a third of all stall samples sat on the full-barrier
//             wait, 1.65 instead of 1.0 ms), classes dealt round robin; 4-byte cp.async into a ring with an mbarrier per
//             stage as in mc_up.cuh, but the window is stored with the two classes of a pair interleaved:
//             [pair][6 rows][WS columns][2], so that ...
//   phase 1   ... one lane per (class pair, source row) item reads the 3 window columns of both classes with 3 LDS.64,
//             interpolates the strip's 2 pixels HORIZONTALLY in the 3-weight form of mc_up.cuh with the pair in the two
//             halves of FFMA2 / FMUL2, and writes {px0.c0, px0.c1, px1.c0, px1.c1} with one STS.128 into
//             rows[warp][pair][row][px][2] (one 128-byte line per pair)
//   phase 2   a lane reads the pair's two interpolated rows around its pixel (2 LDS.64, one wavefront each) and
//             interpolates VERTICALLY: FMUL2 + FFMA2 per class PAIR
//   pass math mc_pass_math's arithmetic with the pair in the packed halves; the partial sums keep their association
//             (class c adds to chain c % 4 of the denominator and c % 2 of the entropy sum: pair k feeds chains
//             {2k % 4, 2k % 4 + 1} = one f32x2 register, and {0, 1}), so every map value has the bits of the pixel-pair
//             kernels.  An odd class count leaves the second half of the last pair unused (scalar tail).
#pragma once

namespace das {

constexpr int kUp1Strip = 2;  // output columns per consumer warp
__host__ __device__ constexpr int up1_tile_w(int NW) { return kUp1Strip * NW; }
__host__ __device__ constexpr int up1_win_cols(int NW) { return (NW + 1) / 2 + 2; }
__host__ __device__ constexpr int up1_win_stride(int NW) { return up1_win_cols(NW) | 1; }  // odd: conflict-free LDS.64 per half warp
constexpr int up1_pairs(int C) { return (C + 1) / 2; }
constexpr size_t up1_stage_bytes(int C, int NW) { return (size_t)up1_pairs(C) * kUpRows * up1_win_stride(NW) * 2 * sizeof(float); }
constexpr size_t up1_rows_bytes(int C, int NW) { return (size_t)NW * up1_pairs(C) * 32 * sizeof(float); }
constexpr size_t up1_wts_bytes(int NW) { return (size_t)NW * 12 * sizeof(float); }

// One Monte-Carlo pass on the C logits of ONE pixel held as class pairs x[k] = {class 2k, class 2k + 1}.
// Same arithmetic and the same association of every sum as mc_pass_math<C, VEC, ...> (mc_kernels.cuh).
template <int C, bool PROBS, bool VOTES>
__device__ __forceinline__ uint32_t mc_pass_math_pairs(f32x2 (&x)[(C + 1) / 2], f32x2 (&acc)[(C + 1) / 2], float& ent) {
    constexpr int CP = (C + 1) / 2, FULL = C / 2;  // FULL pairs have both halves
    float m;
    {
        float t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) t[c] = (c & 1) ? x[c >> 1].y : x[c >> 1].x;
        m = MaxTree<C>::run(t);
    }
    // d_c = x_c + (0 - m): +0 for the maxima (either zero), negative otherwise - see mc_pass_math
    const float nm1 = 0.f - m;
    const f32x2 nm = {nm1, nm1};
#pragma unroll
    for (int k = 0; k < FULL; ++k) x[k] = add2(x[k], nm);
    if (C & 1) x[CP - 1].x = x[CP - 1].x + nm1;
    uint32_t vote = 0;
    if (VOTES) {
        uint32_t lo = 0, hi = 0;
        constexpr int HALF = (C + 1) / 2;
#pragma unroll
        for (int c = 0; c < HALF; ++c) lo = __funnelshift_l(__float_as_uint((c & 1) ? x[c >> 1].y : x[c >> 1].x), lo, 1);
#pragma unroll
        for (int c = HALF; c < C; ++c) hi = __funnelshift_l(__float_as_uint((c & 1) ? x[c >> 1].y : x[c >> 1].x), hi, 1);
        const uint32_t notmax = (lo << (C - HALF)) | hi;
        const uint32_t ismax = ~notmax & (0xffffffffu >> (32 - C));
        vote = (uint32_t)(__clz(ismax) - (32 - C));
    }
    if (PROBS) {
        const f32x2 L2 = {kLog2e, kLog2e};
        f32x2 s01 = {0.f, 0.f}, s23 = {0.f, 0.f}, ap = {0.f, 0.f};  // chains {0,1}, {2,3} of the denominator; {0,1} of sum e*y
#pragma unroll
        for (int k = 0; k < FULL; ++k) {
            const f32x2 y = mul2(x[k], L2);
            const f32x2 e = {ex2_approx(y.x), ex2_approx(y.y)};
            if ((2 * k) & 2) s23 = add2(s23, e);
            else s01 = add2(s01, e);
            ap = fma2(e, y, ap);
            x[k] = e;
        }
        if (C & 1) {  // class C - 1 (even index): chain (C - 1) % 4 of the denominator, chain 0 of sum e*y
            const float y = x[CP - 1].x * kLog2e;
            const float e = ex2_approx(y);
            if ((C - 1) & 2) s23.x += e;
            else s01.x += e;
            ap.x = fmaf(e, y, ap.x);
            x[CP - 1].x = e;
        }
        const float s = (s01.x + s01.y) + (s23.x + s23.y), a = ap.x + ap.y;
        const float inv = rcp_approx(s);
        ent += lg2_approx(s) - a * inv;
        const f32x2 inv2 = {inv, inv};
#pragma unroll
        for (int k = 0; k < FULL; ++k) acc[k] = fma2(x[k], inv2, acc[k]);
        if (C & 1) acc[CP - 1].x = fmaf(x[CP - 1].x, inv, acc[CP - 1].x);
    }
    return vote;
}

template <int C, bool PROBS, bool VOTES, int NW, int NP, int MINB>
__global__ void __launch_bounds__(32 * (NW + NP), MINB) mc_score_up1_kernel(const __grid_constant__ McUpParams q) {
    constexpr int NT = 32 * NW;
    constexpr int CP = (C + 1) / 2;
    constexpr int ITEMS = CP * kUpRows;                     // (class pair, source row) items per strip
    constexpr int P1_ITERS = (ITEMS + 31) / 32;
    constexpr int WC = up1_win_cols(NW), WS = up1_win_stride(NW), TW = up1_tile_w(NW);
    constexpr int kSlots = kUpRows * WC;                    // floats the producer copies per class
    constexpr int SLOT_ITERS = (kSlots + 31) / 32;
    constexpr int kPairFloats = kUpRows * WS * 2;           // floats of one class pair in a staged window
    constexpr uint32_t kStageBytes = (uint32_t)(CP * kPairFloats * sizeof(float));
    constexpr uint32_t kRowsBytes = (uint32_t)(NW * CP * 32 * sizeof(float));
    constexpr uint32_t kWtsBytes = (uint32_t)(NW * 12 * sizeof(float));
    // [NW][CP][32] interpolated rows | [NW][12] strip weights | ring of [CP][6][WS][2] windows
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2 * kUpMaxStages];
    __shared__ uint32_t hist32[VOTES ? C * NT / 4 + 1 : 1];
    __shared__ float lut[VOTES ? 256 : 1];
    __shared__ float red[DAS_N_SCORES][NT / 32];
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist32);

    const McFinParams& f = q.fin;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = q.stages, T = q.n_passes;
    const uint32_t ring0 = tma_smem_u32(smem) + kRowsBytes + kWtsBytes, bar0 = tma_smem_u32(bars);
    const int tiles_per_image = q.tiles_x * q.tiles_y;
    const int total_tiles = q.B * tiles_per_image;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(32 * NP) : "memory");   // full
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (S + s)), "r"(NW) : "memory");  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NW) {
        // ===== producer warps (NP of them, classes dealt round robin): lane -> slots (row, col) of the 6 x WC window,
        // one 4-byte cp.async per class and slot.  One warp cannot issue the copies of 30 consumer warps fast enough
        // (ncu: a third of all stall samples at the consumers' full-barrier wait) =====
        // Fully unrolled: the element offsets of every (class, slot) copy of this warp are formed once per tile, a
        // copy is then one 64-bit multiply-add for the address and the cp.async itself (with NP even, a warp's classes
        // pw, pw + NP, ... all sit in the same half of their pairs, so their shared-memory offsets are immediates).
        static_assert(NP % 2 == 0, "class pairs: a producer warp serves classes of one parity");
        constexpr int CPW = (C + NP - 1) / NP;  // classes per producer warp (upper bound)
        const int pw = warp - NW;
        const uint32_t plane = (uint32_t)(q.h * q.w);  // C * h * w < 2^31 (host check)
        uint32_t soff[SLOT_ITERS];  // byte position of the lane's slots inside the staged window, first class of this warp
#pragma unroll
        for (int j = 0; j < SLOT_ITERS; ++j) {
            const int s = lane + 32 * j;
            soff[j] = (uint32_t)(((s / WC) * WS + s % WC) * 8 + ((pw >> 1) * kPairFloats + (pw & 1)) * 4);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
            const int ty0 = (t_in / q.tiles_x) * kUpTileH, tx0 = (t_in % q.tiles_x) * TW;
            const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
            const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
            uint32_t eoff[CPW][SLOT_ITERS];  // clamped source coordinates: slots past the plane edge repeat the edge
#pragma unroll
            for (int j = 0; j < SLOT_ITERS; ++j) {
                const int s = lane + 32 * j;
                const uint32_t g0 = (uint32_t)(min(r_base + s / WC, q.h - 1) * q.w + min(c_base + s % WC, q.w - 1));
#pragma unroll
                for (int i = 0; i < CPW; ++i) eoff[i][j] = ((uint32_t)(b * C + pw + NP * i)) * plane + g0;
            }
            for (int g = 0; g < T; ++g) {
                tma_mbar_wait(bar0 + 8u * (S + stage), phase ^ 1u);
                const float* src = q.lowres[g];
                const uint32_t dst0 = ring0 + stage * kStageBytes;
#pragma unroll
                for (int i = 0; i < CPW; ++i) {
                    if (pw + NP * i < C) {
#pragma unroll
                        for (int j = 0; j < SLOT_ITERS; ++j)
                            if (lane + 32 * j < kSlots)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + soff[j] + (uint32_t)(i * (NP / 2) * kPairFloats * 4)),
                                             "l"(src + eoff[i][j])
                                             : "memory");
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8u * stage) : "memory");
                if (++stage == S) stage = 0, phase ^= 1u;
            }
        }
        return;
    }

    // ===== consumers: warp -> columns [2 warp, 2 warp + 2) of the tile; lane -> pixel (row lane / 2, column lane % 2) =====
    const SyncNamed<NT> sync;
    if (VOTES) {
        const float Tf = (float)f.T;
        for (int n = tid; n <= f.T; n += NT) {
            const float pr = (float)n / Tf;
            lut[n] = pr * log2f(pr + kEps);  // p * log2(p + 1e-12), p = n / T in float32 (mc_dropout.py:47-48)
        }
        sync();
    }
    const int xp = lane & 1, yy = lane >> 1;
    float* rows = reinterpret_cast<float*>(smem) + warp * (CP * 32);          // warp-private
    float* wts = reinterpret_cast<float*>(smem + kRowsBytes) + warp * 12;     // warp-private
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
        const int ty0 = (t_in / q.tiles_x) * kUpTileH, tx0 = (t_in % q.tiles_x) * TW;
        const int xs0 = tx0 + kUp1Strip * warp;  // first column of this warp's strip
        const int y = ty0 + yy, x = xs0 + xp;
        const bool act = y < q.H && x < q.W;
        // ---- geometry (ATen align_corners) ----
        const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
        const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
        int y0, ys;
        float ly0, ly1;
        up_source(q.rh, min(y, q.H - 1), q.h, y0, ys, ly0, ly1);
        const int top_idx = (y0 - r_base) * 4 + 2 * xp;  // float index inside the interpolated rows of pair 0
        const int bot_idx = top_idx + ys * 4;
        // the strip's 2 pixels read source columns cb, cb+1, cb+2 of the window: pixel j = W0[j] v0 + W1[j] v1 + W2[j] v2
        // with (W0,W1,W2) = (l0,l1,0) or (0,l0,l1); stored duplicated ({w,w}) so that one LDS.128 yields two packed operands
        int cb;
        {
            int x0f, st;
            float t0, t1;
            up_source(q.rw, min(xs0, q.W - 1), q.w, x0f, st, t0, t1);
            cb = x0f - c_base;
            __syncwarp();
            if (lane < kUp1Strip) {
                int x0;
                up_source(q.rw, min(xs0 + lane, q.W - 1), q.w, x0, st, t0, t1);
                const bool sh = x0 != x0f;  // this pixel starts one source column further right
                const float a0 = sh ? 0.f : t0, a1 = sh ? t0 : t1, a2 = sh ? t1 : 0.f;
                wts[2 * lane] = a0, wts[2 * lane + 1] = a0;
                wts[4 + 2 * lane] = a1, wts[4 + 2 * lane + 1] = a1;
                wts[8 + 2 * lane] = a2, wts[8 + 2 * lane + 1] = a2;
            }
            __syncwarp();
        }
        const int i0 = 2 * cb, i1 = 2 * min(cb + 1, WC - 1), i2 = 2 * min(cb + 2, WC - 1);  // never past the window
        const f32x2 LY0 = {ly0, ly0}, LY1 = {ly1, ly1};

        if (VOTES) {
#pragma unroll
            for (int c = 0; c < C; ++c) hist8[c * NT + tid] = 0;  // thread-private counters
        }
        f32x2 acc[CP];
#pragma unroll
        for (int k = 0; k < CP; ++k) acc[k] = f32x2{0.f, 0.f};
        float ent = 0.f;
        uint32_t first_vote = 0;

        // A warp whose strip lies past the right edge of the plane (the last tile of a row: 1024 = 17 x 60 + 4) only keeps
        // the ring moving: it waits for each stage (which also keeps it from running ahead of the barrier phases) and
        // hands it back, leaving its issue slots to the warps that have pixels.  The flag goes through a shuffle so
        // that the compiler sees a warp-uniform branch (the ring position stays in uniform registers).
        if (!__shfl_sync(0xffffffffu, xs0 < q.W ? 1 : 0, 0)) {
            for (int g = 0; g < T; ++g) {
                tma_mbar_wait_idle(bar0 + 8u * stage, phase);
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (S + stage)) : "memory");
                if (++stage == S) stage = 0, phase ^= 1u;
            }
        } else
        for (int g = 0; g < T; ++g) {
            // ---- phase 1: horizontal interpolation of the staged window for this warp's 2 columns ----
            tma_mbar_wait(bar0 + 8u * stage, phase);
            __syncwarp();  // every lane has read the previous pass's rows
            {
                const float* win = reinterpret_cast<const float*>(smem + kRowsBytes + kWtsBytes + stage * kStageBytes) + lane * (WS * 2);
                float2 v0[P1_ITERS], v1[P1_ITERS], v2[P1_ITERS];
#pragma unroll
                for (int k = 0; k < P1_ITERS; ++k) {
                    if (lane + 32 * k < ITEMS) {
                        const float* wk = win + k * 32 * (WS * 2);
                        v0[k] = *reinterpret_cast<const float2*>(wk + i0);
                        v1[k] = *reinterpret_cast<const float2*>(wk + i1);
                        v2[k] = *reinterpret_cast<const float2*>(wk + i2);
                    }
                }
                const float4 w0 = *reinterpret_cast<const float4*>(wts), w1 = *reinterpret_cast<const float4*>(wts + 4),
                             w2 = *reinterpret_cast<const float4*>(wts + 8);
#pragma unroll
                for (int k = 0; k < P1_ITERS; ++k) {
                    if (lane + 32 * k < ITEMS) {
                        const f32x2 V0 = {v0[k].x, v0[k].y}, V1 = {v1[k].x, v1[k].y}, V2 = {v2[k].x, v2[k].y};
                        const f32x2 p0 = fma2(f32x2{w0.x, w0.y}, V0, fma2(f32x2{w1.x, w1.y}, V1, mul2(f32x2{w2.x, w2.y}, V2)));
                        const f32x2 p1 = fma2(f32x2{w0.z, w0.w}, V0, fma2(f32x2{w1.z, w1.w}, V1, mul2(f32x2{w2.z, w2.w}, V2)));
                        // item i = (pair i / 6, source row i % 6) -> rows[pair][row][px][2]
                        const int i = lane + 32 * k;
                        *reinterpret_cast<float4*>(rows + (i / kUpRows) * 32 + (i % kUpRows) * 4) = make_float4(p0.x, p0.y, p1.x, p1.y);
                    }
                }
            }
            __syncwarp();
            if (lane == 0)  // this warp no longer reads the window: hand the slot back to the producer
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (S + stage)) : "memory");
            if (++stage == S) stage = 0, phase ^= 1u;
            // ---- phase 2: vertical interpolation -> the C logits of this lane's pixel, as class pairs ----
            f32x2 xl[CP];
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                const float2 t = *reinterpret_cast<const float2*>(rows + top_idx + k * 32);
                const float2 u = *reinterpret_cast<const float2*>(rows + bot_idx + k * 32);
                xl[k] = fma2(LY0, f32x2{t.x, t.y}, mul2(LY1, f32x2{u.x, u.y}));
            }
            const uint32_t vote = mc_pass_math_pairs<C, PROBS, VOTES>(xl, acc, ent);
            if (VOTES) {
                hist8[vote * NT + tid] += 1;
                if (g == 0) first_vote = vote;
            }
        }

        // ---- finalize the tile (same arithmetic as mc_finalize_kernel) ----
        float sc[DAS_N_SCORES][1];
#pragma unroll
        for (int k = 0; k < DAS_N_SCORES; ++k) sc[k][0] = 0.f;
        if (act) {
            const size_t o = (size_t)b * q.H * q.W + (size_t)y * q.W + x;
            bool valid[1];
            load_valid<C, 1>(f.labels, o, valid);
            if (PROBS)
                probs_scores<C, 1>([&](int c, float* a) { a[0] = (c & 1) ? acc[c >> 1].y : acc[c >> 1].x; }, &ent, (float)f.T, valid, sc);
            if (VOTES) {
                float ve[1];
                hist_vote_entropy<C, 1, NT>(hist8, lut, tid, ve);
                sc[DAS_SCORE_VOTE_ENTROPY][0] = valid[0] ? ve[0] : 0.f;
                if (f.weak_labels) store_weak_labels<1>(f.weak_labels, o, first_vote, valid);
            }
            store_maps<1>(f, o, sc, PROBS, VOTES);
        }
        block_partials<1, NT>(sc, red, f.partials + ((size_t)b * f.blocks_per_image + t_in) * DAS_N_SCORES, tid, sync);
        sync();  // `red` is reused by the next tile
    }
}

}  // namespace das
