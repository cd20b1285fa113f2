// Fused K1+K2 on the network's LOW-RESOLUTION logits: the model's last op,
//     x = F.interpolate(low_res_x, size=input.size()[2:], mode='bilinear', align_corners=True)   models/deeplab.py:59
// is applied inside the scoring kernel.  The [B,C,H,W] logits of a pass are never written to or
// read from HBM: the kernel reads [B,C,h,w] (16x fewer bytes at DeepLab's stride-4 decoder), so the step is no
// longer HBM bound but issue bound, and the interpolation is organised to cost as few issue slots as possible:
//
//   tile      16 output rows x 4 NW columns per CTA (NW consumer warps: 15 -> one 512-thread CTA per SM, or 4 -> three
//             160-thread CTAs per SM for narrow outputs); each consumer warp owns a 16-row x 4-column strip, a lane a
//             horizontal pixel pair of one row (accumulators in registers, as in mc_tma.cuh)
//   producer  one warp copies the 6-row x (NW + 2)-column low-res source window of every class (clamped at the plane
//             edges) into a 4-stage shared-memory ring (4-byte cp.async: no alignment demands, works for 129 x 129
//             planes) and signals an mbarrier (cp.async.mbarrier.arrive.noinc)
//   phase 1   one lane per (class, source row) item interpolates HORIZONTALLY the 4 columns of its warp's strip from
//             3 window columns: out_j = W0[j] v0 + W1[j] v1 + W2[j] v2 with (W0,W1,W2) = (l0,l1,0) or (0,l0,l1) - no
//             per-pixel selects - and stores them with one STS.128 into rows[warp][class][row][4 px] (one 128-byte
//             line per class).  A source row is interpolated once and shared by the ~4 output rows between the same
//             two source rows: ~10 instructions per item, 114 items per 64 pixel pairs at C = 19.  The buffer is
//             warp-private: __syncwarp() orders the two phases, there is no CTA barrier inside the pass loop
//   phase 2   a lane reads the two interpolated rows around its pixel pair (2 LDS.64, one wavefront each) and
//             interpolates VERTICALLY with packed FMUL2 + FFMA2: 3 extra instructions per class and pass over the
//             TMA kernel
//
// Arithmetic: indices and weights exactly as ATen's align_corners path (area_pixel_compute_scale /
// area_pixel_compute_source_index: scale = float(in-1)/float(out-1), src = scale*dst, i0 = int(src), l1 = src-i0,
// l0 = 1-l1), each of the three lerps as fma(l0, a, l1*b) - bit-identical to ATen's vectorised CPU kernel at
// DeepLab's shapes (oracle/restate.py: bilinear_upsample_align_corners).  Everything after the interpolation is
// the arithmetic of mc_kernels.cuh (mc_pass_math, probs_scores, byte histogram, fixed-order block partials).
#pragma once

namespace das {

constexpr int kUpTileH = 16;                // output rows per tile
constexpr int kUpStrip = 4;                 // output columns per consumer warp
constexpr int kUpRows = 6;                  // source rows a tile may touch (host-checked per shape)
constexpr int kUpClassStride = 32;          // floats between the interpolated rows of two classes: one 128-byte line per
                                            // class, so the phase-2 LDS.64 of a warp is a single wavefront
constexpr int kUpMaxStages = 8;
// NW consumer warps per CTA -> tile of 16 rows x 4 NW columns; source window of 6 rows x (NW + 2) columns
__host__ __device__ constexpr int up_tile_w(int NW) { return kUpStrip * NW; }
__host__ __device__ constexpr int up_win_cols(int NW) { return NW + 2; }
__host__ __device__ constexpr int up_win_stride(int NW) { return (NW + 2) | 1; }  // odd: 32 consecutive (class,row) items hit 32 banks
constexpr int up_threads(int NW) { return 32 * (NW + 1); }
constexpr size_t up_stage_bytes(int C, int NW) { return (size_t)C * kUpRows * up_win_stride(NW) * sizeof(float); }
constexpr size_t up_rows_bytes(int C, int NW) { return (size_t)NW * C * kUpClassStride * sizeof(float); }
constexpr size_t up_wts_bytes(int NW) { return (size_t)NW * 12 * sizeof(float); }

struct McUpParams {
    const float* lowres[DAS_MAX_PASS_GROUP];  // [B,C,h,w] per pass
    McFinParams fin;                          // outputs; HW = H*W; blocks_per_image = tiles_x * tiles_y
    int B, n_passes, stages;
    int h, w, H, W, tiles_x, tiles_y;
    float rh, rw;                             // float(h-1)/float(H-1), float(w-1)/float(W-1)  (0 when H / W == 1)
    int num_sms;                              // host only: persistent grid sizing
};

// ATen: source index and the weights of the two neighbours for destination index d (align_corners=True)
__device__ __forceinline__ void up_source(float scale, int d, int n_in, int& i0, int& step, float& l0, float& l1) {
    const float src = __fmul_rn(scale, (float)d);
    i0 = min((int)src, n_in - 1);
    l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
    step = i0 < n_in - 1 ? 1 : 0;
}

template <int C, bool PROBS, bool VOTES, int NW, int MINB>
__global__ void __launch_bounds__(32 * (NW + 1), MINB) mc_score_up_kernel(const __grid_constant__ McUpParams q) {
    constexpr int NT = 32 * NW, VEC = 2;
    constexpr int CR = C * kUpRows;                         // (class, source row) items per tile
    constexpr int P1_ITERS = (CR + 31) / 32;                // phase-1 items per lane
    constexpr int WC = up_win_cols(NW), WS = up_win_stride(NW), TW = up_tile_w(NW);
    constexpr int kSlots = kUpRows * WC;                    // floats the producer copies per class
    constexpr int SLOT_ITERS = (kSlots + 31) / 32;
    constexpr uint32_t kStageBytes = (uint32_t)(C * kUpRows * WS * sizeof(float));
    constexpr uint32_t kRowsBytes = (uint32_t)(NW * C * kUpClassStride * sizeof(float));
    constexpr uint32_t kWtsBytes = (uint32_t)(NW * 12 * sizeof(float));
    // [NW][C][32] interpolated rows | [NW][12] strip weights | ring of [C][6][WS] windows
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[2 * kUpMaxStages];
    __shared__ uint32_t hist32[VOTES ? C * NT * VEC / 4 + 1 : 1];
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
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(32) : "memory");        // full
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (S + s)), "r"(NW) : "memory");  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        // ===== producer warp: lane -> slots (row, col) of the 6 x WC window, one 4-byte cp.async per class and slot =====
        const size_t plane = (size_t)q.h * q.w;
        uint32_t soff[SLOT_ITERS];  // position of the lane's slots inside a class of the staged window (floats)
#pragma unroll
        for (int j = 0; j < SLOT_ITERS; ++j) {
            const int s = lane + 32 * j;
            soff[j] = (uint32_t)((s / WC) * WS + s % WC);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
            const int ty0 = (t_in / q.tiles_x) * kUpTileH, tx0 = (t_in % q.tiles_x) * TW;
            const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
            const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
            // clamped source coordinates: slots past the plane edge repeat the edge (never used with weight > 0)
            uint32_t goff[SLOT_ITERS];
#pragma unroll
            for (int j = 0; j < SLOT_ITERS; ++j) {
                const int s = lane + 32 * j;
                goff[j] = (uint32_t)(min(r_base + s / WC, q.h - 1) * q.w + min(c_base + s % WC, q.w - 1));
            }
            for (int g = 0; g < T; ++g) {
                tma_mbar_wait(bar0 + 8u * (S + stage), phase ^ 1u);
                // The class loop is fully unrolled: the shared-memory offsets of a class are immediates and a copy costs one
                // 64-bit multiply-add for its address + the cp.async.  Rolled (24 instructions per class, 456 per pass from
                // ONE warp that gets a quarter of its scheduler) the producer was the critical path of the whole kernel:
                // votes-only scoring took as long as the full pass (profiles/r2_upsample_notes.md).
                const float* src = q.lowres[g] + (size_t)b * C * plane;
                uint32_t plane_g = (uint32_t)plane;
                asm volatile("" : "+r"(plane_g));  // per-pass value: keeps C x SLOT_ITERS hoisted offsets out of the registers
                uint32_t dst[SLOT_ITERS];
#pragma unroll
                for (int j = 0; j < SLOT_ITERS; ++j) dst[j] = ring0 + stage * kStageBytes + soff[j] * 4u;
#pragma unroll
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int j = 0; j < SLOT_ITERS; ++j)
                        if (lane + 32 * j < kSlots)  // element offsets inside an image fit 32 bits (host check)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst[j] + (uint32_t)(c * kUpRows * WS * 4)),
                                         "l"(src + (goff[j] + (uint32_t)c * plane_g))
                                         : "memory");
                }
                // the lane's arrival fires when all of its copies above have landed
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8u * stage) : "memory");
                if (++stage == S) stage = 0, phase ^= 1u;
            }
        }
        return;
    }

    // ===== consumers: warp -> columns [4 warp, 4 warp + 4) of the tile; lane -> pixel pair (x, x+1) of row yy =====
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
    float* rows = reinterpret_cast<float*>(smem) + warp * (C * kUpClassStride);                 // warp-private
    float* wts = reinterpret_cast<float*>(smem + kRowsBytes) + warp * 12;                       // warp-private
    // phase 1: lane -> items lane, lane + 32, ... of the C * 6 (class, source row) items
    int p1_store[P1_ITERS];
#pragma unroll
    for (int k = 0; k < P1_ITERS; ++k) {
        const int i = lane + 32 * k;
        p1_store[k] = (i / kUpRows) * kUpClassStride + (i % kUpRows) * kUpStrip;
    }
    const bool vec_ok = (q.W & 1) == 0;  // pixel pairs are 8-byte aligned in the maps
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
        const int ty0 = (t_in / q.tiles_x) * kUpTileH, tx0 = (t_in % q.tiles_x) * TW;
        const int xs0 = tx0 + kUpStrip * warp;  // first column of this warp's strip
        const int y = ty0 + yy, x = xs0 + 2 * xp;
        bool act[VEC];
        act[0] = y < q.H && x < q.W;
        act[1] = y < q.H && x + 1 < q.W;
        // ---- geometry (ATen align_corners) ----
        const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
        const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
        int y0, ys;
        float ly0, ly1;
        up_source(q.rh, min(y, q.H - 1), q.h, y0, ys, ly0, ly1);
        const int top_idx = (y0 - r_base) * kUpStrip + 2 * xp;  // float index inside the interpolated rows of class 0
        const int bot_idx = top_idx + ys * kUpStrip;
        // the strip's 4 pixels read source columns cb, cb+1, cb+2 of the window: pixel j = W0[j] v0 + W1[j] v1 + W2[j] v2
        // with (W0,W1,W2) = (l0,l1,0) or (0,l0,l1) - the zero weight adds an exact 0, so the value is ATen's
        // fma(l0, a, l1*b).  The weights are warp-uniform: computed by lanes 0..3, kept in shared memory.
        int cb;
        {
            int x0f, st;
            float t0, t1;
            up_source(q.rw, min(xs0, q.W - 1), q.w, x0f, st, t0, t1);
            cb = x0f - c_base;
            __syncwarp();
            if (lane < kUpStrip) {
                int x0;
                up_source(q.rw, min(xs0 + lane, q.W - 1), q.w, x0, st, t0, t1);
                const bool sh = x0 != x0f;  // this pixel starts one source column further right
                wts[lane] = sh ? 0.f : t0;
                wts[4 + lane] = sh ? t0 : t1;
                wts[8 + lane] = sh ? t1 : 0.f;
            }
            __syncwarp();
        }
        // never read past the window (the padding column holds garbage; columns past the plane edge repeat the edge)
        const int i1 = min(cb + 1, WC - 1), i2 = min(cb + 2, WC - 1);
        const f32x2 LY0 = {ly0, ly0}, LY1 = {ly1, ly1};

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
            // ---- phase 1: horizontal interpolation of the staged window for this warp's 4 columns ----
            tma_mbar_wait(bar0 + 8u * stage, phase);
            __syncwarp();  // every lane has read the previous pass's rows
            {
                const float* win = reinterpret_cast<const float*>(smem + kRowsBytes + kWtsBytes + stage * kStageBytes) + lane * WS;
                float v0[P1_ITERS], v1[P1_ITERS], v2[P1_ITERS];  // all loads first (the stores below may alias for the compiler)
#pragma unroll
                for (int k = 0; k < P1_ITERS; ++k) {
                    if (lane + 32 * k < CR) {
                        const float* wk = win + k * 32 * WS;
                        v0[k] = wk[cb], v1[k] = wk[i1], v2[k] = wk[i2];
                    }
                }
                const float4 w0 = *reinterpret_cast<const float4*>(wts), w1 = *reinterpret_cast<const float4*>(wts + 4),
                             w2 = *reinterpret_cast<const float4*>(wts + 8);
#pragma unroll
                for (int k = 0; k < P1_ITERS; ++k) {
                    if (lane + 32 * k < CR) {
                        const f32x2 V0 = {v0[k], v0[k]}, V1 = {v1[k], v1[k]}, V2 = {v2[k], v2[k]};
                        const f32x2 lo = fma2(f32x2{w0.x, w0.y}, V0, fma2(f32x2{w1.x, w1.y}, V1, mul2(f32x2{w2.x, w2.y}, V2)));
                        const f32x2 hi = fma2(f32x2{w0.z, w0.w}, V0, fma2(f32x2{w1.z, w1.w}, V1, mul2(f32x2{w2.z, w2.w}, V2)));
                        *reinterpret_cast<float4*>(rows + p1_store[k]) = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
            __syncwarp();
            if (lane == 0)  // this warp no longer reads the window: hand the slot back to the producer
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (S + stage)) : "memory");
            if (++stage == S) stage = 0, phase ^= 1u;
            // ---- phase 2: vertical interpolation -> the C logits of this lane's two pixels ----
            float xl[C][VEC];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float2 t = *reinterpret_cast<const float2*>(rows + top_idx + c * kUpClassStride);
                const float2 u = *reinterpret_cast<const float2*>(rows + bot_idx + c * kUpClassStride);
                const f32x2 v = fma2(LY0, f32x2{t.x, t.y}, mul2(LY1, f32x2{u.x, u.y}));
                xl[c][0] = v.x, xl[c][1] = v.y;
            }
            const uint32_t vote_word = mc_pass_math<C, VEC, PROBS, VOTES, false>(xl, acc, ent);
            if (VOTES) {
                hist_add<VEC, NT>(hist8, vote_word, tid);
                if (g == 0) first_vote = vote_word;
            }
        }

        // ---- finalize the tile (same arithmetic as mc_finalize_kernel) ----
        float sc[DAS_N_SCORES][VEC];
#pragma unroll
        for (int k = 0; k < DAS_N_SCORES; ++k)
#pragma unroll
            for (int j = 0; j < VEC; ++j) sc[k][j] = 0.f;
        if (act[0]) {
            const size_t o0 = (size_t)b * q.H * q.W + (size_t)y * q.W + x;
            bool valid[VEC];
            if (vec_ok) {
                load_valid<C, VEC>(f.labels, o0, valid);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    valid[j] = true;
                    if (f.labels != nullptr && act[j]) {
                        const float lab = f.labels[o0 + j];
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
            if (vec_ok) {  // W even: x + 1 < W whenever x < W
                if (VOTES && f.weak_labels) store_weak_labels<VEC>(f.weak_labels, o0, first_vote, valid);
                store_maps<VEC>(f, o0, sc, PROBS, VOTES);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if (act[j]) {
                        const size_t o = o0 + j;
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
        block_partials<VEC, NT>(sc, red, f.partials + ((size_t)b * f.blocks_per_image + t_in) * DAS_N_SCORES, tid, sync);
        sync();  // `red` is reused by the next tile
    }
}

template <int C>
int launch_score_up(const McUpParams& p, int flags, int nw, cudaStream_t st);
int dispatch_score_up(const McUpParams& p, int flags, int nw, cudaStream_t st);

}  // namespace das
