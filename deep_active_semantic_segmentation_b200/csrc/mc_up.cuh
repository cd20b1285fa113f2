// Fused K1+K2 on the network's LOW-RESOLUTION logits: the model's last op,
//     x = F.interpolate(low_res_x, size=input.size()[2:], mode='bilinear', align_corners=True)   models/deeplab.py:59
// is applied inside the scoring kernel.  The [B,C,H,W] logits of a pass are never written to or
// read from HBM: the kernel reads [B,C,h,w] (16x fewer bytes at DeepLab's stride-4 decoder), so the step is no
// longer HBM bound but issue bound, and the interpolation is organised to cost as few issue slots as possible:
//
//   tile      16 x 16 output pixels per CTA; each of the 4 consumer warps owns a 16-row x 4-column strip, a lane
//             owns a horizontal pixel pair of one row (accumulators in registers, as in mc_tma.cuh)
//   producer  one warp copies the <= 6 x 6 low-res source window of every class into a shared-memory ring
//             (4-byte cp.async: no alignment demands, works for 129 x 129 planes) and signals an mbarrier
//   phase 1   a warp interpolates HORIZONTALLY, once per source row, the 4 columns of its strip:
//             rows[warp][c][r < 6][4 px] - shared by the ~4 output rows that lie between the same two source rows
//             (7 instructions per pixel pair and source row, 6/16 of them per output pixel pair).  The strip is
//             warp-private: __syncwarp() orders the two phases, there is no CTA barrier inside the pass loop
//   phase 2   a lane reads the two interpolated rows around its pixel pair (2 LDS.64) and interpolates
//             VERTICALLY with packed FMUL2 + FFMA2: 3 extra instructions per class and pass over the TMA kernel
//
// Arithmetic: indices and weights exactly as ATen's align_corners path (area_pixel_compute_scale /
// area_pixel_compute_source_index: scale = float(in-1)/float(out-1), src = scale*dst, i0 = int(src), l1 = src-i0,
// l0 = 1-l1), each of the three lerps as fma(l0, a, l1*b) - bit-identical to ATen's vectorised CPU kernel at
// DeepLab's shapes (oracle/restate.py: bilinear_upsample_align_corners).  Everything after the interpolation is
// the arithmetic of mc_kernels.cuh (mc_pass_math, probs_scores, byte histogram, fixed-order block partials).
#pragma once

namespace das {

constexpr int kUpTile = 16;                 // output tile edge
constexpr int kUpRows = 6;                  // source rows / columns a tile may touch (host-checked per shape)
constexpr int kUpCols = 6;                  // row stride of the staged window (floats): stride 6 keeps the
                                            // phase-1 reads of 16 (class,row) items x 2 columns on 32 distinct banks
constexpr int kUpStrip = 4;                 // columns per consumer warp
constexpr int kUpThreads = 160;             // 4 consumer warps + 1 producer warp
constexpr int kUpMaxStages = 8;

struct McUpParams {
    const float* lowres[DAS_MAX_PASS_GROUP];  // [B,C,h,w] per pass
    McFinParams fin;                          // outputs; HW = H*W; blocks_per_image = tiles_x * tiles_y
    int B, n_passes, stages;
    int h, w, H, W, tiles_x, tiles_y;
    float rh, rw;                             // float(h-1)/float(H-1), float(w-1)/float(W-1)  (0 when H / W == 1)
};

constexpr int up_ctas_per_sm(int C) { return C <= 24 ? 3 : 2; }
constexpr size_t up_stage_bytes(int C) { return (size_t)C * kUpRows * kUpCols * sizeof(float); }
constexpr size_t up_rows_bytes(int C) { return (size_t)C * kUpRows * kUpTile * sizeof(float); }

// ATen: source index and the weight of the upper neighbour for destination index d (align_corners=True)
__device__ __forceinline__ void up_source(float scale, int d, int n_in, int& i0, int& step, float& l0, float& l1) {
    const float src = __fmul_rn(scale, (float)d);
    i0 = min((int)src, n_in - 1);
    l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
    step = i0 < n_in - 1 ? 1 : 0;
}

template <int C, bool PROBS, bool VOTES>
__global__ void __launch_bounds__(kUpThreads, up_ctas_per_sm(C)) mc_score_up_kernel(const __grid_constant__ McUpParams q) {
    constexpr int NT = 128, VEC = 2;
    constexpr int CR = C * kUpRows;                         // (class, source row) pairs per tile
    constexpr int P1_ITERS = (CR + 15) / 16;                // phase-1 items per thread
    constexpr uint32_t kStageBytes = (uint32_t)(C * kUpRows * kUpCols * sizeof(float));
    constexpr uint32_t kRowsBytes = (uint32_t)(C * kUpRows * kUpTile * sizeof(float));
    extern __shared__ __align__(16) uint8_t smem[];         // [4 warps][C][6][4] interpolated rows | ring of [C][6][6] windows
    __shared__ uint64_t bars[2 * kUpMaxStages];
    __shared__ uint32_t hist32[VOTES ? C * NT * VEC / 4 + 1 : 1];
    __shared__ float lut[VOTES ? 256 : 1];
    __shared__ float red[DAS_N_SCORES][NT / 32];
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(hist32);

    const McFinParams& f = q.fin;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = q.stages, T = q.n_passes;
    const uint32_t ring0 = tma_smem_u32(smem) + kRowsBytes, bar0 = tma_smem_u32(bars);
    const int tiles_per_image = q.tiles_x * q.tiles_y;
    const int total_tiles = q.B * tiles_per_image;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s), "r"(32) : "memory");       // full
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (S + s)), "r"(4) : "memory");  // empty
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        // ===== producer warp: lane -> up to two slots (row, col) of the 6 x 6 window, one cp.async per class =====
        constexpr int kSlots = kUpRows * kUpCols;            // 36 floats per class: slot = row * 6 + col
        const int s0 = lane, s1 = lane + 32;
        const bool v1 = s1 < kSlots;
        const size_t plane = (size_t)q.h * q.w;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
            const int ty0 = (t_in / q.tiles_x) * kUpTile, tx0 = (t_in % q.tiles_x) * kUpTile;
            const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
            const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
            // clamped source coordinates: slots past the plane edge repeat the edge (never used with weight > 0)
            const uint32_t off0 = (uint32_t)(min(r_base + s0 / kUpCols, q.h - 1) * q.w + min(c_base + s0 % kUpCols, q.w - 1));
            const uint32_t off1 = (uint32_t)(min(r_base + s1 / kUpCols, q.h - 1) * q.w + min(c_base + s1 % kUpCols, q.w - 1));
            for (int g = 0; g < T; ++g) {
                tma_mbar_wait(bar0 + 8u * (S + stage), phase ^ 1u);
                const float* src = q.lowres[g] + (size_t)b * C * plane;
                const uint32_t dst = ring0 + stage * kStageBytes;
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(c * kSlots + s0) * 4u),
                                 "l"(src + off0)
                                 : "memory");
                    if (v1)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(c * kSlots + s1) * 4u),
                                     "l"(src + off1)
                                     : "memory");
                    src += plane;
                }
                // the lane's arrival fires when all of its copies above have landed
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar0 + 8u * stage) : "memory");
                if (++stage == S) stage = 0, phase ^= 1u;
            }
        }
        return;
    }

    // ===== consumers: warp -> columns [4 warp, 4 warp + 4) of the 16 x 16 tile; lane -> pixel pair (x, x+1) of row yy =====
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
    float* rows = reinterpret_cast<float*>(smem) + warp * (C * kUpRows * kUpStrip);  // warp-private
    const bool vec_ok = (q.W & 1) == 0;  // pixel pairs are 8-byte aligned in the maps
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_image, t_in = tile % tiles_per_image;
        const int ty0 = (t_in / q.tiles_x) * kUpTile, tx0 = (t_in % q.tiles_x) * kUpTile;
        const int y = ty0 + yy, x = tx0 + kUpStrip * warp + 2 * xp;
        bool act[VEC];
        act[0] = y < q.H && x < q.W;
        act[1] = y < q.H && x + 1 < q.W;
        // ---- geometry of this thread's pixels (ATen align_corners) ----
        const int r_base = min((int)__fmul_rn(q.rh, (float)ty0), q.h - 1);
        const int c_base = min((int)__fmul_rn(q.rw, (float)tx0), q.w - 1);
        int y0, ys;
        float ly0, ly1;
        up_source(q.rh, min(y, q.H - 1), q.h, y0, ys, ly0, ly1);
        const int top_idx = (y0 - r_base) * kUpStrip + 2 * xp;  // float index inside the interpolated rows of class 0
        const int bot_idx = top_idx + ys * kUpStrip;
        float lx0[VEC], lx1[VEC];
        int a_idx[VEC], b_idx[VEC];  // the two source columns of pixel j inside (class, row) item yy (float index)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int x0, xs;
            up_source(q.rw, min(x + j, q.W - 1), q.w, x0, xs, lx0[j], lx1[j]);
            a_idx[j] = yy * kUpCols + (x0 - c_base);
            b_idx[j] = a_idx[j] + xs;
        }
        const f32x2 LX0 = {lx0[0], lx0[1]}, LX1 = {lx1[0], lx1[1]}, LY0 = {ly0, ly0}, LY1 = {ly1, ly1};
        const int h_store = yy * kUpStrip + 2 * xp;

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
            // ---- phase 1: horizontal interpolation of the staged window, (class, row) items yy, yy+16, ... ----
            tma_mbar_wait(bar0 + 8u * stage, phase);
            __syncwarp();  // every lane has read the previous pass's rows
            {
                const float* win = reinterpret_cast<const float*>(smem + kRowsBytes + stage * kStageBytes);
#pragma unroll
                for (int k = 0; k < P1_ITERS; ++k) {
                    if (k * 16 + yy < CR) {
                        const float* wk = win + k * 16 * kUpCols;
                        const f32x2 a = {wk[a_idx[0]], wk[a_idx[1]]}, bb = {wk[b_idx[0]], wk[b_idx[1]]};
                        const f32x2 r = fma2(LX0, a, mul2(LX1, bb));
                        *reinterpret_cast<float2*>(rows + h_store + k * 16 * kUpStrip) = make_float2(r.x, r.y);
                    }
                }
            }
            __syncwarp();
            if (lane == 0)  // this warp no longer reads the window: hand the slot back to the producer
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (S + stage)) : "memory");
            if (++stage == S) stage = 0, phase ^= 1u;
            // ---- phase 2: vertical interpolation -> the C logits of this thread's two pixels ----
            float xl[C][VEC];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float2 t = *reinterpret_cast<const float2*>(rows + top_idx + c * kUpRows * kUpStrip);
                const float2 u = *reinterpret_cast<const float2*>(rows + bot_idx + c * kUpRows * kUpStrip);
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
int launch_score_up(const McUpParams& p, int flags, int ctas_per_sm, cudaStream_t st);
int dispatch_score_up(const McUpParams& p, int flags, int ctas_per_sm, cudaStream_t st);

}  // namespace das
