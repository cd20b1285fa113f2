// K4, tensor-core part: pairwise squared distances of the core-set features on the 5th-generation tensor
// cores (tcgen05.mma, BF16 operands staged by TMA, FP32 accumulators in TMEM), fused with the distance
// epilogue  d2~[c,i] = |f_i|^2 + |f_c|^2 - 2 <f_i, f_c>.
//
// Role in the k-center greedy (reference active_selection/core_set.py:17-38): the reference evaluates
// sklearn float64 distances of every pool row to every new centre.  Here the BF16 tensor-core distances
// are a FILTER with a proven error bound: a row's float64 min-distance can only change when
//      d2~(i,c) - delta * (|f_i|^2 + |f_c|^2)  <=  min_d2[i],
// and only those rows (a few per cent) are re-evaluated exactly in float64 (kcenter.cu).  Selections are
// therefore identical to the exact float64 path; the tensor cores remove ~97 % of its work.
//
// Error bound: bf16 round-to-nearest has relative error <= 2^-9 per operand, so every product is within
// 2^-8 (1 + 2^-10) of the exact one and, by Cauchy-Schwarz, |<a,b>~ - <a,b>| <= 2^-8 |a||b| + fp32
// accumulation error (<= K * 2^-24 |a||b| = 2^-13 |a||b| for K = 2048).  With 2|a||b| <= |a|^2 + |b|^2 the
// distance error is <= (2^-8 + 2^-13 + eps_f32) (|a|^2 + |b|^2); kFilterDelta = 2^-7 doubles that.
//
// Kernel anatomy (one CTA per SM, persistent over 128 x 256 output tiles):
//   warp 0   TMA producer: 4-stage ring of {A 128x64, B 256x64} bf16 tiles, 128-byte swizzle
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16)
//   warp 2   TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warps 4-7 epilogue: tcgen05.ld 32x32b -> registers -> distance -> coalesced stores of the TRANSPOSED
//            tile (out[c, i], i fastest), overlapped with the next tile's main loop
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "das_common.cuh"
#include "gram.cuh"
#include "tma_host.cuh"

namespace das {

constexpr int kBM = 128, kBN = 256, kBK = 64;  // tile; kBK bf16 = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kUmmaK = 16;
constexpr uint32_t kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2, kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 256;
constexpr uint32_t kTmemCols = 512;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major operand tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B:
// rows are 128 bytes, 8-row groups are 1024 bytes apart (SBO), version 1 (Blackwell), layout SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), 1
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                    // descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

// out[c * ld + i] = nrmA[i] + nrmB[c] - 2 <A_i, B_c>   for i < M (rows of A), c < Nn (rows of B)
__global__ void __launch_bounds__(kGemmThreads, 1)
    kc_dist_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const float* __restrict__ nrmA, const float* __restrict__ nrmB, float* __restrict__ out, int M,
                        int Nn, int ld, int num_kb) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + kStages * kStageBytes);
    // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kStages + 2 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (Nn + kBN - 1) / kBN;
    const int tiles = tiles_m * tiles_n;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(full_bar(s), 1), mbar_init(empty_bar(s), 1);
        for (int b = 0; b < 2; ++b) mbar_init(tfull_bar(b), 1), mbar_init(tempty_bar(b), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int tm = t % tiles_m, tn = t / tiles_m;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    mbar_expect_tx(full_bar(stage), kStageBytes);
                    tma_load_2d(sa, &tmA, full_bar(stage), kb * kBK, tm * kBM);
                    tma_load_2d(sb, &tmB, full_bar(stage), kb * kBK, tn * kBN);
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(buf), aphase ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * kBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in 16-byte address units
                        tc_mma_bf16(tmem_d, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));  // frees the smem stage when these MMAs retire
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                tc_commit(tfull_bar(buf));  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> distances -> transposed, coalesced stores =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const int tm = t % tiles_m, tn = t / tiles_m;
            const int buf = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(buf), aphase);
            tc_fence_after();
            const int i = tm * kBM + q * 32 + lane;  // row of A (fast index of the output)
            const float na = i < M ? nrmA[i] : 0.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kBN;
#pragma unroll 1
            for (int ch = 0; ch < kBN / 32; ++ch) {
                uint32_t v[32];
                tc_ld32(taddr + ch * 32, v);
                tc_wait_ld();
                const int c0 = tn * kBN + ch * 32;
                if (c0 < Nn) {
                    const float nbl = (c0 + lane) < Nn ? nrmB[c0 + lane] : 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float nb = __shfl_sync(0xffffffffu, nbl, j);
                        const int c = c0 + j;
                        if (i < M && c < Nn) out[(size_t)c * ld + i] = fmaf(-2.f, __uint_as_float(v[j]), na + nb);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05.mma.cta_group::2): two SMs of a cluster share one 256 x 256 output tile.
// Why: with one CTA per tile the tensor core reads 12 KB of operands per 128-cycle MMA from shared memory
// (96 B/clk) while TMA writes the same 96 B/clk into it - the 128 B/clk shared-memory port caps the tensor pipe
// near 70 % (measured 71 %).  In a pair each SM stages only its 128 rows of A and its 128 rows of B
// (32 KB per k-block instead of 48 KB) and the MMA reads 8 KB per SM: 64 + 64 B/clk.
//   both CTAs : warp 0 = TMA producer (own halves, transaction bytes land on the LEADER's full barrier),
//               warps 4-7 = epilogue of their own 128 accumulator rows
//   leader    : warp 1 = MMA issuer (M256 N256 K16), commits are multicast to both CTAs' barriers
// ---------------------------------------------------------------------------------------------
constexpr int k2Stages = 6;                       // 6 x 32 KB operand ring (measured: 3 stages 0.39 ms, 5 stages 0.33 ms)
constexpr uint32_t k2OutBytes = 32 * 128 * 4;     // epilogue staging: 32 centres x 128 rows fp32
constexpr int k2OutBufs = 1;                      // one staging buffer: the bulk store drains it in ~150 cycles
constexpr uint32_t k2HalfBytes = 128 * kBK * 2, k2StageBytes = 2 * k2HalfBytes;  // A half + B half = 32 KB
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    kc_dist_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ nrmA, const float* __restrict__ nrmB, float* __restrict__ out, int M,
                         int Nn, int ld, int num_kb) {
    constexpr int n_stages = k2Stages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    float* ostage = reinterpret_cast<float*>(smem_al + k2Stages * k2StageBytes);  // [k2OutBufs][32][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + k2Stages * k2StageBytes + k2OutBufs * k2OutBytes);
    // bars[0..S) full (the leader's are used), [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty (leader's)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k2Stages + 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (k2Stages + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * k2Stages + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * k2Stages + 2 + b); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int tiles_m = (M + 255) / 256, tiles_n = (Nn + 255) / 256;
    const int tiles = tiles_m * tiles_n;
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(full_bar(s), 1), mbar_init(empty_bar(s), 1);
        for (int b = 0; b < 2; ++b) mbar_init(tfull_bar(b), 1), mbar_init(tempty_bar(b), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs: barriers initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): own 128 rows of A and of B per k-block =====
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cid; t < tiles; t += ncl) {
                const int tm = t % tiles_m, tn = t / tiles_m;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * k2StageBytes, sb = sa + k2HalfBytes;
                    const uint32_t lead_full = full_bar(stage) & kPeerBitMask;
                    // one arrival (the leader's) + the transaction bytes of BOTH CTAs complete a phase.  The peer does not
                    // arrive: a release.cluster arrive per k-block costs a MEMBAR in the producer thread (measured: tensor
                    // pipe 37 %); its bytes can only land in the right phase because it refills a stage after `empty`.
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * k2StageBytes);
                    tma_load_2d_pair(sa, &tmA, lead_full, kb * kBK, tm * 256 + (int)rank * 128);
                    tma_load_2d_pair(sb, &tmB, lead_full, kb * kBK, tn * 256 + (int)rank * 128);
                    if (++stage == n_stages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // ===== MMA issuer (leader CTA only) =====
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = cid; t < tiles; t += ncl, ++it) {
                const int buf = it & 1;
                const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(buf), aphase ^ 1u);  // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * k2StageBytes, sb = sa + k2HalfBytes;
                    const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k)
                        tc_mma_bf16_pair(tmem_d, adesc + 2u * k, bdesc + 2u * k, kIdesc2, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(empty_bar(stage));  // frees this stage in BOTH CTAs
                    if (++stage == n_stages) stage = 0, phase ^= 1u;
                }
                tc_commit_pair(tfull_bar(buf));  // accumulator halves complete in both CTAs
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 rows of the 256 x 256 tile =====
        const int q = warp & 3;
        const uint32_t lead_tempty0 = map_to_cta(tempty_bar(0), 0), lead_tempty1 = map_to_cta(tempty_bar(1), 0);
        int it = 0;
        for (int t = cid; t < tiles; t += ncl, ++it) {
            const int tm = t % tiles_m, tn = t / tiles_m;
            const int buf = it & 1;
            const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(buf), aphase);
            tc_fence_after();
            const int i0 = tm * 256 + (int)rank * 128;  // first row of this CTA's half
            const int i = i0 + q * 32 + lane;
            const float na = i < M ? nrmA[i] : 0.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 256u;
#pragma unroll 1
            for (int ch = 0; ch < 256 / 32; ++ch) {
                uint32_t v[32];
                tc_ld32(taddr + ch * 32, v);
                tc_wait_ld();
                const int c0 = tn * 256 + ch * 32;
                if (c0 >= Nn || i0 >= M) continue;  // uniform over the four epilogue warps: nothing of this box is inside
                const int ob = ch % k2OutBufs;
                // the bulk store that last read this staging buffer must be done reading it
                if (warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(k2OutBufs - 1) : "memory");
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const float nbl = (c0 + lane) < Nn ? nrmB[c0 + lane] : 0.f;
                float* dst = ostage + ob * (32 * 128) + q * 32 + lane;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float nb = __shfl_sync(0xffffffffu, nbl, j);
                    dst[j * 128] = fmaf(-2.f, __uint_as_float(v[j]), na + nb);  // [centre j][row]: conflict-free
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (warp == 4 && lane == 0) {
                    // transposed tile [32 centres x 128 rows] -> out[c0.., i0..]; rows / centres beyond M / Nn are clipped
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmOut),
                                 "r"(smem_u32(ostage + ob * (32 * 128))), "r"(i0), "r"(c0)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(buf ? lead_tempty1 : lead_tempty0);
        }
        if (warp == 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all tiles are in HBM
    }

    tc_fence_before();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal or be signalled
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// bf16 copy (zero padded to Dp), float64 and float32 squared norms: one warp per row
__global__ void __launch_bounds__(256) kc_prepare_kernel(const float* __restrict__ feats, int N, int D, int Dp,
                                                         __nv_bfloat16* __restrict__ fb, double* __restrict__ nrm64,
                                                         float* __restrict__ nrm32) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* f = feats + (size_t)row * D;
    __nv_bfloat16* o = fb + (size_t)row * Dp;
    double s = 0.0;
    for (int k = lane; k < Dp; k += 32) {
        const float x = k < D ? f[k] : 0.f;
        o[k] = __float2bfloat16_rn(x);
        s = fma((double)x, (double)x, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
        nrm64[row] = s;
        nrm32[row] = (float)s;
    }
}

// bf16 [rows, Dp] row-major -> boxes of box_rows x 64 elements, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const __nv_bfloat16* base, int rows, int Dp, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Dp * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

KcFilterLayout kc_filter_layout(int N, int D, int rows) {
    KcFilterLayout L;
    L.Dp = (D + kBK - 1) / kBK * kBK;
    L.ld = (rows + 31) / 32 * 32;
    size_t off = 0;
    L.fb = off;
    off += align_up((size_t)N * L.Dp * sizeof(__nv_bfloat16), 1024);
    L.nrm64 = off;
    off += align_up((size_t)N * sizeof(double), 256);
    L.nrm32 = off;
    off += align_up((size_t)N * sizeof(float), 256);
    L.stats = off;
    off += 256;
    L.dt = off;
    off += align_up((size_t)N * L.ld * sizeof(float), 256);
    L.total = off;
    return L;
}

constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 + 256;
constexpr size_t kGemm2Smem = (size_t)k2Stages * k2StageBytes + k2OutBufs * k2OutBytes + 1024 + 256;

}  // namespace das

using namespace das;

extern "C" {

int das_kcenter_filter_bytes(int N, int D, int rows, size_t* bytes) {
    if (bytes == nullptr || N <= 0 || D <= 0 || rows <= 0 || rows > N) return DAS_ERR_INVALID_ARG;
    *bytes = kc_filter_layout(N, D, rows).total;
    return DAS_OK;
}

int das_kcenter_filter_build(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end, void* filter,
                             void* stream) {
    DAS_ENTER(h);
    if (feats == nullptr || filter == nullptr) return DAS_ERR_INVALID_ARG;
    if (N <= 0 || D <= 0 || row_begin < 0 || row_end > N || row_begin >= row_end) return DAS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(filter) & 1023u) != 0) return DAS_ERR_MISALIGNED;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = row_end - row_begin;
    const KcFilterLayout L = kc_filter_layout(N, D, rows);
    char* base = static_cast<char*>(filter);
    __nv_bfloat16* fb = reinterpret_cast<__nv_bfloat16*>(base + L.fb);
    double* nrm64 = reinterpret_cast<double*>(base + L.nrm64);
    float* nrm32 = reinterpret_cast<float*>(base + L.nrm32);
    float* dt = reinterpret_cast<float*>(base + L.dt);

    DAS_CUDA(cudaMemsetAsync(base + L.stats, 0, 256, st));
    DAS_LAUNCH(kc_prepare_kernel, (N + 7) / 8, 256, 0, st, feats, N, D, L.Dp, fb, nrm64, nrm32);
    DAS_CHECK_LAUNCH();

    CUtensorMap tmA, tmB;
    int rc = make_map(&tmA, fb + (size_t)row_begin * L.Dp, rows, L.Dp, kBM);  // A: this rank's rows (fast output index)
    if (rc != DAS_OK) return rc;
    // DAS_OPT_GEMM_2CTA = 0: one CTA per 128 x 256 tile (A/B measurements, tests of both kernels)
    if (h->opt[DAS_OPT_GEMM_2CTA]) {
        rc = make_map(&tmB, fb, N, L.Dp, 128);  // B: every row (the candidate centres), 128-row halves
        if (rc != DAS_OK) return rc;
        CUtensorMap tmOut;  // dt [N centres, ld] fp32, valid width = rows; boxes of 32 centres x 128 rows
        {
            const cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)N};
            const cuuint64_t strides[1] = {(cuuint64_t)L.ld * sizeof(float)};
            const cuuint32_t box[2] = {128, 32};
            rc = make_tensor_map(&tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != DAS_OK) return rc;
        }
        DAS_CUDA(cudaFuncSetAttribute(kc_dist_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemm2Smem));
        const int tiles = ((rows + 255) / 256) * ((N + 255) / 256);
        const int pairs = tiles < h->num_sms / 2 ? tiles : h->num_sms / 2;
        DAS_LAUNCH(kc_dist_gemm2_kernel, 2 * pairs, kGemmThreads, kGemm2Smem, st, tmA, tmB, tmOut, nrm32 + row_begin, nrm32, dt,
                   rows, N, L.ld, L.Dp / kBK);
        DAS_CHECK_LAUNCH();
        return DAS_OK;
    }
    rc = make_map(&tmB, fb, N, L.Dp, kBN);  // B: every row (the candidate centres)
    if (rc != DAS_OK) return rc;
    DAS_CUDA(cudaFuncSetAttribute(kc_dist_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    const int tiles = ((rows + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    DAS_LAUNCH(kc_dist_gemm_kernel, grid, kGemmThreads, kGemmSmem, st, tmA, tmB, nrm32 + row_begin, nrm32, dt, rows, N,
               L.ld, L.Dp / kBK);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

}  // extern "C"
