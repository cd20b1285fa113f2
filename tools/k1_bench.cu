// Micro-benchmark of the fused K1+K2 kernel variants (vector width, accumulator home, blocks/SM) on the
// BASELINE config-2 shape.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include \
//        -I deep_active_semantic_segmentation_b200/csrc tools/k1_bench.cu -o /tmp/k1_bench && /tmp/k1_bench
#include <cstdio>
#include <vector>

#include "mc_kernels.cuh"

namespace das {
thread_local int g_last_cuda_error = 0;
std::atomic<unsigned long long> g_launch_count{0};
}  // namespace das
using namespace das;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void fill_logits(float* x, int B, int C, int H, int W, uint32_t seed) {
    const size_t n = (size_t)B * C * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int w = i % W, h = (i / W) % H, c = (i / ((size_t)W * H)) % C, b = i / ((size_t)W * H * C);
        const uint32_t cls = hash32((b * 131 + (h / 32)) * 977 + (w / 32)) % C;
        const float base = ((hash32((uint32_t)i * 2654435761u) >> 8) * (1.f / 8388608.f) - 1.f) * 1.7f;
        const float jit = ((hash32((uint32_t)i * 40503u + seed) >> 8) * (1.f / 8388608.f) - 1.f) * 1.2f;
        x[i] = base + jit + (c == (int)cls ? 3.f : 0.f);
    }
}

constexpr int C_ = 19, H_ = 512, W_ = 1024, T_ = 20;
static int g_iters = 5;

template <int VEC, bool SMEM, int MINB>
int run(const char* name, const McScoreParams& q0, int B, double alg_bytes, std::vector<float>& ref) {
    McScoreParams q = q0;
    const long long per_block = (long long)kAccThreads * VEC;
    q.fin.blocks_per_image = (int)((q.acc.HW + per_block - 1) / per_block);
    auto kern = mc_score_kernel<C_, VEC, true, true, SMEM, MINB>;
    const size_t smem = SMEM ? acc_smem_bytes(C_, VEC) : 0;
    if (smem > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAccThreads, smem));
    dim3 grid(q.fin.blocks_per_image, B);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) kern<<<grid, kAccThreads, smem>>>(q);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    const int iters = g_iters;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) kern<<<grid, kAccThreads, smem>>>(q);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    // checksum of the block partials (pred-entropy column) for cross-variant agreement
    std::vector<float> part((size_t)B * q.fin.blocks_per_image * DAS_N_SCORES);
    CK(cudaMemcpy(part.data(), q.fin.partials, part.size() * sizeof(float), cudaMemcpyDeviceToHost));
    double s[DAS_N_SCORES] = {0};
    for (size_t i = 0; i < part.size(); ++i) s[i % DAS_N_SCORES] += part[i];
    if (ref.empty()) ref.assign(s, s + DAS_N_SCORES);
    double maxrel = 0;
    for (int k = 0; k < DAS_N_SCORES; ++k) maxrel = fmax(maxrel, fabs(s[k] - ref[k]) / fmax(1e-30, fabs(ref[k])));
    printf("%-22s regs=%3d spill=%4zuB smem=%6zuB occ=%d blk/SM (%2d warps)  %.4f ms  %7.1f GB/s  %.3f of 6548  checksum_rel=%.1e\n",
           name, fa.numRegs, (size_t)fa.localSizeBytes, smem + fa.sharedSizeBytes, occ, occ * 4, ms, alg_bytes / ms / 1e6,
           alg_bytes / ms / 1e6 / 6548.2, maxrel);
    return 0;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 8;
    if (argc > 2) g_iters = atoi(argv[2]);
    const size_t HW = (size_t)H_ * W_;
    const size_t n = (size_t)B * C_ * HW;
    McScoreParams q{};
    for (int t = 0; t < T_; ++t) {
        float* p;
        CK(cudaMalloc(&p, n * sizeof(float)));
        fill_logits<<<148 * 8, 256>>>(p, B, C_, H_, W_, 1000 + t);
        q.acc.logits[t] = p;
    }
    float* labels;
    CK(cudaMalloc(&labels, B * HW * sizeof(float)));
    CK(cudaMemset(labels, 0, B * HW * sizeof(float)));
    float* partials;
    CK(cudaMalloc(&partials, (size_t)B * (HW / 128 + 1) * DAS_N_SCORES * sizeof(float)));
    q.acc.sum_p = nullptr; q.acc.sum_ent = nullptr; q.acc.votes = nullptr;
    q.acc.HW = HW; q.acc.C = C_; q.acc.T_cap = T_; q.acc.n_passes = T_; q.acc.pass_begin = 0;
    q.fin = McFinParams{};
    q.fin.labels = labels; q.fin.partials = partials; q.fin.HW = HW; q.fin.C = C_; q.fin.T_cap = T_; q.fin.T = T_;
    CK(cudaDeviceSynchronize());
    const double alg = (double)T_ * n * 4;
    std::vector<float> ref;
    printf("fused K1+K2, B=%d, C=%d, %dx%d, T=%d, algorithmic bytes/launch = %.3f GB\n", B, C_, H_, W_, T_, alg / 1e9);
    run<4, false, 1>("vec4 regs  minb1", q, B, alg, ref);
    run<4, true, 3>("vec4 smem  minb3", q, B, alg, ref);
    run<2, false, 4>("vec2 regs  minb4", q, B, alg, ref);
    run<2, true, 4>("vec2 smem  minb4", q, B, alg, ref);
    run<2, true, 5>("vec2 smem  minb5", q, B, alg, ref);
    run<2, true, 6>("vec2 smem  minb6", q, B, alg, ref);
    run<2, true, 7>("vec2 smem  minb7", q, B, alg, ref);
    run<1, false, 6>("vec1 regs  minb6", q, B, alg, ref);
    run<1, false, 8>("vec1 regs  minb8", q, B, alg, ref);
    return 0;
}
