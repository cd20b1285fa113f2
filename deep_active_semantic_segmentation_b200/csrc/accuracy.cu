// Accuracy-predictor selectors (reference active_selection/accuracy.py): one pass over a [B,C,H,W] logits tensor
// and the float32 labels gives, per image,
//   wrong_count      sum over valid pixels of [label != argmax_c logits]            accuracy.py:30-33
//   p0_sum           sum over valid pixels of softmax(logits)[0]                    accuracy.py:55-58
//   not_argmax_sum   sum over valid pixels of (1 - argmax_c logits)                 accuracy.py:60-64
//   unsure_mean      mean over valid pixels of 4 p1 - 4 p1^2, p1 = softmax[1]       accuracy.py:117-118
//   valid_count      pixels with 0 <= label < num_classes                           accuracy.py:31,56,116
// and optionally the per-pixel map softmax[0] with invalid pixels set to 0 (accuracy.py:159-162), which then goes
// through the same suppress / box-sum / min-max / NMS tail as the vote-entropy maps.
// HBM bound and tiny next to K1 (C = 2 for the error-predictor head): logits are read once, class loop at run time.
#include <math.h>

#include "das_common.cuh"

namespace das {

constexpr int kAccBlock = 256;

__global__ void __launch_bounds__(kAccBlock) accuracy_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                             int C, long long HW, int num_classes, float* __restrict__ p0_map,
                                                             float* __restrict__ partials, int blocks_per_image) {
    __shared__ float red[DAS_ACC_N][kAccBlock / 32];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long pix = (long long)blockIdx.x * kAccBlock + tid;
    float v[DAS_ACC_N] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (pix < HW) {
        const float* x = logits + (size_t)b * C * HW + pix;
        float m = x[0];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            const float xc = x[(size_t)c * HW];
            if (xc > m) m = xc, arg = c;  // first maximal index (torch.argmax)
        }
        float s = 0.f, e0 = 0.f, e1 = 0.f;
        for (int c = 0; c < C; ++c) {
            const float e = ex2_approx((x[(size_t)c * HW] - m) * kLog2e);
            s += e;
            if (c == 0) e0 = e;
            if (c == 1) e1 = e;
        }
        const float inv = __frcp_rn(s);
        const float p0 = e0 * inv, p1 = e1 * inv;
        bool valid = true;
        float lab = 0.f;
        if (labels != nullptr) {
            lab = labels[(size_t)b * HW + pix];
            valid = (lab >= 0.f) && (lab < (float)num_classes);
        }
        if (valid) {
            v[DAS_ACC_WRONG_COUNT] = (labels != nullptr && lab != (float)arg) ? 1.f : 0.f;
            v[DAS_ACC_P0_SUM] = p0;
            v[DAS_ACC_NOT_ARGMAX_SUM] = 1.f - (float)arg;
            v[DAS_ACC_UNSURE_MEAN] = 4.f * p1 - 4.f * (p1 * p1);
            v[DAS_ACC_VALID_COUNT] = 1.f;
        }
        if (p0_map != nullptr) p0_map[(size_t)b * HW + pix] = valid ? p0 : 0.f;
    }
#pragma unroll
    for (int k = 0; k < DAS_ACC_N; ++k) {
        const float s = warp_sum(v[k]);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (tid < DAS_ACC_N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAccBlock / 32; ++w) s += red[tid][w];
        partials[((size_t)b * blocks_per_image + blockIdx.x) * DAS_ACC_N + tid] = s;
    }
}

// one warp per (image, score): fixed-order fp64 sum of the block partials
__global__ void accuracy_reduce_kernel(const float* partials, int blocks_per_image, float* image_scores) {
    __shared__ double tot[DAS_ACC_N];
    const int b = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    if (k < DAS_ACC_N)
        for (int i = lane; i < blocks_per_image; i += 32) s += (double)partials[((size_t)b * blocks_per_image + i) * DAS_ACC_N + k];
    s = warp_sum(s);
    if (k < DAS_ACC_N && lane == 0) tot[k] = s;
    __syncthreads();
    if (threadIdx.x < DAS_ACC_N) {
        const int j = threadIdx.x;
        float out = (float)tot[j];
        if (j == DAS_ACC_UNSURE_MEAN)  // torch mean over the masked selection: float32 sum / count, NaN when empty
            out = __fdiv_rn((float)tot[j], (float)tot[DAS_ACC_VALID_COUNT]);
        image_scores[(size_t)b * DAS_ACC_N + j] = out;
    }
}

}  // namespace das

using namespace das;

extern "C" {

int das_accuracy_workspace_bytes(int B, int H, int W, size_t* bytes) {
    if (bytes == nullptr || B <= 0 || H <= 0 || W <= 0) return DAS_ERR_INVALID_ARG;
    const long long HW = (long long)H * W;
    const size_t blocks = (size_t)((HW + kAccBlock - 1) / kAccBlock);
    *bytes = align_up((size_t)B * blocks * DAS_ACC_N * sizeof(float), 256);
    return DAS_OK;
}

int das_accuracy_scores(das_handle* h, const float* logits, int B, int C, int H, int W, const float* labels,
                        int num_classes, float* p0_map, float* image_scores, void* workspace, void* stream) {
    DAS_ENTER(h);
    if (logits == nullptr || image_scores == nullptr || workspace == nullptr) return DAS_ERR_INVALID_ARG;
    if (B <= 0 || C < 1 || H <= 0 || W <= 0 || num_classes < 1) return DAS_ERR_INVALID_ARG;
    if (B > 65535) return DAS_ERR_UNSUPPORTED;
    const long long HW = (long long)H * W;
    const int blocks = (int)((HW + kAccBlock - 1) / kAccBlock);
    cudaStream_t st = (cudaStream_t)stream;
    float* partials = static_cast<float*>(workspace);
    DAS_LAUNCH(accuracy_kernel, dim3(blocks, B), kAccBlock, 0, st, logits, labels, C, HW, num_classes, p0_map, partials, blocks);
    DAS_CHECK_LAUNCH();
    DAS_LAUNCH(accuracy_reduce_kernel, B, 32 * 8, 0, st, partials, blocks, image_scores);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

}  // extern "C"
