// Input perturbation of the MC-noise selectors, drawn on the device:
//     noise = np.random.normal(0, 0.125, image_batch.shape); model(image_batch + noise)      mc_noise.py:26-27
// The reference draws the noise with numpy on the host and uploads it every pass; torch.normal on the device costs
// 28 us per pass for a batch of eight 513 x 513 images (two passes over the data + a generic generator) - as much as the
// scoring kernel.  Here: one pass, out = x + sigma * N(0,1), Philox4x32-10 (the counter-based generator of cuRAND /
// PyTorch) + Box-Muller, four normals per thread per counter, 128-bit loads and stores: HBM bound (read 4 + write 4 bytes
// per element).  The stream is a pure function of (seed, stream_id, element index): reproducible, independent per pass.
#include "das_common.cuh"

namespace das {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
}
// two uniforms in (0, 1] -> two independent N(0,1) (Box-Muller; |z| <= 6.7 with 32-bit uniforms)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (a + 1) / 2^32 in (0, 1]
    const float v = (float)b * 2.3283064365386963e-10f;           // [0, 1)
    const float r = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    z0 = r * c, z1 = r * s;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) add_gaussian_noise_kernel(const float* __restrict__ x, size_t n, float sigma,
                                                                 uint32_t k0, uint32_t k1, uint32_t s0, uint32_t s1,
                                                                 float* __restrict__ out) {
    const size_t groups = (n + 3) / 4;  // one Philox counter per group of four consecutive elements
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), s0, s1};
        philox4x32_10(c, k0, k1);
        float z[4];
        box_muller(c[0], c[1], z[0], z[1]);
        box_muller(c[2], c[3], z[2], z[3]);
        const size_t i = g * 4;
        if (VEC4 && i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4*>(x + i);
            *reinterpret_cast<float4*>(out + i) =
                make_float4(fmaf(sigma, z[0], v.x), fmaf(sigma, z[1], v.y), fmaf(sigma, z[2], v.z), fmaf(sigma, z[3], v.w));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + j < n) out[i + j] = fmaf(sigma, z[j], x[i + j]);
        }
    }
}

}  // namespace das

using namespace das;

extern "C" {

int das_add_gaussian_noise(das_handle* h, const float* x, size_t n, float sigma, uint64_t seed, uint64_t stream_id,
                           float* out, void* stream) {
    DAS_ENTER(h);
    if (x == nullptr || out == nullptr) return DAS_ERR_INVALID_ARG;
    if (n == 0) return DAS_OK;
    const size_t groups = (n + 3) / 4;
    const size_t want = (groups + 255) / 256, cap = (size_t)h->num_sms * 16;
    const int grid = (int)(want < cap ? want : cap);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), s0 = (uint32_t)stream_id, s1 = (uint32_t)(stream_id >> 32);
    if (aligned16(x) && aligned16(out))
        DAS_LAUNCH((add_gaussian_noise_kernel<true>), grid, 256, 0, (cudaStream_t)stream, x, n, sigma, k0, k1, s0, s1, out);
    else
        DAS_LAUNCH((add_gaussian_noise_kernel<false>), grid, 256, 0, (cudaStream_t)stream, x, n, sigma, k0, k1, s0, s1, out);
    DAS_CHECK_LAUNCH();
    return DAS_OK;
}

}  // extern "C"
