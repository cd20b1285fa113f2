// Probe: does a 1-D tiled TMA load accept an innermost start coordinate that is not 16-byte aligned?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma1d_probe tma1d_probe.cu && ./tma1d_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int start, float* out) {
    __shared__ __align__(1024) float buf[256];
    __shared__ uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(b));
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(d),
                     "l"(&map), "r"(b), "r"(start)
                     : "memory");
    }
    __syncthreads();
    asm volatile(
        "{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b)
        : "memory");
    out[threadIdx.x] = buf[threadIdx.x];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int start = argc > 1 ? atoi(argv[1]) : 0;
    const size_t n = 100000;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *dsrc, *dout;
    cudaMalloc(&dsrc, n * 4);
    cudaMalloc(&dout, 1024);
    cudaMemcpy(dsrc, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[1] = {n}, strides[1] = {0};
    cuuint32_t box[1] = {256}, es[1] = {1};
    CUresult r = ((Enc)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, dsrc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d start=%d\n", (int)r, start);
    probe<<<1, 256>>>(map, start, dout);
    cudaError_t e = cudaDeviceSynchronize();
    float o[256];
    cudaMemcpy(o, dout, 1024, cudaMemcpyDeviceToHost);
    printf("sync: %s  out[0]=%g out[255]=%g (want %d, %d)\n", cudaGetErrorString(e), o[0], o[255], start, start + 255);
    return 0;
}
