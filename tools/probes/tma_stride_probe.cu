// Probe: semantics of CUtensorMap elementStrides (traversal stride) for a 4-D tiled load.
// Tensor [N=2][H=16][W=32][C=16] fp32, value = n*100000 + h*1000 + w*10 + c/16.  Box {16, BW, BH, 2} with strides {1,4,4,1}.
// Prints how many bytes arrived (mbarrier completes only if expect_tx matches) and which source pixels landed per row.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int expect_bytes, int cx, int cy) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -1.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(expect_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(dst), "l"(&tm), "r"(0), "r"(cx), "r"(cy), "r"(0), "r"(b) : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64(); uint32_t ok = 0;
        while (!ok && clock64() - t0 < 200000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        out[16384] = ok ? 1.f : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
    const int N = 2, H = 16, W = 32, C = 16;
    std::vector<float> h((size_t)N * H * W * C);
    for (int n = 0; n < N; n++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) for (int c = 0; c < C; c++)
        h[(((size_t)n * H + y) * W + x) * C + c] = n * 100000.f + y * 1000.f + x * 10.f + c / 16.f;
    float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 16385 * 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    for (int variant = 0; variant < 2; variant++) {
        // variant 0: boxDim = traversal extent {16, 16, 8, 2} (expect 4x2 pixels);  variant 1: boxDim {16, 16*4.., } see below
        cuuint32_t box[4] = {16, (cuuint32_t)(variant == 0 ? 16 : 4), (cuuint32_t)(variant == 0 ? 8 : 2), 2};
        cuuint32_t es[4] = {1, 4, 4, 1};
        cuuint64_t dims[4] = {C, W, H, N};
        cuuint64_t strides[3] = {C * 4ull, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
        CUtensorMap tm;
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d encode -> %d\n", variant, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int expect = 16 * 4 * 2 * 2 * 4;      // 4 x 2 pixels x 2 images x 64 B
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        probe<<<1, 128, 65536>>>(tm, o, expect, 1, 2);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> res(16385);
        cudaMemcpy(res.data(), o, 16385 * 4, cudaMemcpyDeviceToHost);
        printf("variant %d: sync=%s barrier_completed=%g\n", variant, cudaGetErrorString(e), res[16384]);
        int rows = 0;
        for (int i = 0; i < 1024 && rows < 24; i++) {
            const float v = res[i * 16];
            if (v >= 0.f) { printf("  row %d: n=%d y=%d x=%d\n", i, (int)(v / 100000), ((int)v % 100000) / 1000, ((int)v % 1000) / 10); rows++; }
        }
        int filled = 0; for (int i = 0; i < 16384; i++) filled += res[i] >= 0.f;
        printf("  floats written: %d\n", filled);
    }
    return 0;
}
