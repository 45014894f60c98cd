// Throughput probe: 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) per SM sub-partition on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
// Each warp runs N independent accumulator chains; the kernel reports cycles per warp-instruction for 1..8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, long long* cycles, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 w = make_float2(1.0001f + threadIdx.x * 1e-9f, 0.9999f - threadIdx.x * 1e-9f), c = make_float2(0.5f + threadIdx.x * 1e-9f, 0.25f);   // per-thread values: register (not uniform / immediate) operands
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, w.x, c.x); a[i].y = fmaf(a[i].y, w.y, c.y); }   // 2 FFMA
                else a[i] = __ffma2_rn(a[i], w, c);                                                       // 1 FFMA2
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; mode++) {
            if (mode == 0) probe<0><<<1, warps * 32>>>(out, cyc, iters); else probe<1><<<1, warps * 32>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double fmas_per_thread = (double)iters * 4 * 8 * 2;                 // scalar FMAs per thread
            const double warp_fma_per_smsp = fmas_per_thread * (warps / 4.0);         // warp-wide scalar-FMA "rows" per sub-partition
            printf("%2d warps/CTA (%d per SMSP) %s: %lld cycles, %.2f cycles per 32-lane FMA per SMSP (%.1f lane-FMA/clk/SM)\n", warps, warps / 4,
                   mode ? "FFMA2" : "FFMA ", h, h / warp_fma_per_smsp, 4 * 32 * warp_fma_per_smsp / h);
        }
    }
    return 0;
}
