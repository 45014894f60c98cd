// TMA / mbarrier PTX wrappers and the cuTensorMapEncodeTiled entry point, shared by the tcgen05 convolution
// (kpd_conv_umma.cu) and the TMA-staged depthwise convolution (kpd_dw.cu).
#pragma once
#include "kpd_common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace kpd {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait (about one second of SM clocks): a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0; long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ll) { printf("kpd: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
        }
    }
}
// Epilogue-side wait: one lane polls with back-off so the idle epilogue warps do not steal issue slots from the
// single-thread TMA producer / MMA issuer that share their SM sub-partitions; the rest of the warp parks at __syncwarp.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
    if (lane == 0) {
        uint32_t spins = 0; long long t0 = 0;
        while (!mbar_try_wait(bar, parity)) {
            __nanosleep(128);
            if ((++spins & 1023u) == 0) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 2000000000ll) { printf("kpd: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
            }
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace kpd
