// Dispatch of the trunk's 1x1 convolutions (GEMMs over all B*H*W pixels) between the two implementations.
//   reference ops: every 1x1 conv of the trunk -- tv:models/mobilenetv3.py:87-108 (expand / project), features.12,
//   dll/models/backbone.py:33-37 (FPN laterals + nearest top-down add).
#include "kpd_common.cuh"
#include <stdlib.h>

namespace kpd {

int depthwise_tma(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s);

// Default (fp32 / bf16 modes): tcgen05 3xTF32 (kpd_pwu.cu), except for the shapes where its per-tile pipeline
// (TMA -> split -> MMA -> epilogue) measured slower than the CUDA-core kernel (kpd_pwf.cu): K <= 16 (one half-empty slab per
// tile), very narrow outputs with a residual, and the K = 24 lateral with the upsampled residual.  Strict fp32 always takes
// the exact FFMA kernel.
int pointwise_gemm(const PwMma& p, int sm_count, cudaStream_t s) {
    if (p.Wu && !p.exact) {
        // (the tcgen05 kernel stages the SE factors of a tile's images in shared memory: maps of < 32 pixels -- inputs below 160 x 160 --
        //  would need more than its 4 x 576 floats)
        const int hw = p.hw > 0 ? p.hw : 1;
        const bool se_fits = !p.se_scale || (int64_t)((127 + hw - 1) / hw + 1) * p.K <= 4 * 576;
        const bool prefer_ffma = (p.K <= 16) || (p.N <= 24 && p.res) || (p.res && p.res_mode == 2 && p.K <= 24) || !se_fits;
        if (!prefer_ffma) return pointwise_umma(p, sm_count, s);
    }
    if (!p.Wraw) { set_error("pointwise_gemm: missing weights"); return KPD_ERR_ARG; }
    return pointwise_ffma(p, sm_count, s);
}

// depthwise k x k of a block that is not fused with its expand GEMM (features.1 has none; strict mode; bands that do not fit)
int depthwise_batched(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s) {
    return depthwise_tma(m, in, out, pool, pool_capacity, pool_parts, B, H, sm_count, s);
}

}  // namespace kpd
