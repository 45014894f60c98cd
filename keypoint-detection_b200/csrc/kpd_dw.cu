// Batch-wide depthwise k x k convolution (+ folded BN + activation, optional SE squeeze), NHWC fp32, with TMA-staged tiles.
//   reference op: tv:models/mobilenetv3.py:92-104 (the depthwise Conv2dNormActivation of every InvertedResidual block)
//   and the AdaptiveAvgPool2d of its SqueezeExcitation (tv:ops/misc.py:243-251).
//
// Work item = (image, group of <= 32 channels, TH x TW output tile).  Persistent CTAs walk the items with a ring of
// shared-memory stages: ONE 4-D TMA box {CG channels, in_w, in_h, 1 image} brings the zero-padded input tile (halo
// included -- out-of-image coordinates are zero-filled by TMA, which implements the padding) and one 2-D box brings the
// group's k*k tap weights; both complete on the stage's mbarrier.  The loads of the next `stages` items are in flight
// while the 256 threads compute the current one, so there is no index arithmetic, register staging or exposed DRAM latency on
// the load side -- the tiled predecessor (dw_tiled_kernel, kpd_pw.cu) spent most of its time there (ncu: 24% warps active,
// <= 20% of DRAM bandwidth).  The compute phase is the same sliding-window FMA loop: per tap row the (XR-1)*stride+K input
// vectors and K weights are read from shared memory once and reused from registers for XR outputs.
#include "kpd_tma.cuh"
#include <stdlib.h>

namespace kpd {

struct DwArgs {
    const float* bias; float* out; float* pool;
    int B, H, Ho, C, stride, act;
    int TH, TW, tiles_x, tiles, CG, in_h, in_w;
    int stages, stage_bytes, tile_bytes, w_bytes, tx_bytes;   // tile_bytes = weight offset in the stage (128-aligned); tx_bytes = bytes the two boxes deliver
    int total;
};

constexpr int DW_THREADS = 256, DW_XR = 4;

template <int K>
__global__ void __launch_bounds__(DW_THREADS, 2) dw_tma_kernel(const __grid_constant__ CUtensorMap tmIn,
                                                            const __grid_constant__ CUtensorMap tmW, const DwArgs a) {
    extern __shared__ __align__(128) uint8_t dw_smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem) + 127) & ~uintptr_t(127));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)a.stages * a.stage_bytes);
    float* red = reinterpret_cast<float*>(bars + 8);           // [2][32] SE squeeze partials, double-buffered by item parity
    const int tid = threadIdx.x, lane = tid & 31;
    const int pad = (K - 1) / 2, CG = a.CG;
    const int step = gridDim.x;

    auto issue = [&](int item, int s) {                        // one thread: both TMA loads of `item` into stage s
        const int tile = item % a.tiles, rest = item / a.tiles;
        const int b = rest % a.B, grp = rest / a.B;
        const int ty0 = (tile / a.tiles_x) * a.TH, tx0 = (tile % a.tiles_x) * a.TW;
        const uint32_t bar = smem_u32(bars + s), dst = smem_u32(base + (size_t)s * a.stage_bytes);
        mbar_expect_tx(bar, (uint32_t)a.tx_bytes);
        tma_load_4d(dst, &tmIn, bar, grp * CG, tx0 * a.stride - pad, ty0 * a.stride - pad, b);
        tma_load_2d(dst + a.tile_bytes, &tmW, bar, grp * CG, 0);
    };

    if (tid == 0) {
        for (int s = 0; s < a.stages; s++) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) red[tid] = 0.f;
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < a.stages; s++) { const int item = blockIdx.x + s * step; if (item < a.total) issue(item, s); }
    }

    int n = 0;
    for (int item = blockIdx.x; item < a.total; item += step, n++) {
        const int s = n % a.stages;
        const uint32_t phase = (uint32_t)(n / a.stages) & 1u;
        // item decode (a handful of 32-bit divisions per item, amortised over the tile)
        const int tile = item % a.tiles, rest = item / a.tiles;
        const int b = rest % a.B, grp = rest / a.B;
        const int c0 = grp * CG, cg = min(CG, a.C - c0), cg4 = cg >> 2;       // cg4 is a power of two (host-checked)
        const int cg4_sh = 31 - __clz(cg4);
        const int ty0 = (tile / a.tiles_x) * a.TH, tx0 = (tile % a.tiles_x) * a.TW;
        const int th = min(a.TH, a.Ho - ty0), tw = min(a.TW, a.Ho - tx0);
        const int in_w = a.in_w;
        const float* tilep = reinterpret_cast<const float*>(base + (size_t)s * a.stage_bytes);     // [in_h][in_w][CG]
        const float* wsm = reinterpret_cast<const float*>(base + (size_t)s * a.stage_bytes + a.tile_bytes);   // [K*K][CG]
        const int c4 = tid & (cg4 - 1);                                       // fixed for the thread within an item
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + c4);
        float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f);
        float* outb = a.out + (size_t)b * a.Ho * a.Ho * a.C + c0;
        const int runs = (tw + DW_XR - 1) / DW_XR;
        const int nitems = (th * runs) << cg4_sh;

        mbar_wait(smem_u32(bars + s), phase);

        for (int it = tid; it < nitems; it += DW_THREADS) {
            const int q = it >> cg4_sh;
            const int oy = q / runs;
            const int ox0 = (q - oy * runs) * DW_XR;
            float4 acc[DW_XR];
#pragma unroll
            for (int j = 0; j < DW_XR; j++) acc[j] = bb;
            if (a.stride == 1) {
                constexpr int SEG = DW_XR - 1 + K;
#pragma unroll
                for (int ky = 0; ky < K; ky++) {
                    const float* rp = tilep + ((size_t)(oy + ky) * in_w + ox0) * CG + c4 * 4;
                    float4 seg[SEG], wk[K];
#pragma unroll
                    for (int i = 0; i < SEG; i++) seg[i] = *reinterpret_cast<const float4*>(rp + (size_t)i * CG);   // a partial last run over-reads inside the stage: unused lanes
#pragma unroll
                    for (int kx = 0; kx < K; kx++) wk[kx] = *reinterpret_cast<const float4*>(wsm + (ky * K + kx) * CG + c4 * 4);
#pragma unroll
                    for (int j = 0; j < DW_XR; j++)
#pragma unroll
                        for (int kx = 0; kx < K; kx++) {
                            acc[j].x = fmaf(seg[j + kx].x, wk[kx].x, acc[j].x); acc[j].y = fmaf(seg[j + kx].y, wk[kx].y, acc[j].y);
                            acc[j].z = fmaf(seg[j + kx].z, wk[kx].z, acc[j].z); acc[j].w = fmaf(seg[j + kx].w, wk[kx].w, acc[j].w);
                        }
                }
            } else {
                constexpr int SEG = (DW_XR - 1) * 2 + K;
#pragma unroll
                for (int ky = 0; ky < K; ky++) {
                    const float* rp = tilep + ((size_t)(oy * 2 + ky) * in_w + ox0 * 2) * CG + c4 * 4;
                    float4 seg[SEG], wk[K];
#pragma unroll
                    for (int i = 0; i < SEG; i++) seg[i] = *reinterpret_cast<const float4*>(rp + (size_t)i * CG);
#pragma unroll
                    for (int kx = 0; kx < K; kx++) wk[kx] = *reinterpret_cast<const float4*>(wsm + (ky * K + kx) * CG + c4 * 4);
#pragma unroll
                    for (int j = 0; j < DW_XR; j++)
#pragma unroll
                        for (int kx = 0; kx < K; kx++) {
                            acc[j].x = fmaf(seg[2 * j + kx].x, wk[kx].x, acc[j].x); acc[j].y = fmaf(seg[2 * j + kx].y, wk[kx].y, acc[j].y);
                            acc[j].z = fmaf(seg[2 * j + kx].z, wk[kx].z, acc[j].z); acc[j].w = fmaf(seg[2 * j + kx].w, wk[kx].w, acc[j].w);
                        }
                }
            }
#pragma unroll
            for (int j = 0; j < DW_XR; j++) {
                if (ox0 + j >= tw) break;
                float4 r;
                r.x = apply_act(acc[j].x, a.act); r.y = apply_act(acc[j].y, a.act); r.z = apply_act(acc[j].z, a.act); r.w = apply_act(acc[j].w, a.act);
                reinterpret_cast<float4*>(outb + ((size_t)(ty0 + oy) * a.Ho + tx0 + ox0 + j) * a.C)[c4] = r;
                ssum.x += r.x; ssum.y += r.y; ssum.z += r.z; ssum.w += r.w;
            }
        }
        float* redb = red + (n & 1) * 32;
        if (a.pool) {
            // lanes with equal (lane mod cg4) hold the same channel group: butterfly over the other lane bits, then one
            // shared atomic per (warp, channel)
            for (int o = 16; o >= cg4; o >>= 1) {
                ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, o); ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, o);
                ssum.z += __shfl_xor_sync(0xffffffffu, ssum.z, o); ssum.w += __shfl_xor_sync(0xffffffffu, ssum.w, o);
            }
            if (lane < cg4) {
                atomicAdd(&redb[lane * 4 + 0], ssum.x); atomicAdd(&redb[lane * 4 + 1], ssum.y);
                atomicAdd(&redb[lane * 4 + 2], ssum.z); atomicAdd(&redb[lane * 4 + 3], ssum.w);
            }
        }
        __syncthreads();                                      // stage s is free again; this item's squeeze partials are complete
        if (a.pool && tid < cg) { atomicAdd(a.pool + (size_t)b * a.C + c0 + tid, redb[tid]); redb[tid] = 0.f; }
        if (tid == 0) { const int nxt = item + a.stages * step; if (nxt < a.total) issue(nxt, s); }
    }
}

int depthwise_tma(const MBBlock& m, const float* in, float* out, float* pool, int B, int H, int sm_count, cudaStream_t s) {
    const int Ho = H / m.stride, C = m.exp, K = m.k;
    if (C & 3) { set_error("depthwise_tma: unsupported channel count %d", C); return KPD_ERR_UNSUPPORTED; }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    // channel group: <= 32 channels such that (group/4) is a power of two for every group incl. the last partial one
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    int CG = 32;
    if (C < 32) CG = C;
    else if ((C % 32) != 0 && !pow2((C % 32) / 4)) CG = 16;
    if (!pow2(CG / 4) || (C % CG != 0 && !pow2((C % CG) / 4))) { set_error("depthwise_tma: channel grouping failed for C=%d", C); return KPD_ERR_UNSUPPORTED; }
    static const int tile_kb = getenv("KPD_DW_TILE_KB") ? atoi(getenv("KPD_DW_TILE_KB")) : 40;
    static const int stages_env = getenv("KPD_DW_STAGES") ? atoi(getenv("KPD_DW_STAGES")) : 2;
    const int TW = std::min(Ho, 28);
    int TH = std::min(Ho, 28);
    while (TH > 1 && (size_t)((TH - 1) * m.stride + K) * ((TW - 1) * m.stride + K) * CG * 4 > (size_t)tile_kb * 1024) TH--;
    int tiles_y = (Ho + TH - 1) / TH;
    TH = (Ho + tiles_y - 1) / tiles_y;                    // balance the row tiles
    const int tiles_x = (Ho + TW - 1) / TW;
    DwArgs a{};
    a.bias = m.b_dw; a.out = out; a.pool = pool; a.B = B; a.H = H; a.Ho = Ho; a.C = C; a.stride = m.stride; a.act = m.act;
    a.TH = TH; a.TW = TW; a.tiles_x = tiles_x; a.tiles = tiles_x * tiles_y; a.CG = CG;
    a.in_h = (TH - 1) * m.stride + K; a.in_w = (TW - 1) * m.stride + K;
    a.tile_bytes = a.in_h * a.in_w * CG * 4;               // multiple of 64 (CG >= 16)
    a.w_bytes = K * K * CG * 4;
    if (a.in_h > 256 || a.in_w > 256 || CG < 16) { set_error("depthwise_tma: unsupported tile %dx%dx%d", a.in_h, a.in_w, CG); return KPD_ERR_UNSUPPORTED; }
    a.tx_bytes = a.tile_bytes + a.w_bytes;
    a.tile_bytes = (a.tile_bytes + 127) & ~127;             // weights start 128-byte aligned; the TMA box itself is dense
    // a partial last run reads up to (XR-1) pixels past the tile row: keep that inside the stage
    a.stage_bytes = (a.tile_bytes + a.w_bytes + DW_XR * m.stride * CG * 4 + 127) & ~127;
    a.stages = std::max(2, std::min(stages_env, 8));
    const int64_t total = (int64_t)a.tiles * B * ((C + CG - 1) / CG);
    if (total >= ((int64_t)1 << 31)) { set_error("depthwise_tma: too many work items"); return KPD_ERR_UNSUPPORTED; }
    a.total = (int)total;
    const size_t smem = (size_t)a.stages * a.stage_bytes + 128 /*align*/ + 64 /*barriers*/ + 256 /*red*/;
    if (smem > 227 * 1024) { set_error("depthwise_tma: tile does not fit in shared memory"); return KPD_ERR_UNSUPPORTED; }
    const int ctas_per_sm = std::max(1, std::min(4, (int)((227 * 1024) / (smem + 1024))));

    CUtensorMap tmIn, tmW;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)H * C * 4, (cuuint64_t)H * H * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)CG, (cuuint32_t)a.in_w, (cuuint32_t)a.in_h, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("depthwise_tma: cuTensorMapEncodeTiled(input) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(K * K)};
        cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {(cuuint32_t)CG, (cuuint32_t)(K * K)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(m.w_dw), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("depthwise_tma: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    const int grid = (int)std::min<int64_t>(total, (int64_t)sm_count * ctas_per_sm);
    static uint64_t attr3 = 0, attr5 = 0;
    if (K == 3) {
        KPD_CUDA(ensure_smem_attr(dw_tma_kernel<3>, 227 * 1024, attr3));
        dw_tma_kernel<3><<<grid, DW_THREADS, smem, s>>>(tmIn, tmW, a);
    } else if (K == 5) {
        KPD_CUDA(ensure_smem_attr(dw_tma_kernel<5>, 227 * 1024, attr5));
        dw_tma_kernel<5><<<grid, DW_THREADS, smem, s>>>(tmIn, tmW, a);
    } else { set_error("depthwise_tma: unsupported kernel size %d", K); return KPD_ERR_UNSUPPORTED; }
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
