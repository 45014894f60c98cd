// FPN level 0 in composed form on tcgen05: fpn0 = ReLU(BN(conv3x3(lateral0))) with lateral0 = L0 f0 + up4(lat1) NEVER materialised.
//   reference ops: dll/models/backbone.py:33-39 (lateral_convs[0], the x4 nearest top-down add, fpn_convs[0] = conv3x3 + BN + ReLU)
//   and the global avg/max pooling of ChannelAttention (keypoint_model.py:33-44), reduced in the epilogue.
//
// conv3x3 is linear, so the two summands of lateral0 are convolved separately into the same TMEM accumulator:
//   (A) conv3x3(L0 f0)      = a 3x3 conv straight from the 16-channel stem output with composed weights
//                             W1[o][tap][i] = sum_c Wc[o][c][tap] L0[c][i]                       (K = 9 x 16 instead of 9 x 128)
//   (B) conv3x3(up4(lat1)): up4(lat1) is constant over 4x4 blocks, so an output pixel at sub-position (py,px) of its block
//                             sees at most 2 x 2 distinct coarse pixels; its 9 taps collapse to <= 4 summed 128x128 matrices
//                             W2[(py,px,ny,nx)] applied to lat1[cy+ny, cx+nx]                    (36 matrices instead of 16 x 9 taps)
// Per coarse pixel that is 36 + 18 = 54 units of 128x128 MACs instead of 144: 2.7x fewer MMAs and, since these fp32-operand
// MMAs are L2->SMEM bandwidth bound, 2.7x less operand traffic; the 1.64 GB lateral0 tensor (B=256) and its streaming kernel
// disappear.  Zero padding is preserved: taps outside the image read TMA out-of-bounds zeros in both forms.
//
// GEMM tiling: rows of an M=128 tile are the pixels {2 images} x {4 coarse rows} x {16 coarse columns} at ONE sub-position
// (py,px), i.e. output pixels (4cy+py, 4cx+px); all rows then share the B operand.  (A) loads them with a TMA traversal
// stride of 4 in x and y (box {16 ch, 64, 16, 2}, elementStrides {1,4,4,1}, 64-byte swizzled rows); (B) loads the coarse
// neighbour tile as a plain box {32 ch, 16, 4, 2} (128-byte swizzled rows) 4 K-slabs per neighbour.  Two tiles (adjacent
// coarse column blocks) share every weight stage.  Warp roles, mbarrier rings and the TMEM double buffer are those of
// conv_umma_kernel (kpd_conv_umma.cu).
#include "kpd_tma.cuh"
#include <stdlib.h>

namespace kpd {

struct Fpn0Params {
    int B, H0, W0, Hc, Wc;             // fine map H0 x W0 = 4 Hc x 4 Wc
    int tiles_cx, tiles_cy, cxgroups;  // coarse tiles of 16 x 4; cxgroups = pairs of column tiles sharing weights
    int stages, total_groups, parts;   // parts = pooling partials per image = tiles_cy * tiles_cx * 16
    int mat_base[16];                  // first W2 matrix of sub-position sp = py*4 + px
    const float* bias; float* out; float* pool_part;
};

constexpr int F0_THREADS = 192, F0_MT = 2, F0_SLOT = 16384, F0_STAGE = (F0_MT + 1) * F0_SLOT;

__device__ __forceinline__ int f0_first(int p) { return p == 0 ? -1 : 0; }            // first coarse-neighbour offset along an axis
__device__ __forceinline__ int f0_count(int p) { return (p == 0 || p == 3) ? 2 : 1; }   // neighbours along an axis

__global__ void __launch_bounds__(F0_THREADS, 1)
fpn0_fused_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmL,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const Fpn0Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * F0_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);
    float* red = sbias + 128;                          // [4 warps][32][33] transpose tiles
    float* wred = red + 4 * 32 * 33;                   // [4 warps][2][128] per-warp channel sum / max

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
    constexpr int TMEM_COLS = 512;                     // 2 accumulator stages x 2 tiles x 128 columns

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmF) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmL) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
        for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // group g -> (image pair, coarse row tile, column-tile pair, sub-position); the 16 sub-positions of a spatial tile are
    // consecutive, so their A operands (the same f0 window and lat1 tile) are served by L2
    auto decode = [&](int g, int& sp, int& cxg, int& cyt, int& pair) {
        sp = g & 15; int r = g >> 4;
        cxg = r % p.cxgroups; r /= p.cxgroups;
        cyt = r % p.tiles_cy; pair = r / p.tiles_cy;
    };

    if (warp == 0) {
        if (lane == 0) {                                // ---------------- TMA producer
            int stage = 0; uint32_t phase = 0;
            for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
                int sp, cxg, cyt, pair; decode(grp, sp, cxg, cyt, pair);
                const int py = sp >> 2, px = sp & 3, n0 = pair * 2, cy0 = cyt * 4, cx0 = cxg * (F0_MT * 16);
                // (A) 9 taps of the composed 16-channel conv: strided pixel box of f0, 64-byte rows
                int dy = -1, dx = -1;
                for (int tap = 0; tap < 9; tap++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), (uint32_t)((F0_MT + 1) * 8192));
                    const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
#pragma unroll
                    for (int i = 0; i < F0_MT; i++)
                        tma_load_4d(sa + i * F0_SLOT, &tmF, full_bar(stage), 0, 4 * (cx0 + 16 * i) + px + dx, 4 * cy0 + py + dy, n0);
                    tma_load_2d(sa + F0_MT * F0_SLOT, &tmW1, full_bar(stage), tap * 16, 0);
                    if (++dx == 2) { dx = -1; dy++; }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                // (B) the coarse neighbours of this sub-position: 4 K-slabs of 32 channels each, 128-byte rows
                const int fy = f0_first(py), fx = f0_first(px), ncx = f0_count(px), nb = f0_count(py) * ncx;
                for (int j = 0; j < nb; j++) {
                    const int ny = fy + j / ncx, nx = fx + j % ncx, mat = p.mat_base[sp] + j;
#pragma unroll
                    for (int kb = 0; kb < 4; kb++) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        mbar_expect_tx(full_bar(stage), (uint32_t)F0_STAGE);
                        const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
#pragma unroll
                        for (int i = 0; i < F0_MT; i++)
                            tma_load_4d(sa + i * F0_SLOT, &tmL, full_bar(stage), kb * 32, cx0 + 16 * i + nx, cy0 + ny, n0);
                        tma_load_2d(sa + F0_MT * F0_SLOT, &tmW2, full_bar(stage), kb * 32, mat * 128);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                // ---------------- MMA issuer
            // instruction descriptor: D=f32, A/B = tf32 (2), K-major both, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
                const int sp = grp & 15;
                const int nb = f0_count(sp >> 2) * f0_count(sp & 3);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * F0_MT * 128);
                for (int tap = 0; tap < 9; tap++) {                     // (A): 64-byte rows, 2 K steps of 8 floats
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
                    const uint64_t bdesc = make_sdesc_rb(sa + F0_MT * F0_SLOT, 64);
#pragma unroll
                    for (int i = 0; i < F0_MT; i++) {
                        const uint64_t adesc = make_sdesc_rb(sa + i * F0_SLOT, 64);
#pragma unroll
                        for (int k4 = 0; k4 < 2; k4++)
                            tc_mma<false>(d_tmem + (uint32_t)(i * 128), adesc + 2 * k4, bdesc + 2 * k4, idesc, (tap | k4) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                for (int kk = 0; kk < nb * 4; kk++) {                    // (B): 128-byte rows, 4 K steps
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
                    const uint64_t bdesc = make_sdesc_rb(sa + F0_MT * F0_SLOT, 128);
#pragma unroll
                    for (int i = 0; i < F0_MT; i++) {
                        const uint64_t adesc = make_sdesc_rb(sa + i * F0_SLOT, 128);
#pragma unroll
                        for (int k4 = 0; k4 < 4; k4++)
                            tc_mma<false>(d_tmem + (uint32_t)(i * 128), adesc + 2 * k4, bdesc + 2 * k4, idesc, 1u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {                                            // ---------------- epilogue warps 2..5
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const int cxl = row & 15, cyl = (row >> 4) & 3, nl = row >> 6;      // TMA box order: column fastest, then row, then image
        int acc = 0; uint32_t acc_phase = 0;
        for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
            int sp, cxg, cyt, pair; decode(grp, sp, cxg, cyt, pair);
            const int py = sp >> 2, px = sp & 3;
            mbar_wait_warp(tfull_bar(acc), acc_phase, lane);
            tc_fence_after();
            for (int ti = 0; ti < F0_MT; ti++) {
                const int cxt = cxg * F0_MT + ti;
                if (cxt >= p.tiles_cx) break;
                const int cx = cxt * 16 + cxl, cy = cyt * 4 + cyl, n = pair * 2 + nl;
                const bool valid = (cx < p.Wc) && (n < p.B);
                const size_t pix = ((size_t)n * p.H0 + (4 * cy + py)) * p.W0 + (4 * cx + px);
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((acc * F0_MT + ti) * 128);
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32(taddr + c0, r);
                    tc_wait_ld();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaxf(__uint_as_float(r[j]) + sbias[c0 + j], 0.f);      // folded BN shift + ReLU
                    if (valid) {
                        float4* o = reinterpret_cast<float4*>(p.out + pix * 128 + c0);
#pragma unroll
                        for (int q = 0; q < 8; q++) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                    // fused global avg/max pooling: per-warp transpose so that lane j reduces channel c0+j over the warp's 32 pixels
                    // (masked rows add 0; every value is >= 0 after the ReLU, so 0 never wins a max that has a valid row)
                    float* rt = red + lg * (32 * 33);
#pragma unroll
                    for (int j = 0; j < 32; j++) rt[lane * 33 + j] = valid ? v[j] : 0.f;
                    __syncwarp();
                    float sm_ = 0.f, mx_ = 0.f;
                    for (int rr = 0; rr < 32; rr++) { const float x = rt[rr * 33 + lane]; sm_ += x; mx_ = fmaxf(mx_, x); }
                    wred[(lg * 2 + 0) * 128 + c0 + lane] = sm_;
                    wred[(lg * 2 + 1) * 128 + c0 + lane] = mx_;
                    __syncwarp();
                }
                // rows 0..63 belong to image pair*2, rows 64..127 to the next one: combine warps {0,1} and {2,3}
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int et = (warp - 2) * 32 + lane;                      // 0..127 -> channel
                const int part = (cyt * p.tiles_cx + cxt) * 16 + sp;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int n = pair * 2 + h;
                    if (n < p.B) {
                        float* dst = p.pool_part + ((size_t)n * p.parts + part) * 256;
                        // warp index in wred = TMEM lane group (warp & 3): lane groups 0,1 hold rows 0..63, groups 2,3 rows 64..127
                        dst[et] = wred[((2 * h) * 2 + 0) * 128 + et] + wred[((2 * h + 1) * 2 + 0) * 128 + et];
                        dst[128 + et] = fmaxf(wred[((2 * h) * 2 + 1) * 128 + et], wred[((2 * h + 1) * 2 + 1) * 128 + et]);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// pooling partials per image written by the kernel for an S x S input
int fpn0_fused_parts(int S) { const int Hc = S / 8; return (Hc / 4) * ((Hc + 15) / 16) * 16; }

// f0t: tf32-rounded stem output [B][S/2][S/2][16]; lat1: tf32-rounded lateral level 1 [B][S/8][S/8][128] (top-down sum included);
// out: fpn0 [B][S/2][S/2][128] fp32; pool_part: [B][fpn0_fused_parts(S)][2][128].  Buffers must extend one image past an odd B.
int fpn0_fused(const kpd_plan* plan, const float* f0t, const float* lat1, float* out, float* pool_part, int B, int S, cudaStream_t s) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    if (S % 32 != 0 || !plan->fpn0c_w1 || !plan->fpn0c_w2) { set_error("fpn0_fused: unsupported size %d or missing composed weights", S); return KPD_ERR_UNSUPPORTED; }
    Fpn0Params p{};
    p.B = B; p.H0 = S / 2; p.W0 = S / 2; p.Hc = S / 8; p.Wc = S / 8;
    p.tiles_cx = (p.Wc + 15) / 16; p.tiles_cy = p.Hc / 4; p.cxgroups = (p.tiles_cx + F0_MT - 1) / F0_MT;
    p.parts = p.tiles_cy * p.tiles_cx * 16;
    const int pairs = (B + 1) / 2;
    const int64_t groups = (int64_t)pairs * p.tiles_cy * p.cxgroups * 16;
    if (groups >= ((int64_t)1 << 31)) { set_error("fpn0_fused: too many tiles"); return KPD_ERR_UNSUPPORTED; }
    p.total_groups = (int)groups;
    int m = 0;
    for (int sp = 0; sp < 16; sp++) {
        p.mat_base[sp] = m;
        const int py = sp >> 2, px = sp & 3;
        m += ((py == 0 || py == 3) ? 2 : 1) * ((px == 0 || px == 3) ? 2 : 1);
    }
    p.bias = plan->fpn_conv[0].bias; p.out = out; p.pool_part = pool_part;
    const int misc = 1024 /*align*/ + 8 * (2 * 8 + 4) + 16 + 4 * 128 + 4 * (4 * 32 * 33 + 8 * 128) + 64;
    int stages = (227 * 1024 - misc) / F0_STAGE; if (stages > 8) stages = 8;
    if (stages < 2) { set_error("fpn0_fused: not enough shared memory"); return KPD_ERR_UNSUPPORTED; }
    p.stages = stages;
    const size_t smem = (size_t)stages * F0_STAGE + misc;
    const cuuint64_t NB = (cuuint64_t)(pairs * 2);            // image extent rounded up to the pair the boxes span

    CUtensorMap tmF, tmL, tmW1, tmW2;
    {
        cuuint64_t dims[4] = {16, (cuuint64_t)p.W0, (cuuint64_t)p.H0, NB};
        cuuint64_t strides[3] = {64, (cuuint64_t)p.W0 * 64, (cuuint64_t)p.H0 * p.W0 * 64};
        cuuint32_t box[4] = {16, 64, 16, 2};                  // traversal extents: every 4th pixel -> 16 x 4 pixels x 2 images
        cuuint32_t es[4] = {1, 4, 4, 1};
        CUresult r = enc(&tmF, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(f0t), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fpn0_fused: cuTensorMapEncodeTiled(f0) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[4] = {128, (cuuint64_t)p.Wc, (cuuint64_t)p.Hc, NB};
        cuuint64_t strides[3] = {512, (cuuint64_t)p.Wc * 512, (cuuint64_t)p.Hc * p.Wc * 512};
        cuuint32_t box[4] = {32, 16, 4, 2};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(lat1), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fpn0_fused: cuTensorMapEncodeTiled(lat1) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[2] = {144, 128};
        cuuint64_t strides[1] = {144 * 4};
        cuuint32_t box[2] = {16, 128};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmW1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plan->fpn0c_w1), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fpn0_fused: cuTensorMapEncodeTiled(W1) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[2] = {128, 36 * 128};
        cuuint64_t strides[1] = {512};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmW2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plan->fpn0c_w2), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("fpn0_fused: cuTensorMapEncodeTiled(W2) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    static uint64_t attr = 0;
    KPD_CUDA(ensure_smem_attr(fpn0_fused_kernel, 227 * 1024, attr));
    const int grid = std::min(p.total_groups, plan->sm_count);
    fpn0_fused_kernel<<<grid, F0_THREADS, smem, s>>>(tmF, tmL, tmW1, tmW2, p);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
