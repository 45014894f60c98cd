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
// (py,px), i.e. output pixels (4cy+py, 4cx+px); all rows then share the B operand.  For (A) the stem writes a tf32 copy of f0
// as 16 sub-position planes (space-to-depth by 4), so a tap's pixels are a dense box {16 ch, 16, 4, 2} of one plane (64-byte
// swizzled rows; a stride-4 TMA traversal of the plain layout works too but is request-bound); (B) loads the coarse
// neighbour tile as a plain box {32 ch, 16, 4, 2} (128-byte swizzled rows) 4 K-slabs per neighbour.  Two tiles (adjacent
// coarse column blocks) share every weight stage.  Warp roles, mbarrier rings and the TMEM double buffer are those of
// conv_umma_kernel (kpd_conv_umma.cu).
#include "kpd_tma.cuh"
#include <stdlib.h>

namespace kpd {

struct Fpn0Params {
    int B, NB, H0, W0, Hc, Wc;         // fine map H0 x W0 = 4 Hc x 4 Wc; NB = B rounded up to even (image extent of the TMA maps)
    int tiles_cx, tiles_cy, cxgroups;  // coarse tiles of 16 x 4; cxgroups = pairs of column tiles sharing weights
    int stages, total_groups, parts;   // parts = pooling partials per image = tiles_cy * tiles_cx * 16
    int mat_base[16];                  // first W2 matrix of sub-position sp = py*4 + px
    const float* bias; float* out; float* pool_part;
    int skip;                          // timing attribution only (KPD_FPN0_SKIP): 1 no part (A), 2 no part (B), 4 no output stores, 8 no pooling
};

constexpr int F0_EPI_WARPS = 8 /* 4 per tile of a group */, F0_CW = 16 /* channels per epilogue chunk */, F0_RS = F0_CW + 4 /* padded row of the epilogue transpose tile */, F0_THREADS = 64 + 32 * F0_EPI_WARPS, F0_MT = 2, F0_SLOT = 16384, F0_STAGE = (F0_MT + 1) * F0_SLOT;

__device__ __forceinline__ int f0_first(int p) { return p == 0 ? -1 : 0; }            // first coarse-neighbour offset along an axis
__device__ __forceinline__ int f0_count(int p) { return (p == 0 || p == 3) ? 2 : 1; }   // neighbours along an axis

__global__ void __launch_bounds__(F0_THREADS, 1)
fpn0_fused_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmL,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const Fpn0Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // (offset arithmetic on the __shared__ array, not an integer round trip: shared-memory accesses must compile to LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * F0_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);
    float* red = sbias + 128;                          // [8 warps][32][F0_RS] transpose tiles (16-byte aligned rows)
    float* wred = red + F0_EPI_WARPS * 32 * F0_RS;        // [2 tiles][4 lane groups][2][128] channel sum / max over a lane group's 32 rows

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
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), F0_EPI_WARPS); }
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
                // (A) 9 taps of the composed 16-channel conv, two taps per stage.  f0 is stored as 16 sub-position planes
                // [qy*4+qx][image][Hc][Wc][16] (space-to-depth by 4), so the pixels (4cy+py+dy, 4cx+px+dx) of a tile are a DENSE box
                // of plane ((py+dy)&3, (px+dx)&3) shifted by the coarse carry -- 1 KB contiguous per box row instead of a stride-4 gather
                for (int t0 = (p.skip & 1) ? 8 : 0; t0 < 9; t0 += 2) {
                    const int nt = min(2, 9 - t0);
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), (uint32_t)(nt * (F0_MT + 1) * 8192));
                    const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
                    for (int t = 0; t < nt; t++) {
                        const int tap = t0 + t, ty = py + tap / 3 - 1, tx = px + tap % 3 - 1;
                        const int plane = (ty & 3) * 4 + (tx & 3);
                        const int ny = ty < 0 ? -1 : (ty > 3 ? 1 : 0), nx = tx < 0 ? -1 : (tx > 3 ? 1 : 0);
#pragma unroll
                        for (int i = 0; i < F0_MT; i++)
                            tma_load_4d(sa + i * F0_SLOT + t * 8192, &tmF, full_bar(stage), 0, cx0 + 16 * i + nx, cy0 + ny, plane * p.NB + n0);
                        tma_load_2d(sa + F0_MT * F0_SLOT + t * 8192, &tmW1, full_bar(stage), tap * 16, 0);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                // (B) the coarse neighbours of this sub-position: 4 K-slabs of 32 channels each, 128-byte rows
                const int fy = f0_first(py), fx = f0_first(px), ncx = f0_count(px), nb = (p.skip & 2) ? 0 : f0_count(py) * ncx;
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
                const int nb = (p.skip & 2) ? 0 : f0_count(sp >> 2) * f0_count(sp & 3);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * F0_MT * 128);
                for (int t0 = (p.skip & 1) ? 8 : 0; t0 < 9; t0 += 2) {  // (A): two taps per stage, 64-byte rows, 2 K steps of 8 floats each
                    const int nt = min(2, 9 - t0);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * F0_STAGE);
                    for (int t = 0; t < nt; t++) {
                        const uint64_t bdesc = make_sdesc_rb(sa + F0_MT * F0_SLOT + t * 8192, 64);
#pragma unroll
                        for (int i = 0; i < F0_MT; i++) {
                            const uint64_t adesc = make_sdesc_rb(sa + i * F0_SLOT + t * 8192, 64);
#pragma unroll
                            for (int k4 = 0; k4 < 2; k4++)
                                tc_mma<false>(d_tmem + (uint32_t)(i * 128), adesc + 2 * k4, bdesc + 2 * k4, idesc, ((p.skip & 1) ? k4 : (t0 | t | k4)) != 0 ? 1u : 0u);
                        }
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
    } else {                                            // ---------------- epilogue: warps 2..5 take tile 0 of a group, warps 6..9 tile 1
        // (one warp per tile row block is a ~900-cycle dependent chain per 32-channel chunk -- tcgen05.ld, bias/ReLU, transpose,
        // stores, pooling -- so the two tiles of a group are drained side by side; measured 0.63 -> 0.35 ms of fixed cost)
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int ti = (warp - 2) >> 2;                 // which tile of the group
        const int row = lg * 32 + lane;
        const int cxl = row & 15, cyl = (row >> 4) & 3, nl = row >> 6;      // TMA box order: column fastest, then row, then image
        int acc = 0; uint32_t acc_phase = 0;
        for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
            int sp, cxg, cyt, pair; decode(grp, sp, cxg, cyt, pair);
            const int py = sp >> 2, px = sp & 3;
            mbar_wait_warp(tfull_bar(acc), acc_phase, lane);
            tc_fence_after();
            {
                const int cxt = cxg * F0_MT + ti;
                if (cxt < p.tiles_cx) {
                const int cx = cxt * 16 + cxl, cy = cyt * 4 + cyl, n = pair * 2 + nl;
                const bool valid = (cx < p.Wc) && (n < p.B);
                const size_t pix = ((size_t)n * p.H0 + (4 * cy + py)) * p.W0 + (4 * cx + px);
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((acc * F0_MT + ti) * 128);
                // The accumulators arrive one tile row (pixel) per thread.  Storing them row-per-thread sends every 16-byte piece
                // as its own half-empty 32-byte sector request (ncu: 104 M write requests for 1.64 GB, the L1->crossbar request
                // path the busiest unit of the kernel).  Instead each 32-channel chunk goes through a padded shared-memory tile
                // and is re-read as (row = lane/8, 4 channels = lane%8): one warp store then covers four full 128-byte lines, and
                // the same registers feed the pooling sums / maxima.
                const int n_w = pair * 2 + (lg >> 1), cy_w = cyt * 4 + ((lg & 1) << 1);     // image / first coarse row of this warp's 32 rows
                float* rt = red + (warp - 2) * (32 * F0_RS);
                float* wr = wred + ti * (8 * 128);
                // transposed read-back: lane = (row lane/4 of 8, 4 channels lane%4), 4 passes: a warp store = eight 64-byte runs (full
                // sectors).  16-channel chunks keep the transpose tiles at 20 KB, which buys the operand ring its fourth stage.
                const int rr8 = lane >> 2, q = lane & 3;
                for (int c0 = 0; c0 < 128; c0 += F0_CW) {
                    uint32_t r[F0_CW];
                    tc_ld16(taddr + c0, r);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < F0_CW / 4; j++) {
                        float4 v;
                        v.x = fmaxf(__uint_as_float(r[4 * j + 0]) + sbias[c0 + 4 * j + 0], 0.f);     // folded BN shift + ReLU
                        v.y = fmaxf(__uint_as_float(r[4 * j + 1]) + sbias[c0 + 4 * j + 1], 0.f);
                        v.z = fmaxf(__uint_as_float(r[4 * j + 2]) + sbias[c0 + 4 * j + 2], 0.f);
                        v.w = fmaxf(__uint_as_float(r[4 * j + 3]) + sbias[c0 + 4 * j + 3], 0.f);
                        *reinterpret_cast<float4*>(rt + lane * F0_RS + 4 * j) = v;
                    }
                    __syncwarp();
                    float4 sm4 = make_float4(0.f, 0.f, 0.f, 0.f), mx4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int it = 0; it < 4; it++) {
                        const int rr = it * 8 + rr8;                                         // tile row within this warp's 32
                        const int cx_r = cxt * 16 + (rr & 15), cy_r = cy_w + (rr >> 4);
                        const float4 x = *reinterpret_cast<const float4*>(rt + rr * F0_RS + 4 * q);
                        if (n_w < p.B && cx_r < p.Wc) {
                            if (!(p.skip & 8)) {                                             // values are >= 0 after the ReLU
                                sm4.x += x.x; sm4.y += x.y; sm4.z += x.z; sm4.w += x.w;
                                mx4.x = fmaxf(mx4.x, x.x); mx4.y = fmaxf(mx4.y, x.y); mx4.z = fmaxf(mx4.z, x.z); mx4.w = fmaxf(mx4.w, x.w);
                            }
                            if (!(p.skip & 4))
                                *reinterpret_cast<float4*>(p.out + (((size_t)n_w * p.H0 + (4 * cy_r + py)) * p.W0 + (4 * cx_r + px)) * 128 + c0 + 4 * q) = x;
                        }
                    }
#pragma unroll
                    for (int o = 4; o <= 16; o <<= 1) {                                      // combine the 8 row groups (lanes q, q+4, ..., q+28)
                        sm4.x += __shfl_xor_sync(0xffffffffu, sm4.x, o); sm4.y += __shfl_xor_sync(0xffffffffu, sm4.y, o);
                        sm4.z += __shfl_xor_sync(0xffffffffu, sm4.z, o); sm4.w += __shfl_xor_sync(0xffffffffu, sm4.w, o);
                        mx4.x = fmaxf(mx4.x, __shfl_xor_sync(0xffffffffu, mx4.x, o)); mx4.y = fmaxf(mx4.y, __shfl_xor_sync(0xffffffffu, mx4.y, o));
                        mx4.z = fmaxf(mx4.z, __shfl_xor_sync(0xffffffffu, mx4.z, o)); mx4.w = fmaxf(mx4.w, __shfl_xor_sync(0xffffffffu, mx4.w, o));
                    }
                    if (lane < 4) {
                        *reinterpret_cast<float4*>(wr + (lg * 2 + 0) * 128 + c0 + 4 * q) = sm4;
                        *reinterpret_cast<float4*>(wr + (lg * 2 + 1) * 128 + c0 + 4 * q) = mx4;
                    }
                    __syncwarp();
                }
                // rows 0..63 belong to image pair*2, rows 64..127 to the next one: combine lane groups {0,1} and {2,3} of this tile
                // (named barrier 1 + ti over the tile's 4 warps)
                asm volatile("bar.sync %0, 128;" ::"r"(1 + ti) : "memory");
                const int et = ((warp - 2) & 3) * 32 + lane;                // 0..127 -> channel
                const int part = (cyt * p.tiles_cx + cxt) * 16 + sp;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int n = pair * 2 + h;
                    if (n < p.B) {
                        float* dst = p.pool_part + ((size_t)n * p.parts + part) * 256;
                        dst[et] = wr[((2 * h) * 2 + 0) * 128 + et] + wr[((2 * h + 1) * 2 + 0) * 128 + et];
                        dst[128 + et] = fmaxf(wr[((2 * h) * 2 + 1) * 128 + et], wr[((2 * h + 1) * 2 + 1) * 128 + et]);
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + ti) : "memory");
                }
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

// f0t: tf32-rounded stem output as 16 sub-position planes [qy*4+qx][NB][S/8][S/8][16] (NB = B rounded up to even); lat1: tf32-rounded lateral level 1 [B][S/8][S/8][128] (top-down sum included);
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
    static const int skip_env = getenv("KPD_FPN0_SKIP") ? atoi(getenv("KPD_FPN0_SKIP")) : 0;
    p.skip = skip_env;
    const int misc = 1024 /*align*/ + 8 * (2 * 8 + 4) + 16 + 4 * 128 + 4 * (F0_EPI_WARPS * 32 * F0_RS + 2 * 8 * 128) + 64;
    int stages = (227 * 1024 - misc) / F0_STAGE; if (stages > 8) stages = 8;
    if (stages < 2) { set_error("fpn0_fused: not enough shared memory"); return KPD_ERR_UNSUPPORTED; }
    p.stages = stages;
    const size_t smem = (size_t)stages * F0_STAGE + misc;
    const cuuint64_t NB = (cuuint64_t)(pairs * 2);            // image extent rounded up to the pair the boxes span
    p.NB = pairs * 2;

    CUtensorMap tmF, tmL, tmW1, tmW2;
    {
        cuuint64_t dims[4] = {16, (cuuint64_t)p.Wc, (cuuint64_t)p.Hc, 16 * NB};       // [plane][image] folded into the outer dimension
        cuuint64_t strides[3] = {64, (cuuint64_t)p.Wc * 64, (cuuint64_t)p.Hc * p.Wc * 64};
        cuuint32_t box[4] = {16, 16, 4, 2};
        cuuint32_t es[4] = {1, 1, 1, 1};
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
