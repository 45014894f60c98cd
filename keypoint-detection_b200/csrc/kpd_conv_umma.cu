// Dense 3x3 / 1x1 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//   reference ops replaced: dll/models/backbone.py:39 fpn_convs[0] (conv3x3 128->128 + BN + ReLU),
//   dll/models/heatmap_head.py:49-68,31-39 (conv3x3 64->256, 256->256, 256->64, each + bias + BN + ReLU).
//
// GEMM view: D[M=128 pixels][N=Cout] += A[M][K] * B[N][K]^T with K = taps*Cin.
//   * A is never materialised: for every tap the [128 pixels x BK channels] operand tile is ONE 4-D TMA box
//     {BK ch, tw, th, tn} of the NHWC input, shifted by the tap offset; out-of-image coordinates are zero-filled
//     by TMA, which implements the padding.  The box lands in shared memory as 128 rows of 128 bytes with the
//     128-byte swizzle, i.e. exactly the canonical K-major UMMA operand layout.
//   * B (BN-folded weights, [Cout][taps*Cin] K-major) is a 2-D TMA box {BK, Cout}.
//   * one elected thread issues tcgen05.mma (M=128, N=Cout, K=32 bytes per instruction), accumulators live in TMEM
//     (2 accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1), tcgen05.commit releases
//     shared-memory stages / publishes accumulators through mbarriers.
//   * persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc),
//     warps 2.. epilogue: tcgen05.ld -> +bias -> ReLU -> bf16/fp32 rows into a 128B-swizzled shared-memory staging tile ->
//     one TMA tensor store per (warp, 128 bytes of channels): full-line writes, out-of-image pixels clipped by the TMA unit
//     (a row-per-thread st.global epilogue issues 32 half-empty sector requests per instruction and made the K = 576
//     conv 64->256 epilogue-bound at 2.7x its MMA time, profiles/r02_conv1_sass_summary_before_tma_store.txt).
#include "kpd_tma.cuh"
#include <stdlib.h>

namespace kpd {

struct ConvParams {
    int N, H, W, Cin, Cout, taps;      // N = allocated instances (tensor-map extent)
    int tw, th, tn;                    // pixel box of one 128-row tile: tw*th*tn == 128
    int tiles_w, tiles_h, tiles_n;
    int act, out_bf16, stages, round_out_tf32;
    int mt;                            // M-tiles (128 pixels each) that share one weight stage: 1, 2 or 4 (2*mt*Cout <= 512 TMEM columns)
    int row_bytes;                     // bytes of K per shared-memory row: 128 (default) or 64 (Cin*elt == 64, e.g. the 16-ch fp32 lateral)
    const float* res; int res_up;      // optional residual added in the epilogue: nearest-upsampled [N][H/up][W/up][Cout] fp32 (FPN top-down)
    float* pool_part;                  // optional: per-tile channel sum/max partials [n][tiles_h*tiles_w][2][Cout] (tn must be 1)
    int out_cm;                        // 1: write [N][Cout/E16][H][W][E16] (16-byte channel chunks, the halo kernel's input layout)
    const float* bias;
    void* out;                         // NHWC [N,H,W,Cout], bf16 or fp32
    const int32_t* n_valid;            // optional device count of live instances (<= N)
    // optional fused tail of the heat-map head (Cout == 64): final_layer[3] 1x1 conv 64 -> 17 + bias + sigmoid
    // (heatmap_head.py:40-44,111) evaluated on the fp32 accumulators after bias + ReLU and written as NCHW fp32 maps
    // [slot][17][H][W]; the conv's own NHWC output is then not stored at all.
    const float* fin_w; const float* fin_b;   // [64][17], [17]
    float* fin_out; const int32_t* fin_slot;  // heat maps; output slot of each instance
};
constexpr int FIN_K = 17, FIN_KP = 20;   // joints; padded row of the shared-memory copy (float4 broadcast loads)

// Store 32 consecutive output channels (c0..c0+31) of pixel (n,h,w): NHWC bf16 / fp32, or chunk-major (16-byte chunks).
__device__ __forceinline__ void epi_store(const ConvParams& p, const float (&v)[32], int n, int h, int w, size_t pix, int c0) {
    if (p.out_cm) {                 // chunk-major: 16-byte chunk q of channel block c0 -> plane (c0*elt/16 + q)
        const size_t plane = (size_t)p.H * p.W;
        if (p.out_bf16) {
            uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)n * (p.Cout >> 3) + (c0 >> 3)) * plane + (size_t)h * p.W + w;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 u;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
                u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
                u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
                o[(size_t)q * plane] = u;
            }
        } else {
            float4* o = reinterpret_cast<float4*>(p.out) + ((size_t)n * (p.Cout >> 2) + (c0 >> 2)) * plane + (size_t)h * p.W + w;
#pragma unroll
            for (int q = 0; q < 8; q++) o[(size_t)q * plane] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    } else if (p.out_bf16) {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + c0);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
            o[q] = u;
        }
    } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.Cout + c0);
#pragma unroll
        for (int q = 0; q < 8; q++) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

constexpr int UMMA_THREADS = 192;
constexpr int TS_WARP_BYTES = 4096;      // TMA-store staging: 32 pixel rows x 128 bytes of channels per epilogue warp, double-buffered

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// MT = M-tiles (128 pixels each) sharing one weight stage, RB = bytes of K per shared-memory row (128, or 64 for the
// 16-channel fp32 lateral); both compile-time so that the single MMA-issuing thread runs a fully unrolled, lean loop
// (its per-instruction overhead directly limits tensor throughput).
// EW = epilogue warps: 4 (one per TMEM lane group) or 8 (two per lane group, each draining half of the output channels --
// the epilogue of a tile is a dependent chain of tcgen05.ld -> bias/ReLU -> convert -> stores per 32-channel chunk, and with
// K as short as 576 (head conv 64->256) one warp per lane group cannot drain a tile in the time the MMAs need for the next).
// TS = TMA-store epilogue (plain convs: no fused pooling / residual / final 1x1, NHWC output).
// CG = 2: CTA PAIRS (tcgen05 cta_group::2).  The two CTAs of a cluster work on two M-tiles at once; each loads its own A tile and
// HALF of the weight tile (Cout/2 rows), one thread of the leader CTA issues M = 256 MMAs that read both CTAs' shared memory and
// write both CTAs' tensor memory.  The tensor core fetches its operands from shared memory at ~64 B/clk; at N = 256 that is
// exactly what one MMA needs per its math time (A 4 KB + B 8 KB per 182 cycles), so any hiccup starves it -- a pair needs
// 4 + 4 KB per CTA (profiles/r02_conv_row_shift_tap_reuse.txt).  TMA loads of both CTAs complete on the leader's `full` barrier,
// tcgen05.commit is multicast to both CTAs' `empty` / `tmem_full` barriers, the peer's epilogue warps arrive on the leader's
// `tmem_empty` through the cluster address window.  (MT = 1 and the TMA-store epilogue only: the 256-wide head convs.)
template <bool BF16, int MT, int RB, int EW, bool TS, int CG>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x (A | B)] | store staging (TS) | barriers | tmem slot | bias
    // (offset arithmetic on the __shared__ array, not an integer round trip: the bias / weight reads of the epilogue must compile
    //  to LDS -- as generic loads they were the top stall of the epilogue warps)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int A_BYTES = 128 * RB;
    const int B_BYTES = (p.Cout / CG) * RB;            // CG = 2: this CTA's half of the weight rows
    const int STAGE_BYTES = MT * A_BYTES + B_BYTES;
    uint8_t* stage_out = smem + (size_t)p.stages * STAGE_BYTES;                 // [EW][2][TS_WARP_BYTES], 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + (TS ? EW * 2 * TS_WARP_BYTES : 0));
    // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);
    float* red = sbias + p.Cout;                      // pooling only: [4 warps][32][33] transpose tiles
    float* wred = red + 4 * 32 * 33;                  // pooling only: [4 warps][2][Cout] per-warp channel sum / max
    float* finw = sbias + p.Cout;                     // fused final 1x1 only: [64][FIN_KP] + bias [FIN_KP] (16-byte aligned: every block before it is)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;      // 0 = leader of the pair
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };

    int tmem_cols = 32; while (tmem_cols < 2 * MT * p.Cout) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (TS) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CG * EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = p.bias[i];
    if (p.fin_w) {
        for (int i = threadIdx.x; i < 65 * FIN_KP; i += blockDim.x) {
            const int c = i / FIN_KP, k = i - c * FIN_KP;
            finw[i] = k < FIN_K ? (c < 64 ? p.fin_w[c * FIN_K + k] : p.fin_b[k]) : 0.f;
        }
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();      // pair: both CTAs' barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // work split: a pair walks pair-groups of CG tiles; the CTA takes tile pg * CG + rank (a short last group repeats the last tile)
    const int wstart = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, wstep = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    const int n_live = p.n_valid ? min(*p.n_valid, p.N) : p.N;
    const int tiles_n_live = (n_live + p.tn - 1) / p.tn;
    const int total_tiles = tiles_n_live * p.tiles_h * p.tiles_w;
    const int total_groups = CG == 2 ? (total_tiles + 1) / 2 : (total_tiles + MT - 1) / MT;      // MT consecutive tiles (or the pair's two) share every weight stage
    constexpr int BK = RB / (BF16 ? 2 : 4);            // elements per K slab (one swizzled row)
    const int cchunks = p.Cin / BK;
    const int kiters = p.taps * cchunks;
    constexpr int ksteps = RB / 32;                    // MMAs per slab (32 bytes of K each)
    const int ksz = p.taps == 9 ? 3 : 1, pad = p.taps == 9 ? 1 : 0;

    if (warp == 0) {
        if (lane == 0) {                                // ---------------- TMA producer
            int stage = 0; uint32_t phase = 0;
            for (int grp = wstart; grp < total_groups; grp += wstep) {
                int w0[MT], h0[MT], n0[MT];             // tile origins: decoded once per group, not per K step
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    const int tile = min(CG == 2 ? grp * 2 + (int)crank : grp * MT + i, total_tiles - 1);         // a short last group repeats its last tile
                    const int iw = tile % p.tiles_w, ih = (tile / p.tiles_w) % p.tiles_h, in = tile / (p.tiles_w * p.tiles_h);
                    w0[i] = iw * p.tw; h0[i] = ih * p.th; n0[i] = in * p.tn;
                }
                int tap = 0, cc = 0, dy = -pad, dx = -pad;
                for (int kk = 0; kk < kiters; kk++) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    if constexpr (CG == 2) {
                        // the leader arms its barrier for both CTAs' bytes; either CTA's loads complete on it
                        if (crank == 0) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)STAGE_BYTES);
                        const uint32_t lbar = mapa_shared(full_bar(stage), 0);
                        tma_load_4d_cg2(sa, &tmA, lbar, cc * BK, w0[0] + dx, h0[0] + dy, n0[0]);
                        tma_load_2d_cg2(sa + A_BYTES, &tmB, lbar, tap * p.Cin + cc * BK, (int)crank * (p.Cout >> 1));
                    } else {
                    mbar_expect_tx(full_bar(stage), (uint32_t)STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < MT; i++)
                        tma_load_4d(sa + i * A_BYTES, &tmA, full_bar(stage), cc * BK, w0[i] + dx, h0[i] + dy, n0[i]);
                    tma_load_2d(sa + MT * A_BYTES, &tmB, full_bar(stage), tap * p.Cin + cc * BK, 0);
                    }
                    if (++cc == cchunks) { cc = 0; tap++; if (++dx == ksz - pad) { dx = -pad; dy++; } }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {                  // ---------------- MMA issuer (pairs: the leader CTA only)
            // instruction descriptor: D=f32 [4,6)=1, A/B format [7,10)/[10,13) (1 = bf16, 2 = tf32), K-major both,
            // N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t fmt = BF16 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Cout >> 3) << 17) | (((uint32_t)(128 * CG) >> 4) << 24);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int grp = wstart; grp < total_groups; grp += wstep) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * p.Cout);
                for (int kk = 0; kk < kiters; kk++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint64_t bdesc = make_sdesc_rb(sa + MT * A_BYTES, RB);
#pragma unroll
                    for (int i = 0; i < MT; i++) {
                        const uint64_t adesc = make_sdesc_rb(sa + i * A_BYTES, RB);
#pragma unroll
                        for (int k4 = 0; k4 < ksteps; k4++) { // 32-byte K steps inside the swizzled row
                            if constexpr (CG == 2) tc_mma_cg2<BF16>(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kk | k4) != 0 ? 1u : 0u);
                            else tc_mma<BF16>(d_tmem + (uint32_t)(i * p.Cout), adesc + 2 * k4, bdesc + 2 * k4, idesc, (kk | k4) != 0 ? 1u : 0u);
                        }
                    }
                    if constexpr (CG == 2) tc_commit_cg2(empty_bar(stage)); else
                    tc_commit(empty_bar(stage));        // frees this smem stage once the MMAs above retire
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                if constexpr (CG == 2) tc_commit_cg2(tfull_bar(acc)); else
                tc_commit(tfull_bar(acc));              // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {                                            // ---------------- epilogue warps 2..5
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        // EW == 8: two warps per lane group.  They split the M-tiles of a group (fused final 1x1: MT = 4, two tiles each) or
        // the output channels (TMA-store epilogue)
        const int whalf = (warp - 2) >> 2;
        const bool split_tiles = EW == 8 && !TS;
        const int cbeg = (EW == 8 && TS) ? whalf * (p.Cout >> 1) : 0, cend = (EW == 8 && TS) ? cbeg + (p.Cout >> 1) : p.Cout;
        const int ti_beg = split_tiles ? whalf * (MT / 2) : 0, ti_end = split_tiles ? ti_beg + MT / 2 : MT;
        const int row = lg * 32 + lane;
        const int rw = row % p.tw, rh = (row / p.tw) % p.th, rn = row / (p.tw * p.th);
        // TMA-store epilogue: this warp's 32 accumulator rows are the pixel box {tw, 32 / tw} of instance rn0 at row offset rh0
        const int rh0 = ((lg * 32) / p.tw) % p.th, rn0 = (lg * 32) / (p.tw * p.th);
        uint8_t* sbuf = stage_out + (size_t)(warp - 2) * 2 * TS_WARP_BYTES;
        uint32_t nstore = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int grp = wstart; grp < total_groups; grp += wstep) {
          mbar_wait_warp(tfull_bar(acc), acc_phase, lane);
          tc_fence_after();
          for (int ti = ti_beg; ti < ti_end; ti++) {
            const int tile = CG == 2 ? grp * 2 + (int)crank : grp * MT + ti;
            if (tile >= total_tiles) break;
            const int iw = tile % p.tiles_w, ih = (tile / p.tiles_w) % p.tiles_h, in = tile / (p.tiles_w * p.tiles_h);
            const int w = iw * p.tw + rw, h = ih * p.th + rh, n = in * p.tn + rn;
            const bool valid = (w < p.W) && (h < p.H) && (n < n_live);
            const size_t pix = ((size_t)n * p.H + h) * p.W + w;
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((acc * MT + ti) * p.Cout);
            if constexpr (TS) {
                // 128 bytes of output channels (64 bf16 / 32 fp32) per pixel row -> swizzled staging rows -> one tensor store
                constexpr int SC = BF16 ? 64 : 32;
                for (int c0 = cbeg; c0 < cend; c0 += SC) {
                    uint8_t* buf = sbuf + (nstore & 1u) * TS_WARP_BYTES;
                    // the store that read this buffer two chunks ago has drained it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    uint8_t* rowp = buf + lane * 128;
#pragma unroll
                    for (int hh = 0; hh < SC / 32; hh++) {
                        uint32_t r[32];
                        tc_ld32(taddr + c0 + 32 * hh, r);
                        tc_wait_ld();
                        const float4* b4 = reinterpret_cast<const float4*>(sbias + c0 + 32 * hh);
                        float v[32];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 bb = b4[q];
                            v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bb.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb.y;
                            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb.w;
                        }
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            float x = p.act == ACT_RELU ? fmaxf(v[j], 0.f) : v[j];
                            v[j] = p.round_out_tf32 ? tf32_rna(x) : x;
                        }
                        if constexpr (BF16) {
#pragma unroll
                            for (int q = 0; q < 4; q++) {          // 16-byte chunk 4*hh + q of the row: channels 8q .. 8q+7 of this half
                                uint4 u;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
                                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
                                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
                                u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
                                u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
                                *reinterpret_cast<uint4*>(rowp + (((4 * hh + q) ^ (lane & 7)) << 4)) = u;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; q++)
                                *reinterpret_cast<float4*>(rowp + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA unit
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&tmO, smem_u32(buf), c0, iw * p.tw, ih * p.th + rh0, in * p.tn + rn0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    nstore++;
                }
                continue;
            }
            if (p.fin_w) {
                // fused final_layer[3] + sigmoid: the thread owns one pixel, its 64 conv outputs arrive as two 32-column TMEM
                // loads, the 17 joint logits accumulate in registers (weights = shared-memory float4 broadcasts) and leave as
                // 17 NCHW stores (a warp covers 4 rows x 8 pixels of every joint plane: full 32-byte sectors)
                float a[FIN_KP];
#pragma unroll
                for (int k = 0; k < FIN_KP; k++) a[k] = finw[64 * FIN_KP + k];
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32(taddr + c0, r);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float x = fmaxf(__uint_as_float(r[j]) + sbias[c0 + j], 0.f);
                        const float4* wp = reinterpret_cast<const float4*>(finw + (c0 + j) * FIN_KP);
#pragma unroll
                        for (int q = 0; q < FIN_KP / 4; q++) {
                            const float4 w4 = wp[q];
                            a[4 * q + 0] = fmaf(x, w4.x, a[4 * q + 0]); a[4 * q + 1] = fmaf(x, w4.y, a[4 * q + 1]);
                            a[4 * q + 2] = fmaf(x, w4.z, a[4 * q + 2]); a[4 * q + 3] = fmaf(x, w4.w, a[4 * q + 3]);
                        }
                    }
                }
                if (valid) {
                    float* o = p.fin_out + ((size_t)p.fin_slot[n] * FIN_K * p.H + h) * p.W + w;
                    const size_t plane = (size_t)p.H * p.W;
#pragma unroll
                    for (int k = 0; k < FIN_K; k++) o[k * plane] = 1.f / (1.f + expf(-a[k]));
                }
                continue;
            }
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                uint32_t r[32];
                tc_ld32(taddr + c0, r);
                tc_wait_ld();
                float v[32];
                if (p.res && valid) {                 // + nearest-upsampled coarser lateral (backbone.py:35-37)
                    const float4* rp = reinterpret_cast<const float4*>(
                        p.res + (((size_t)n * (p.H / p.res_up) + h / p.res_up) * (p.W / p.res_up) + w / p.res_up) * p.Cout + c0);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 t4 = __ldg(rp + q);
                        r[4 * q + 0] = __float_as_uint(__uint_as_float(r[4 * q + 0]) + t4.x); r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + t4.y);
                        r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + t4.z); r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + t4.w);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    float x = __uint_as_float(r[j]) + sbias[c0 + j];
                    x = p.act == ACT_RELU ? fmaxf(x, 0.f) : x;
                    v[j] = p.round_out_tf32 ? tf32_rna(x) : x;
                }
                if (valid) epi_store(p, v, n, h, w, pix, c0);
                if (p.pool_part) {
                    // fused global avg/max pooling (keypoint_model.py:37,40): per-warp transpose through shared memory so
                    // that lane j reduces channel c0+j over this warp's 32 pixels, from the fp32 accumulators
                    float* rt = red + lg * (32 * 33);
#pragma unroll
                    for (int j = 0; j < 32; j++) rt[lane * 33 + j] = valid ? v[j] : 0.f;
                    __syncwarp();
                    float sm_ = 0.f, mx_ = -INFINITY;
                    for (int rr = 0; rr < 32; rr++) { const float x = rt[rr * 33 + lane]; sm_ += x; mx_ = fmaxf(mx_, x); }
                    wred[(lg * 2 + 0) * p.Cout + c0 + lane] = sm_;
                    wred[(lg * 2 + 1) * p.Cout + c0 + lane] = mx_;
                    __syncwarp();
                }
            }
            if (p.pool_part) {                       // combine the 4 epilogue warps (named barrier 1, 128 threads)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int et = (warp - 2) * 32 + lane;
                float* dst = p.pool_part + ((size_t)in * (p.tiles_h * p.tiles_w) + (tile - in * (p.tiles_h * p.tiles_w))) * 2 * p.Cout;
                for (int c = et; c < p.Cout; c += 128) {
                    dst[c] = (wred[0 * p.Cout + c] + wred[2 * p.Cout + c]) + (wred[4 * p.Cout + c] + wred[6 * p.Cout + c]);
                    dst[p.Cout + c] = fmaxf(fmaxf(wred[1 * p.Cout + c], wred[3 * p.Cout + c]), fmaxf(wred[5 * p.Cout + c], wred[7 * p.Cout + c]));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CG == 2 && crank != 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0)); else mbar_arrive(tempty_bar(acc)); }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (TS && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every tensor store has completed
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();      // pair: neither CTA leaves while the other can still reach its memory
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ tap-reuse variant
// conv_umma_halo_kernel: the classic kernel re-loads the A operand for each of the 9 taps, which makes N <= 128 convs
// (FPN-0, head conv3) L2->SMEM bandwidth bound.  Here the input lives in a chunk-major layout [N][C/E16][H][W][E16]
// (E16 = elements per 16 bytes) and ONE TMA box {10*E16, 18, 8 chunks} brings an (8+2)x(16+2) pixel halo tile of a
// 128-byte channel slab into shared memory as [chunk][row][col][16 B] -- exactly the canonical NO-SWIZZLE K-major UMMA
// layout for an 8-wide x 16-high output tile (core matrix = 8 pixels of a row x 16 B; SBO = halo row pitch = 160 B;
// LBO = chunk plane = 2880 B).  Every tap is then just a different descriptor start address (+ (ky*10+kx)*16 B), so
// the 23 KB slab is read from L2 once and used by 9 x 4 MMAs.  Weights stream per tap as before (128B-swizzled).
constexpr int HALO_W = 10, HALO_H = 18, HALO_CHUNK = HALO_W * HALO_H * 16;   // bytes per 16-byte-channel plane
constexpr int HALO_SLAB = 8 * HALO_CHUNK;                                    // 8 chunks = 128 bytes of channels

__device__ __forceinline__ uint64_t make_sdesc_noswz(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(HALO_CHUNK >> 4) << 16) | ((uint64_t)((HALO_W * 16) >> 4) << 32) | (1ull << 46);
}

template <bool BF16>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p,
                      int a_stages, int b_stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int B_BYTES = p.Cout * 128;
    uint8_t* smB = smem;                                           // [b_stages][Cout*128]   (1024-aligned: swizzled)
    uint8_t* smA = smem + (size_t)b_stages * B_BYTES;              // [a_stages][HALO_SLAB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smA + (size_t)a_stages * HALO_SLAB);
    // bars: a_full[a], a_empty[a], b_full[b], b_empty[b], tmem_full[2], tmem_empty[2]
    const int AS = a_stages, BS = b_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * AS + 2 * BS + 4);
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto afull = [&](int i) { return bar0 + 8u * i; };
    auto aempty = [&](int i) { return bar0 + 8u * (AS + i); };
    auto bfull = [&](int i) { return bar0 + 8u * (2 * AS + i); };
    auto bempty = [&](int i) { return bar0 + 8u * (2 * AS + BS + i); };
    auto tfull = [&](int i) { return bar0 + 8u * (2 * AS + 2 * BS + i); };
    auto tempty = [&](int i) { return bar0 + 8u * (2 * AS + 2 * BS + 2 + i); };
    int tmem_cols = 32; while (tmem_cols < 2 * p.Cout) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int i = 0; i < AS; i++) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); }
        for (int i = 0; i < BS; i++) { mbar_init(bfull(i), 1); mbar_init(bempty(i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) sbias[i] = p.bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_live = p.n_valid ? min(*p.n_valid, p.N) : p.N;
    const int total_tiles = n_live * p.tiles_h * p.tiles_w;
    constexpr int BK = BF16 ? 64 : 32;                 // channels per 128-byte slab
    constexpr int E16 = BF16 ? 8 : 4;                  // elements per 16-byte chunk
    const int slabs = p.Cin / BK;

    if (warp == 0) {
        if (lane == 0) {                                // ---------------- TMA producer (A slabs + per-tap weight tiles)
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int iw = tile % p.tiles_w, ih = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
                for (int sl = 0; sl < slabs; sl++) {
                    mbar_wait(aempty(as), aph ^ 1);
                    mbar_expect_tx(afull(as), (uint32_t)HALO_SLAB);
                    // coords: {element in the W*E16 row, h, chunk, n}; negative / past-the-end coordinates are zero-filled
                    tma_load_4d(smem_u32(smA + (size_t)as * HALO_SLAB), &tmA, afull(as), (iw * 8 - 1) * E16, ih * 16 - 1, sl * 8, n);
                    if (++as == AS) { as = 0; aph ^= 1; }
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(bempty(bs), bph ^ 1);
                        mbar_expect_tx(bfull(bs), (uint32_t)B_BYTES);
                        tma_load_2d(smem_u32(smB + (size_t)bs * B_BYTES), &tmB, bfull(bs), tap * p.Cin + sl * BK, 0);
                        if (++bs == BS) { bs = 0; bph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                // ---------------- MMA issuer
            const uint32_t fmt = BF16 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((128u >> 4) << 24);
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
                for (int sl = 0; sl < slabs; sl++) {
                    mbar_wait(afull(as), aph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smA + (size_t)as * HALO_SLAB);
                    for (int tap = 0; tap < 9; tap++) {
                        mbar_wait(bfull(bs), bph);
                        tc_fence_after();
                        const uint64_t adesc = make_sdesc_noswz(sa + (uint32_t)((tap / 3) * HALO_W + tap % 3) * 16u);
                        const uint64_t bdesc = make_sdesc(smem_u32(smB + (size_t)bs * B_BYTES));
#pragma unroll
                        for (int k4 = 0; k4 < 4; k4++)  // 32 bytes of K = two 16-byte chunk planes per MMA
                            tc_mma<BF16>(d_tmem, adesc + (uint64_t)((2 * k4 * HALO_CHUNK) >> 4), bdesc + 2 * k4, idesc, (sl | tap | k4) != 0 ? 1u : 0u);
                        tc_commit(bempty(bs));
                        if (++bs == BS) { bs = 0; bph ^= 1; }
                    }
                    tc_commit(aempty(as));
                    if (++as == AS) { as = 0; aph ^= 1; }
                }
                tc_commit(tfull(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {                                            // ---------------- epilogue warps 2..5
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const int rw = row & 7, rh = row >> 3;          // 8-wide x 16-high output tile
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int iw = tile % p.tiles_w, ih = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
            const int w = iw * 8 + rw, h = ih * 16 + rh;
            const bool valid = (w < p.W) && (h < p.H);
            const size_t pix = ((size_t)n * p.H + h) * p.W + w;
            mbar_wait_warp(tfull(acc), acc_phase, lane);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * p.Cout);
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                uint32_t r[32];
                tc_ld32(taddr + c0, r);
                tc_wait_ld();
                if (valid) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float x = __uint_as_float(r[j]) + sbias[c0 + j];
                        x = p.act == ACT_RELU ? fmaxf(x, 0.f) : x;
                        v[j] = p.round_out_tf32 ? tf32_rna(x) : x;
                    }
                    epi_store(p, v, n, h, w, pix, c0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
int conv_dense_umma(const kpd_plan* plan, const DenseConv& c, int prec_bf16, const void* in, void* out, int out_bf16,
                    int N, int H, int W, int act, int round_out_tf32, const int32_t* n_valid, cudaStream_t s, float* pool_part,
                    const float* res, int res_up, const FinalFuse* fin) {
    if (N <= 0) return 0;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    const int elt = prec_bf16 ? 2 : 4;
    const int row_bytes = (c.cin * elt) % 128 == 0 ? 128 : 64;      // 64-byte swizzle for the 16-channel fp32 lateral
    const int BK = row_bytes / elt;
    if (c.cin % BK != 0 || c.cout % 16 != 0 || c.cout > 256 || c.cout < 16) { set_error("conv_dense_umma: unsupported channels %d->%d", c.cin, c.cout); return KPD_ERR_UNSUPPORTED; }
    ConvParams p{};
    p.N = N; p.H = H; p.W = W; p.Cin = c.cin; p.Cout = c.cout; p.taps = c.taps;
    if (W % 16 == 0 && H % 8 == 0) { p.tw = 16; p.th = 8; p.tn = 1; } else { p.tw = 8; p.th = 8; p.tn = 2; }
    p.tiles_w = (W + p.tw - 1) / p.tw; p.tiles_h = (H + p.th - 1) / p.th; p.tiles_n = (N + p.tn - 1) / p.tn;
    p.act = act; p.round_out_tf32 = round_out_tf32; p.out_bf16 = out_bf16; p.bias = c.bias; p.out = out; p.n_valid = n_valid;
    p.pool_part = pool_part; p.row_bytes = row_bytes; p.res = res; p.res_up = res_up > 0 ? res_up : 1;
    if (fin) {
        if (c.cout != 64 || pool_part || res) { set_error("conv_dense_umma: the fused final 1x1 needs a plain 64-channel conv"); return KPD_ERR_UNSUPPORTED; }
        p.fin_w = fin->w; p.fin_b = fin->b; p.fin_out = fin->out; p.fin_slot = fin->slot;
    }
    int mt = 1;                                          // M-tiles per weight stage: as many as TMEM (512 columns, double-buffered) allows, <= 4
    while (mt < 4 && 2 * (mt * 2) * c.cout <= 512) mt *= 2;
    if (row_bytes == 64) mt = 1;
    p.mt = mt;
    if (pool_part && p.tn != 1) { set_error("conv_dense_umma: fused pooling needs W %% 16 == 0 and H %% 8 == 0"); return KPD_ERR_UNSUPPORTED; }
    // Epilogue form: TMA tensor stores for the plain NHWC convs (head convs 64->256, 256->256; the FPN levels 1-3); the fused
    // variants (pooling, residual, final 1x1) keep their register epilogues; the fused final 1x1 runs two epilogue warps per
    // TMEM lane group, each on half of the group's M-tiles (its 64x17 FMAs per pixel made a 4-warp epilogue the bottleneck).
    const bool ts = !fin && !pool_part && !res && row_bytes == 128 && c.cout % (prec_bf16 && out_bf16 ? 64 : 32) == 0 && (out_bf16 != 0) == (prec_bf16 != 0) && 32 % p.tw == 0;
    const bool ew8 = fin && mt == 4;
    const int ew = ew8 ? 8 : 4;
    // CTA pairs (cta_group::2) for the 256-wide convs with the TMA-store epilogue and a long K loop: each CTA stages half of the
    // weight tile.  Measured at C2: conv 256->256 (K = 2304) 0.639 -> 0.624 ms; conv 64->256 (K = 576) 0.207 -> 0.221 ms -- with nine
    // slabs per tile the pair's cross-CTA hand-offs at every tile boundary cost more than the halved weight fill saves, so short-K
    // convs stay single-CTA.
    const bool pair = ts && mt == 1 && c.cout == 256 && c.taps * c.cin >= 1024 && plan->sm_count >= 2;
    const int stage_bytes = mt * 128 * row_bytes + (pair ? c.cout / 2 : c.cout) * row_bytes;
    const int misc = 1024 /*align*/ + (ts ? ew * 2 * TS_WARP_BYTES : 0) + 8 * (2 * 8 + 4) + 16 + 4 * c.cout + 64 + (pool_part ? 4 * (4 * 32 * 33 + 8 * c.cout) : 0) + (fin ? 4 * 65 * FIN_KP + 16 : 0);
    int stages = (227 * 1024 - misc) / stage_bytes; if (stages > 8) stages = 8;
    if (stages < 2) { set_error("conv_dense_umma: not enough shared memory"); return KPD_ERR_UNSUPPORTED; }
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + misc;

    CUtensorMap tmA, tmB, tmO;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c.cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)c.cin * elt, (cuuint64_t)W * c.cin * elt, (cuuint64_t)H * W * c.cin * elt};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, prec_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         const_cast<void*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        const cuuint64_t K = (cuuint64_t)c.taps * c.cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)c.cout};
        cuuint64_t strides[1] = {K * elt};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(pair ? c.cout / 2 : c.cout)};
        cuuint32_t es[2] = {1, 1};
        const void* wptr = prec_bf16 ? static_cast<const void*>(c.w_bf16) : static_cast<const void*>(c.w_tf32);
        CUresult r = enc(&tmB, prec_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(wptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    tmO = tmB;                                           // placeholder when the epilogue stores from registers
    if (ts) {
        // output tensor [N][H][W][Cout] as seen by one epilogue warp: 128 bytes of channels x its 32 pixels (tw x 32/tw); pixels
        // outside the image (partial edge tiles) and instances past N are clipped by the TMA unit
        const int oelt = out_bf16 ? 2 : 4;
        cuuint64_t dims[4] = {(cuuint64_t)c.cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)c.cout * oelt, (cuuint64_t)W * c.cout * oelt, (cuuint64_t)H * W * c.cout * oelt};
        cuuint32_t box[4] = {(cuuint32_t)(128 / oelt), (cuuint32_t)p.tw, (cuuint32_t)(32 / p.tw), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmO, out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(out) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    int grid = std::min((total_tiles + mt - 1) / mt, plan->sm_count);
    if (pair) grid = std::min(2 * ((total_tiles + 1) / 2), plan->sm_count & ~1);
    static uint64_t attr[20] = {};
#define KPD_LAUNCH_UMMA(BF, MTV, RBV, EWV, TSV, SLOT)                                                                     \
    do {                                                                                                                  \
        KPD_CUDA(ensure_smem_attr(conv_umma_kernel<BF, MTV, RBV, EWV, TSV, 1>, 227 * 1024, attr[SLOT]));                  \
        conv_umma_kernel<BF, MTV, RBV, EWV, TSV, 1><<<grid, 64 + 32 * EWV, smem, s>>>(tmA, tmB, tmO, p);                  \
    } while (0)
#define KPD_LAUNCH_UMMA_PAIR(BF, SLOT)                                                                                    \
    do {                                                                                                                  \
        KPD_CUDA(ensure_smem_attr(conv_umma_kernel<BF, 1, 128, 4, true, 2>, 227 * 1024, attr[SLOT]));                     \
        cudaLaunchConfig_t cfg{};                                                                                         \
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * 4); cfg.dynamicSmemBytes = smem; cfg.stream = s;          \
        cudaLaunchAttribute at[1];                                                                                        \
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; \
        cfg.attrs = at; cfg.numAttrs = 1;                                                                                 \
        KPD_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<BF, 1, 128, 4, true, 2>, tmA, tmB, tmO, p));                   \
    } while (0)
    if (pair) {
        if (prec_bf16) KPD_LAUNCH_UMMA_PAIR(true, 15); else KPD_LAUNCH_UMMA_PAIR(false, 16);
    } else
    if (row_bytes == 64) {
        if (prec_bf16 || mt != 1) { set_error("conv_dense_umma: 64-byte K rows are built for tf32, mt=1"); return KPD_ERR_UNSUPPORTED; }
        KPD_LAUNCH_UMMA(false, 1, 64, 4, false, 0);
    } else if (ew8) {
        if (prec_bf16) KPD_LAUNCH_UMMA(true, 4, 128, 8, false, 1); else KPD_LAUNCH_UMMA(false, 4, 128, 8, false, 2);
    } else if (ts) {
        if (prec_bf16) { if (mt == 4) KPD_LAUNCH_UMMA(true, 4, 128, 4, true, 3); else if (mt == 2) KPD_LAUNCH_UMMA(true, 2, 128, 4, true, 4); else KPD_LAUNCH_UMMA(true, 1, 128, 4, true, 5); }
        else { if (mt == 4) KPD_LAUNCH_UMMA(false, 4, 128, 4, true, 6); else if (mt == 2) KPD_LAUNCH_UMMA(false, 2, 128, 4, true, 7); else KPD_LAUNCH_UMMA(false, 1, 128, 4, true, 8); }
    } else if (prec_bf16) {
        if (mt == 4) KPD_LAUNCH_UMMA(true, 4, 128, 4, false, 9); else if (mt == 2) KPD_LAUNCH_UMMA(true, 2, 128, 4, false, 10); else KPD_LAUNCH_UMMA(true, 1, 128, 4, false, 11);
    } else {
        if (mt == 4) KPD_LAUNCH_UMMA(false, 4, 128, 4, false, 12); else if (mt == 2) KPD_LAUNCH_UMMA(false, 2, 128, 4, false, 13); else KPD_LAUNCH_UMMA(false, 1, 128, 4, false, 14);
    }
#undef KPD_LAUNCH_UMMA
#undef KPD_LAUNCH_UMMA_PAIR
    KPD_LAUNCH_CHECK();
    return 0;
}

// Tap-reuse launcher.  `in` is chunk-major [N][Cin/E16][H][W][E16]; taps must be 9.
int conv_dense_umma_halo(const kpd_plan* plan, const DenseConv& c, int prec_bf16, const void* in, void* out, int out_bf16,
                         int out_cm, int N, int H, int W, int act, int round_out_tf32, const int32_t* n_valid, cudaStream_t s) {
    if (N <= 0) return 0;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    const int elt = prec_bf16 ? 2 : 4, E16 = 16 / elt, BK = 128 / elt;
    if (c.taps != 9 || c.cin % BK != 0 || c.cout % 16 != 0 || c.cout > 256 || c.cout < 16) { set_error("conv_dense_umma_halo: unsupported conv %d->%d taps %d", c.cin, c.cout, c.taps); return KPD_ERR_UNSUPPORTED; }
    ConvParams p{};
    p.N = N; p.H = H; p.W = W; p.Cin = c.cin; p.Cout = c.cout; p.taps = 9;
    p.tw = 8; p.th = 16; p.tn = 1;
    p.tiles_w = (W + 7) / 8; p.tiles_h = (H + 15) / 16; p.tiles_n = N;
    p.act = act; p.round_out_tf32 = round_out_tf32; p.out_bf16 = out_bf16; p.out_cm = out_cm; p.bias = c.bias; p.out = out; p.n_valid = n_valid;
    const int b_bytes = c.cout * 128;
    const int a_stages = 3;
    const int misc = 1024 + 8 * (2 * 3 + 2 * 8 + 4) + 16 + 4 * c.cout + 64;
    int b_stages = (227 * 1024 - misc - a_stages * HALO_SLAB) / b_bytes; if (b_stages > 8) b_stages = 8;
    if (b_stages < 2) { set_error("conv_dense_umma_halo: not enough shared memory"); return KPD_ERR_UNSUPPORTED; }
    const size_t smem = (size_t)b_stages * b_bytes + (size_t)a_stages * HALO_SLAB + misc;
    CUtensorMap tmA, tmB;
    {
        const cuuint64_t CH = (cuuint64_t)c.cin / E16;
        cuuint64_t dims[4] = {(cuuint64_t)W * E16, (cuuint64_t)H, CH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, CH * H * W * 16};
        cuuint32_t box[4] = {(cuuint32_t)(HALO_W * E16), (cuuint32_t)HALO_H, 8, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmA, prec_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         const_cast<void*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A halo) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        const cuuint64_t K = (cuuint64_t)9 * c.cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)c.cout};
        cuuint64_t strides[1] = {K * elt};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)c.cout};
        cuuint32_t es[2] = {1, 1};
        const void* wptr = prec_bf16 ? static_cast<const void*>(c.w_bf16) : static_cast<const void*>(c.w_tf32);
        CUresult r = enc(&tmB, prec_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(wptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    const int grid = std::min(total_tiles, plan->sm_count);
    if (prec_bf16) {
        KPD_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_umma_halo_kernel<true><<<grid, UMMA_THREADS, smem, s>>>(tmA, tmB, p, a_stages, b_stages);
    } else {
        KPD_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_umma_halo_kernel<false><<<grid, UMMA_THREADS, smem, s>>>(tmA, tmB, p, a_stages, b_stages);
    }
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
