// Batch-wide pointwise (1x1) convolution on tensor cores and batch-wide depthwise convolution, NHWC fp32.
//   reference ops: every 1x1 conv of the trunk -- tv:models/mobilenetv3.py:87-108 (expand / project), features.12,
//   dll/models/backbone.py:33-37 (FPN laterals + nearest top-down add) -- and tv:models/mobilenetv3.py:92-104 (depthwise).
//
// pw_mma_kernel: out[M][N] = act(A[M][K] * W[K][N] + bias) (+ residual), M = batch*H*W pixels.  CTA tile 128 pixels x
// up to 128 output channels, 8 warps x 16 rows; A is streamed through shared memory in K chunks of 32 with register
// prefetch (optionally scaled by the per-image SE excitation while it is staged); B fragments are pre-packed on the
// host in mma.sync m16n8k8 order (hi/lo TF32 split) and come from L1/L2.  3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,
// fp32 accumulate) keeps the trunk at fp32-level accuracy, which the channel top-k needs.
#include "kpd_common.cuh"
#include <stdlib.h>

namespace kpd {

__device__ __forceinline__ void mma_tf32_pw(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


constexpr int PW_BM = 128, PW_KC = 32, PW_LDA = PW_KC + 4, PW_NT = 16, PW_THREADS = 256;

constexpr int PW_STAGES = 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;                       // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}

// Persistent CTAs walk the (m-tile, n-chunk) items; the A operand streams through a 4-stage cp.async ring that runs
// up to 3 K-chunks (or whole tiles, when K <= 32) ahead of the MMAs, across item boundaries -- so the many small-K GEMMs
// of the trunk are bandwidth- instead of latency-bound.  The SE excitation (a per-image, per-input-channel factor) is
// applied to the A fragments as they are read from shared memory.
__global__ void __launch_bounds__(PW_THREADS) pw_mma_kernel(const PwMma p, int tiles_m, int tiles_n) {
    extern __shared__ __align__(16) float As_all[];      // [PW_STAGES][PW_BM * PW_LDA]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int n8 = p.N >> 3, K = p.K;
    const uint32_t M = (uint32_t)p.M;                    // 32-bit pixel indices (host-checked): no 64-bit div/mod on the item path
    const int sk4 = tid & 7, srow = tid >> 3;            // staging map: float4 (row = srow + 32*i, k4 = sk4), i = 0..3
    const int total = tiles_m * tiles_n;

    int l_item = blockIdx.x, l_kc = 0;                   // load cursor (runs ahead of the compute cursor)
    auto issue_load = [&](int stage) {
        if (l_item < total) {
            float* dst = As_all + (size_t)stage * (PW_BM * PW_LDA);
            const uint32_t m0 = (uint32_t)(l_item / tiles_n) * PW_BM;
            const int k = l_kc + sk4 * 4;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t m = m0 + srow + 32 * i;
                const bool ok = (m < M) && (k < K);
                cp_async16(dst + (srow + 32 * i) * PW_LDA + sk4 * 4, ok ? (const void*)(p.A + (size_t)m * K + k) : (const void*)p.A, ok);
            }
            l_kc += PW_KC;
            if (l_kc >= K) { l_kc = 0; l_item += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < PW_STAGES - 1; s++) issue_load(s);

    int stage = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int tm = item / tiles_n;
        const uint32_t m0 = (uint32_t)tm * PW_BM;
        const int nt0 = (item - tm * tiles_n) * PW_NT, ntc = min(PW_NT, n8 - nt0);
        const uint32_t r[2] = {m0 + warp * 16 + g, m0 + warp * 16 + g + 8};
        const float* sc0 = nullptr; const float* sc1 = nullptr;
        if (p.se_scale) {
            sc0 = p.se_scale + (size_t)(min(r[0], M - 1) / (uint32_t)p.hw) * K + t;
            sc1 = p.se_scale + (size_t)(min(r[1], M - 1) / (uint32_t)p.hw) * K + t;
        }
        float acc[PW_NT][4];
#pragma unroll
        for (int j = 0; j < PW_NT; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
        for (int kc = 0; kc < K; kc += PW_KC) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PW_STAGES - 2) : "memory");
            __syncthreads();                              // chunk `stage` has landed; everyone is done with the slot refilled next
            issue_load((stage + PW_STAGES - 1) % PW_STAGES);
            const float* As = As_all + (size_t)stage * (PW_BM * PW_LDA);
            const int ksteps = min(PW_KC, K - kc) >> 3;
            const float* a0p = As + (warp * 16 + g) * PW_LDA + t;
            const float* a1p = a0p + 8 * PW_LDA;
            for (int ks = 0; ks < ksteps; ks++) {
                float av[4] = {a0p[ks * 8], a1p[ks * 8], a0p[ks * 8 + 4], a1p[ks * 8 + 4]};
                if (sc0) {
                    const int k = kc + ks * 8;
                    av[0] *= __ldg(sc0 + k); av[1] *= __ldg(sc1 + k); av[2] *= __ldg(sc0 + k + 4); av[3] *= __ldg(sc1 + k + 4);
                }
                uint32_t ah[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float h = tf32_rna(av[i]);
                    ah[i] = __float_as_uint(h); al[i] = __float_as_uint(tf32_rna(av[i] - h));
                }
                const float4* wp = p.Wf + ((size_t)((kc >> 3) + ks) * n8 + nt0) * 32 + lane;
#pragma unroll
                for (int j = 0; j < PW_NT; j++) {
                    if (j < ntc) {
                        const float4 b = __ldg(wp + j * 32);
                        mma_tf32_pw(acc[j], al, __float_as_uint(b.x), __float_as_uint(b.y));
                        mma_tf32_pw(acc[j], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                        mma_tf32_pw(acc[j], ah, __float_as_uint(b.x), __float_as_uint(b.y));
                    }
                }
            }
            stage = (stage + 1) % PW_STAGES;
        }
        // epilogue: rows r0 = g, r1 = g + 8 of this warp's 16-row slab; columns n = (nt0+j)*8 + 2t, +1
        size_t rbase[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (r[h] >= M) continue;
            if (p.res_mode == 1) rbase[h] = (size_t)r[h] * p.N;
            else if (p.res_mode == 2) {
                const uint32_t row = r[h] / (uint32_t)p.W, x = r[h] - row * p.W, b = row / (uint32_t)p.H, y = row - b * p.H;
                rbase[h] = ((size_t)(b * (p.H / p.up) + y / p.up) * (p.W / p.up) + x / p.up) * p.N;
            }
        }
#pragma unroll
        for (int j = 0; j < PW_NT; j++) {
            if (j >= ntc) break;
            const int n = (nt0 + j) * 8 + 2 * t;
            const float b0 = p.bias ? __ldg(p.bias + n) : 0.f, b1 = p.bias ? __ldg(p.bias + n + 1) : 0.f;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (r[h] >= M) continue;
                float v0 = apply_act(acc[j][2 * h] + b0, p.act), v1 = apply_act(acc[j][2 * h + 1] + b1, p.act);
                if (p.res_mode) {
                    const float2 q = __ldg(reinterpret_cast<const float2*>(p.res + rbase[h] + n));
                    v0 += q.x; v1 += q.y;
                }
                if (p.round_tf32) { v0 = tf32_rna(v0); v1 = tf32_rna(v1); }
                *reinterpret_cast<float2*>(p.out + (size_t)r[h] * p.N + n) = make_float2(v0, v1);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

int pointwise_mma(const PwMma& p, int sm_count, cudaStream_t s) {
    // Dispatch between the three GEMM implementations.  Default (non-strict modes): tcgen05 3xTF32 (kpd_pwu.cu), except for the
    // shapes where its per-tile pipeline (TMA -> split -> MMA -> epilogue) measured slower than the CUDA-core kernel: K <= 16
    // (one half-empty slab per tile), very narrow outputs with a residual, and the K = 24 lateral with the upsampled residual.
    // Strict fp32 always takes the exact FFMA kernel.  KPD_PW_FFMA=1 / KPD_PW_UMMA=1 / KPD_PW_MMA=1 force one implementation.
    static const bool use_mma = getenv("KPD_PW_MMA") != nullptr;       // 3xTF32 mma.sync predecessor, kept for A/B timing
    static const bool force_umma = getenv("KPD_PW_UMMA") != nullptr;
    static const bool force_ffma = getenv("KPD_PW_FFMA") != nullptr;
    if (!use_mma && p.Wu && !p.exact && !force_ffma) {
        const bool prefer_ffma = (p.K <= 16) || (p.N <= 24 && p.res) || (p.res && p.res_mode == 2 && p.K <= 24);
        if (force_umma || !prefer_ffma) return pointwise_umma(p, sm_count, s);
    }
    if (!use_mma && p.Wraw) return pointwise_ffma(p, sm_count, s);
    if ((p.K & 7) || (p.N & 7)) { set_error("pointwise_mma: K and N must be multiples of 8 (got %d, %d)", p.K, p.N); return KPD_ERR_UNSUPPORTED; }
    if (p.M <= 0) return 0;
    if (p.M >= ((int64_t)1 << 31) - PW_BM) { set_error("pointwise_mma: M = %lld pixels exceeds the 32-bit pixel index", (long long)p.M); return KPD_ERR_UNSUPPORTED; }
    const int tiles_m = (int)((p.M + PW_BM - 1) / PW_BM), tiles_n = ((p.N >> 3) + PW_NT - 1) / PW_NT;
    const size_t smem = sizeof(float) * PW_STAGES * PW_BM * PW_LDA;           // 72 KB: 2-3 CTAs per SM
    static uint64_t attr_mask = 0;
    KPD_CUDA(ensure_smem_attr(pw_mma_kernel, (int)smem, attr_mask));
    const int grid = (int)std::min<int64_t>((int64_t)tiles_m * tiles_n, (int64_t)sm_count * 2);
    pw_mma_kernel<<<grid, PW_THREADS, smem, s>>>(p, tiles_m, tiles_n);
    KPD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Depthwise k x k conv (+ folded BN + act), NHWC fp32, batch-wide, shared-memory tiled; optional SE squeeze.
// CTA = (image, group of <= 32 channels, TH x TW output tile).  The zero-padded input tile (with halo) and the group's
// tap weights are staged in shared memory once (coalesced float4), so the k*k tap loop is branch-free LDS + FMA.  A thread
// keeps one 4-channel group (256 % channels-per-group/4 == 0), so the squeeze accumulates in registers; one shared atomic
// per thread and one global atomic per (CTA, channel) finish it.
template <int K>
__global__ void __launch_bounds__(256, 2) dw_tiled_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       float* __restrict__ pool, int H, int C, int stride, int act,
                                                       int TH, int TW, int tiles_x, int CG) {
    extern __shared__ __align__(16) float smem[];
    const int Ho = H / stride, pad = (K - 1) / 2;
    const int b = blockIdx.z, c0 = blockIdx.y * CG, cg = min(CG, C - c0), cg4 = cg >> 2;
    const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
    const int th = min(TH, Ho - ty0), tw = min(TW, Ho - tx0);
    const int in_h = (th - 1) * stride + K, in_w = (tw - 1) * stride + K;
    float* tile = smem;                                   // [in_h][in_w][cg]
    float* wsm = smem + (size_t)((TH - 1) * stride + K) * ((TW - 1) * stride + K) * CG;   // [K*K][cg]
    float* red = wsm + K * K * CG;                        // [cg]
    const int tid = threadIdx.x;
    const int iy0 = ty0 * stride - pad, ix0 = tx0 * stride - pad;
    const float* inb = in + (size_t)b * H * H * C + c0;
    for (int i = tid; i < in_h * in_w * cg4; i += 256) {
        const int c4 = i % cg4, pxl = i / cg4;
        const int yy = pxl / in_w, xx = pxl - yy * in_w;
        const int iy = iy0 + yy, ix = ix0 + xx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < H) v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)iy * H + ix) * C) + c4);
        reinterpret_cast<float4*>(tile + (size_t)pxl * cg)[c4] = v;
    }
    for (int i = tid; i < K * K * cg4; i += 256) {
        const int c4 = i % cg4, tp = i / cg4;
        reinterpret_cast<float4*>(wsm + tp * cg)[c4] = __ldg(reinterpret_cast<const float4*>(w + (size_t)tp * C + c0) + c4);
    }
    if (tid < cg) red[tid] = 0.f;
    __syncthreads();
    const int c4 = tid % cg4;                             // fixed for the thread: 256 % cg4 == 0
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + c4);
    float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f);
    float* outb = out + (size_t)b * Ho * Ho * C + c0;
    // Sliding window: an item is a run of XR consecutive outputs of one row for one 4-channel group.  Per tap row the
    // (XR-1)*stride+K input vectors and the K weights are loaded from shared memory once and reused from registers
    // (~3x fewer LDS than one load pair per tap; the plain form is shared-memory-bandwidth bound).
    constexpr int XR = 4;
    const int runs = (tw + XR - 1) / XR;
    for (int item = tid; item < th * runs * cg4; item += 256) {
        const int q = item / cg4;
        const int oy = q / runs, ox0 = (q - oy * runs) * XR;
        float4 acc[XR];
#pragma unroll
        for (int j = 0; j < XR; j++) acc[j] = bb;
        if (stride == 1) {
            constexpr int SEG = XR - 1 + K;
#pragma unroll
            for (int ky = 0; ky < K; ky++) {
                const float* rp = tile + ((size_t)(oy + ky) * in_w + ox0) * cg + c4 * 4;
                float4 seg[SEG], wk[K];
#pragma unroll
                for (int i = 0; i < SEG; i++) seg[i] = *reinterpret_cast<const float4*>(rp + (size_t)i * cg);   // a partial last run over-reads into the weight area: unused lanes
#pragma unroll
                for (int kx = 0; kx < K; kx++) wk[kx] = *reinterpret_cast<const float4*>(wsm + (ky * K + kx) * cg + c4 * 4);
#pragma unroll
                for (int j = 0; j < XR; j++)
#pragma unroll
                    for (int kx = 0; kx < K; kx++) {
                        acc[j].x = fmaf(seg[j + kx].x, wk[kx].x, acc[j].x); acc[j].y = fmaf(seg[j + kx].y, wk[kx].y, acc[j].y);
                        acc[j].z = fmaf(seg[j + kx].z, wk[kx].z, acc[j].z); acc[j].w = fmaf(seg[j + kx].w, wk[kx].w, acc[j].w);
                    }
            }
        } else {
            constexpr int SEG = (XR - 1) * 2 + K;
#pragma unroll
            for (int ky = 0; ky < K; ky++) {
                const float* rp = tile + ((size_t)(oy * 2 + ky) * in_w + ox0 * 2) * cg + c4 * 4;
                float4 seg[SEG], wk[K];
#pragma unroll
                for (int i = 0; i < SEG; i++) seg[i] = *reinterpret_cast<const float4*>(rp + (size_t)i * cg);   // a partial last run over-reads into the weight area: unused lanes
#pragma unroll
                for (int kx = 0; kx < K; kx++) wk[kx] = *reinterpret_cast<const float4*>(wsm + (ky * K + kx) * cg + c4 * 4);
#pragma unroll
                for (int j = 0; j < XR; j++)
#pragma unroll
                    for (int kx = 0; kx < K; kx++) {
                        acc[j].x = fmaf(seg[2 * j + kx].x, wk[kx].x, acc[j].x); acc[j].y = fmaf(seg[2 * j + kx].y, wk[kx].y, acc[j].y);
                        acc[j].z = fmaf(seg[2 * j + kx].z, wk[kx].z, acc[j].z); acc[j].w = fmaf(seg[2 * j + kx].w, wk[kx].w, acc[j].w);
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < XR; j++) {
            if (ox0 + j >= tw) break;
            float4 r;
            r.x = apply_act(acc[j].x, act); r.y = apply_act(acc[j].y, act); r.z = apply_act(acc[j].z, act); r.w = apply_act(acc[j].w, act);
            reinterpret_cast<float4*>(outb + ((size_t)(ty0 + oy) * Ho + tx0 + ox0 + j) * C)[c4] = r;
            ssum.x += r.x; ssum.y += r.y; ssum.z += r.z; ssum.w += r.w;
        }
    }
    if (pool) {
        atomicAdd(&red[c4 * 4 + 0], ssum.x); atomicAdd(&red[c4 * 4 + 1], ssum.y);
        atomicAdd(&red[c4 * 4 + 2], ssum.z); atomicAdd(&red[c4 * 4 + 3], ssum.w);
        __syncthreads();
        if (tid < cg) atomicAdd(pool + (size_t)b * C + c0 + tid, red[tid]);
    }
}

int depthwise_tma(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s);

int depthwise_batched(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s) {
    static const bool tiled = getenv("KPD_DW_TILED") != nullptr;      // predecessor kept for A/B timing
    if (!tiled) return depthwise_tma(m, in, out, pool, pool_capacity, pool_parts, B, H, sm_count, s);
    if (pool_parts) *pool_parts = 1;                                  // the tiled kernel accumulates [B][C] with atomics
    if (pool) KPD_CUDA(cudaMemsetAsync(pool, 0, sizeof(float) * B * m.exp, s));
    const int Ho = H / m.stride, C = m.exp;
    if (C & 3) { set_error("depthwise_batched: unsupported channel count %d", C); return KPD_ERR_UNSUPPORTED; }
    // channel group: <= 32 channels with (group/4) dividing 256 for every group incl. the last partial one
    int CG = 32;
    if (C < 32) CG = C;
    else if ((C % 32) != 0 && (256 % ((C % 32) / 4)) != 0) CG = 16;
    if (256 % (CG / 4) != 0 || (C % CG != 0 && 256 % ((C % CG) / 4) != 0)) { set_error("depthwise_batched: channel grouping failed for C=%d", C); return KPD_ERR_UNSUPPORTED; }
    // output tile: full width up to 28, height so that the input tile stays <= ~40 KB
    const int TW = std::min(Ho, 28);
    int TH = std::min(Ho, 28);
    while (TH > 1 && (size_t)((TH - 1) * m.stride + m.k) * ((TW - 1) * m.stride + m.k) * CG * 4 > 40 * 1024) TH--;
    int tiles_y = (Ho + TH - 1) / TH;
    TH = (Ho + tiles_y - 1) / tiles_y;                    // balance the row tiles
    const int tiles_x = (Ho + TW - 1) / TW;
    const size_t smem = sizeof(float) * ((size_t)((TH - 1) * m.stride + m.k) * ((TW - 1) * m.stride + m.k) * CG + (size_t)m.k * m.k * CG + CG);
    dim3 grid(tiles_x * tiles_y, (C + CG - 1) / CG, B);
    (void)sm_count;
    static uint64_t attr3 = 0, attr5 = 0;
    KPD_CUDA(ensure_smem_attr(dw_tiled_kernel<3>, 64 * 1024, attr3));
    KPD_CUDA(ensure_smem_attr(dw_tiled_kernel<5>, 64 * 1024, attr5));
    if (m.k == 3) dw_tiled_kernel<3><<<grid, 256, smem, s>>>(in, m.w_dw, m.b_dw, out, pool, H, C, m.stride, m.act, TH, TW, tiles_x, CG);
    else          dw_tiled_kernel<5><<<grid, 256, smem, s>>>(in, m.w_dw, m.b_dw, out, pool, H, C, m.stride, m.act, TH, TW, tiles_x, CG);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
