// Fused inverted-residual (MBConv) block: ONE kernel per block, one thread-block cluster per image.
//   reference: tv:models/mobilenetv3.py:53-116 (InvertedResidual: expand 1x1 + BN + act -> depthwise kxk + BN + act ->
//   SqueezeExcitation (tv:ops/misc.py:225-261) -> project 1x1 + BN -> (+ residual)), eval-mode BN folded.
//
// Mapping: the C CTAs of a cluster split the depthwise output rows of one image into bands.  Each CTA
//   1. streams its input rows once (coalesced float4) in chunks, expands each chunk with a register-tiled SIMT GEMM
//      into a row ring buffer in shared memory (the expanded tensor never touches HBM), runs the depthwise conv out of
//      the ring into a band-resident buffer and keeps per-channel SE sums in registers (a thread owns fixed channels),
//   2. completes the SE squeeze across the cluster through distributed shared memory (cluster.sync), evaluates the
//      tiny excitation MLP redundantly per CTA,
//   3. scales the band, projects it (1x1 GEMM), adds the residual and writes the output (coalesced float4).
// HBM traffic per block = input (+ halo rows re-read by the neighbouring CTA) + output: the expanded / depthwise
// tensors (3-6x larger) stay on chip.  All arithmetic is fp32 FMA so the trunk keeps the reference's precision.
#include "kpd_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace kpd {

// n / d for 0 <= n < 2^31 without the ~25-instruction integer division sequence (host precomputes mul/sh)
struct FastDiv { uint32_t d, mul, sh; };
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f; f.d = d; uint32_t l = 0; while ((1u << l) < d) l++;
    f.mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - d)) / d + 1); f.sh = l;
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return (int)((__umulhi((uint32_t)n, f.mul) + (uint32_t)n) >> f.sh); }

struct MBArgs {
    const float* in; float* out;
    const float *w_exp, *b_exp, *w_dw, *b_dw, *w_fc1, *b_fc1, *w_fc2, *b_fc2, *w_proj, *b_proj;
    const float4 *wf_exp, *wf_proj;   // 3xTF32 mma.sync B fragments (null -> fp32 FMA GEMM)
    FastDiv fd_h, fd_ho, fd_e4, fd_c4, fd_r;
    int skip;                          // debug bitmask: 1 expand GEMM, 2 depthwise, 4 SE MLP, 8 project GEMM (timing attribution)
    int lde, ldx;                      // padded shared-memory row strides (floats) of the exp-channel / cin-channel tiles
    int cin, exp, cout, k, stride, se, sq, act, res;
    int H, Ho, C, rows_per_cta;
    int R;                 // ring capacity in rows (>= k)
    int chunk_rows;        // max input rows expanded per GEMM pass (x staging capacity)
    int p_exp, p_proj;     // pixels per thread item of the two GEMMs (1, 2, 4 or 8)
    int off_ring, off_dwo, off_x, off_pool, off_scale, off_hid, smem_floats;   // shared-memory carve (floats)
};

constexpr int MB_THREADS = 512;

// out[m][n4*4..] = bias + sum_k A[m][k] * W[k][n]  for m in [0,M), n in [0,N): register tile P pixels x 4 channels,
// K consumed 8 at a time (8 coalesced LDG.128 of W, 2 broadcast LDS.128 of A per pixel, 32 FMA per pixel).
template <int P, typename Epi>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ A, int lda, int M, int K,
                                          const float* __restrict__ W, int N, const float* __restrict__ bias, Epi epi) {
    const int N4 = N >> 2;
    const int groups = (M + P - 1) / P;
    for (int item = threadIdx.x; item < groups * N4; item += MB_THREADS) {
        const int n4 = item % N4, m0 = (item / N4) * P;
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + n4);
        float4 acc[P];
        const float* ap[P];
#pragma unroll
        for (int j = 0; j < P; j++) { acc[j] = b4; ap[j] = A + (size_t)min(m0 + j, M - 1) * lda; }
        for (int k0 = 0; k0 < K; k0 += 8) {
            float4 w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k0 + i) * N) + n4);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const float4 x0 = *reinterpret_cast<const float4*>(ap[j] + k0);
                const float4 x1 = *reinterpret_cast<const float4*>(ap[j] + k0 + 4);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    acc[j].x = fmaf(xs[i], w[i].x, acc[j].x); acc[j].y = fmaf(xs[i], w[i].y, acc[j].y);
                    acc[j].z = fmaf(xs[i], w[i].z, acc[j].z); acc[j].w = fmaf(xs[i], w[i].w, acc[j].w);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; j++)
            if (m0 + j < M) epi(m0 + j, n4, acc[j]);
    }
}

template <typename Epi>
__device__ __forceinline__ void gemm_dispatch(int P, const float* A, int lda, int M, int K, const float* W, int N,
                                              const float* bias, Epi epi) {
    if (P >= 4) gemm_tile<4>(A, lda, M, K, W, N, bias, epi);
    else if (P == 2) gemm_tile<2>(A, lda, M, K, W, N, bias, epi);
    else gemm_tile<1>(A, lda, M, K, W, N, bias, epi);
}

// ---- tensor-core variant of the 1x1 GEMMs: mma.sync m16n8k8 TF32 with the 3xTF32 split (a = a_hi + a_lo,
// b = b_hi + b_lo; a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate) => fp32-level accuracy (~2^-21 relative).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// epi(m, n, v0, v1) receives out[m][n], out[m][n+1] (bias added).  A: shared memory, row stride lda with
// lda % 32 in {4,12,20,28} (conflict-free fragment loads).  Wf: packed B fragments (see MBBlock::wf_exp).
template <typename Epi>
__device__ __forceinline__ void gemm_mma(const float* __restrict__ A, int lda, int M, int K, const float4* __restrict__ Wf,
                                         int N, const float* __restrict__ bias, Epi epi) {
    constexpr int NT = 2;                                 // n8 tiles per warp item
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int n8 = N >> 3, mtiles = (M + 15) >> 4, ngroups = (n8 + NT - 1) / NT;
    for (int item = warp; item < mtiles * ngroups; item += MB_THREADS / 32) {
        const int mt = item % mtiles, ng = item / mtiles;
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const float* a0p = A + (size_t)min(r0, M - 1) * lda + t;
        const float* a1p = A + (size_t)min(r1, M - 1) * lda + t;
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
        // B fragments are prefetched one k-step ahead (they come from L2: the big shared-memory carve-out leaves little L1)
        const float4* wp = Wf + ((size_t)ng * NT) * 32 + lane;
        const size_t wstep = (size_t)n8 * 32;
        float4 bc[NT], bn[NT];
#pragma unroll
        for (int j = 0; j < NT; j++) bc[j] = (ng * NT + j < n8) ? __ldg(wp + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < K; k0 += 8) {
            if (k0 + 8 < K) {
#pragma unroll
                for (int j = 0; j < NT; j++) bn[j] = (ng * NT + j < n8) ? __ldg(wp + wstep + j * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            wp += wstep;
            const float av[4] = {a0p[k0], a1p[k0], a0p[k0 + 4], a1p[k0 + 4]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float h = tf32_rna(av[i]);
                ah[i] = __float_as_uint(h); al[i] = __float_as_uint(tf32_rna(av[i] - h));
            }
#pragma unroll
            for (int j = 0; j < NT; j++) {
                mma_tf32(acc[j], al, __float_as_uint(bc[j].x), __float_as_uint(bc[j].y));
                mma_tf32(acc[j], ah, __float_as_uint(bc[j].z), __float_as_uint(bc[j].w));
                mma_tf32(acc[j], ah, __float_as_uint(bc[j].x), __float_as_uint(bc[j].y));
            }
#pragma unroll
            for (int j = 0; j < NT; j++) bc[j] = bn[j];
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
            if (ng * NT + j >= n8) break;
            const int n = (ng * NT + j) * 8 + 2 * t;
            const float b0 = __ldg(bias + n), b1 = __ldg(bias + n + 1);
            if (r0 < M) epi(r0, n, acc[j][0] + b0, acc[j][1] + b1);
            if (r1 < M) epi(r1, n, acc[j][2] + b0, acc[j][3] + b1);
        }
    }
}

template <bool MMA>
__global__ void __launch_bounds__(MB_THREADS, 1) mbconv_kernel(const MBArgs a) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / a.C;
    float* ring = sm + a.off_ring;        // [R][H][lde]   expanded input rows (zero rows outside the image)
    float* dwo = sm + a.off_dwo;          // [rows_per_cta*Ho][lde]
    float* xs = sm + a.off_x;             // [chunk_rows*H][ldx]
    float* pool = sm + a.off_pool;        // [exp] this CTA's partial channel sums
    float* scale = sm + a.off_scale;      // [exp]
    float* hid = sm + a.off_hid;          // [sq]

    const int H = a.H, Ho = a.Ho, E = a.exp, E4 = a.exp >> 2, k = a.k, pad = (a.k - 1) / 2, R = a.R, st = a.stride;
    const int LDE = a.lde, LDX = a.ldx;
    const int oy0 = rank * a.rows_per_cta, oy1 = min(Ho, oy0 + a.rows_per_cta);
    const float* inb = a.in + (size_t)b * H * H * a.cin;

    for (int i = tid; i < E; i += MB_THREADS) pool[i] = 0.f;
    const int dw_lanes = max(1, MB_THREADS / E4);        // pixel lanes of the depthwise stage
    const int dw_e4 = tid % E4, dw_px = tid / E4;
    const float4 dw_bias = __ldg(reinterpret_cast<const float4*>(a.b_dw) + dw_e4);
    float4 se_sum = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const int last_row = (oy1 - 1) * st - pad + k - 1;   // last input row this band reads
    int next_row = oy0 * st - pad;                       // next input row to expand into the ring
    int oy = oy0;
    while (oy < oy1) {
        // ---- 1a. expand as many new input rows as the ring / staging buffer can take
        const int row_lo = oy * st - pad;                // oldest row still needed
        int n_new = min(last_row - next_row + 1, R - (next_row - row_lo));
        if (a.w_exp) n_new = min(n_new, a.chunk_rows);
        if (n_new > 0) {
            const int r0 = next_row, r1 = next_row + n_new;
            const int v0 = max(r0, 0), v1 = min(r1, H);  // rows inside the image
            for (int r = r0; r < r1; r++) {              // rows outside the image are zero (conv padding)
                if (r >= 0 && r < H) continue;
                float* z = ring + (size_t)(((r % R) + R) % R) * H * LDE;
                for (int i = tid; i < H * E4; i += MB_THREADS) {
                    const int px = fdiv(i, a.fd_e4);
                    *reinterpret_cast<float4*>(z + (size_t)px * LDE + (i - px * E4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (v1 > v0) {
                const float4* xg = reinterpret_cast<const float4*>(inb + (size_t)v0 * H * a.cin);
                if (a.w_exp == nullptr) {                // exp == cin: the ring holds the raw input rows
                    for (int r = v0; r < v1; r++) {
                        float* d = ring + (size_t)(r % R) * H * LDE;
                        const float4* src = xg + (size_t)(r - v0) * H * E4;
                        for (int i = tid; i < H * E4; i += MB_THREADS) {
                            const int px = fdiv(i, a.fd_e4);
                            *reinterpret_cast<float4*>(d + (size_t)px * LDE + (i - px * E4) * 4) = __ldg(src + i);
                        }
                    }
                } else {
                    const int C4 = a.cin >> 2;
                    const int n4 = (v1 - v0) * H * C4;
                    for (int i = tid; i < n4; i += MB_THREADS) {
                        const int px = fdiv(i, a.fd_c4);
                        *reinterpret_cast<float4*>(xs + (size_t)px * LDX + (i - px * C4) * 4) = __ldg(xg + i);
                    }
                    __syncthreads();
                    const int act = a.act;
                    const int M = (a.skip & 1) ? 0 : (v1 - v0) * H;
                    const int ring_px0 = (v0 % R) * H, ring_npx = R * H;     // ring pixel of GEMM row m = (ring_px0 + m) mod R*H
                    if constexpr (MMA) {
                        gemm_mma(xs, LDX, M, a.cin, a.wf_exp, E, a.b_exp, [&](int m, int n, float x0, float x1) {
                            int rp = ring_px0 + m; if (rp >= ring_npx) rp -= ring_npx;
                            *reinterpret_cast<float2*>(ring + (size_t)rp * LDE + n) =
                                make_float2(apply_act(x0, act), apply_act(x1, act));
                        });
                    } else {
                        gemm_dispatch(a.p_exp, xs, LDX, M, a.cin, a.w_exp, E, a.b_exp, [&](int m, int n4i, float4 v) {
                            int rp = ring_px0 + m; if (rp >= ring_npx) rp -= ring_npx;
                            float4 o;
                            o.x = apply_act(v.x, act); o.y = apply_act(v.y, act);
                            o.z = apply_act(v.z, act); o.w = apply_act(v.w, act);
                            *reinterpret_cast<float4*>(ring + (size_t)rp * LDE + n4i * 4) = o;
                        });
                    }
                }
            }
            next_row = r1;
            __syncthreads();
        }
        // ---- 1b. depthwise conv (+ folded BN + act) of every output row whose inputs are now in the ring.
        //          A thread owns one 4-channel group for the whole band, so the SE squeeze accumulates in registers.
        int oy_end = oy;
        while (oy_end < oy1 && oy_end * st - pad + k - 1 < next_row) oy_end++;
        if (tid < dw_lanes * E4 && !(a.skip & 2)) {
            const int npx = (oy_end - oy) * Ho;
            constexpr int DG = 8;                                     // pixels per register group: a tap's weights are loaded once per group
            for (int q0 = dw_px; q0 < npx; q0 += dw_lanes * DG) {
                float4 acc[DG];
                int slot[DG], px[DG];                          // ring slot of the first tap row, first tap column
#pragma unroll
                for (int u = 0; u < DG; u++) {
                    const bool live = q0 + u * dw_lanes < npx;
                    const int q = live ? q0 + u * dw_lanes : 0;
                    const int qy = fdiv(q, a.fd_ho);
                    const int iy0 = (oy + qy) * st - pad + R;    // >= 0 (R >= k > pad)
                    slot[u] = iy0 - fdiv(iy0, a.fd_r) * R;
                    px[u] = live ? (q - qy * Ho) * st - pad : -(1 << 20);     // dead lanes fail the column test below
                    acc[u] = dw_bias;
                }
                for (int ky = 0; ky < k; ky++) {
                    for (int kx = 0; kx < k; kx++) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(a.w_dw + (size_t)(ky * k + kx) * E) + dw_e4);
#pragma unroll
                        for (int u = 0; u < DG; u++) {
                            const int ix = px[u] + kx;               // rows outside the image are zero rows of the ring
                            if (ix < 0 || ix >= H) continue;
                            int sl = slot[u] + ky; if (sl >= R) sl -= R;
                            const float4 v = *reinterpret_cast<const float4*>(ring + ((size_t)sl * H + ix) * LDE + dw_e4 * 4);
                            acc[u].x = fmaf(v.x, w.x, acc[u].x); acc[u].y = fmaf(v.y, w.y, acc[u].y);
                            acc[u].z = fmaf(v.z, w.z, acc[u].z); acc[u].w = fmaf(v.w, w.w, acc[u].w);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < DG; u++) {
                    const int q = q0 + u * dw_lanes;
                    if (q >= npx) break;
                    float4 r;
                    r.x = apply_act(acc[u].x, a.act); r.y = apply_act(acc[u].y, a.act);
                    r.z = apply_act(acc[u].z, a.act); r.w = apply_act(acc[u].w, a.act);
                    *reinterpret_cast<float4*>(dwo + ((size_t)(oy - oy0) * Ho + q) * LDE + dw_e4 * 4) = r;
                    se_sum.x += r.x; se_sum.y += r.y; se_sum.z += r.z; se_sum.w += r.w;
                }
            }
        }
        oy = oy_end;
        __syncthreads();
    }
    if (a.se && tid < dw_lanes * E4) {                     // dw_lanes contributors per channel, once per band
        atomicAdd(pool + dw_e4 * 4 + 0, se_sum.x); atomicAdd(pool + dw_e4 * 4 + 1, se_sum.y);
        atomicAdd(pool + dw_e4 * 4 + 2, se_sum.z); atomicAdd(pool + dw_e4 * 4 + 3, se_sum.w);
    }

    // ---- 2. squeeze-excitation across the cluster (DSMEM), excitation MLP per CTA
    if (a.se) {
        cluster.sync();                                   // every CTA's `pool` is final and visible cluster-wide
        for (int c = tid; c < E; c += MB_THREADS) {
            float s = 0.f;
            for (int r = 0; r < a.C; r++) s += cluster.map_shared_rank(pool, r)[c];
            scale[c] = s * (1.f / (float)(Ho * Ho));      // mean, reused below as the MLP input
        }
        __syncthreads();
        for (int j = tid; j < ((a.skip & 4) ? 0 : a.sq); j += MB_THREADS) {
            float acc0 = __ldg(a.b_fc1 + j), acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 4
            for (int c = 0; c < E; c += 4) {
                acc0 = fmaf(scale[c + 0], __ldg(a.w_fc1 + (size_t)(c + 0) * a.sq + j), acc0);
                acc1 = fmaf(scale[c + 1], __ldg(a.w_fc1 + (size_t)(c + 1) * a.sq + j), acc1);
                acc2 = fmaf(scale[c + 2], __ldg(a.w_fc1 + (size_t)(c + 2) * a.sq + j), acc2);
                acc3 = fmaf(scale[c + 3], __ldg(a.w_fc1 + (size_t)(c + 3) * a.sq + j), acc3);
            }
            hid[j] = fmaxf((acc0 + acc1) + (acc2 + acc3), 0.f);
        }
        __syncthreads();
        for (int c = tid; c < ((a.skip & 4) ? 0 : E); c += MB_THREADS) {
            float acc0 = __ldg(a.b_fc2 + c), acc1 = 0.f;
#pragma unroll 8
            for (int j = 0; j < a.sq; j += 2) {
                acc0 = fmaf(hid[j], __ldg(a.w_fc2 + (size_t)j * E + c), acc0);
                acc1 = fmaf(hid[j + 1], __ldg(a.w_fc2 + (size_t)(j + 1) * E + c), acc1);
            }
            scale[c] = apply_act(acc0 + acc1, ACT_HSIGMOID);
        }
        __syncthreads();
        const int n4 = (oy1 - oy0) * Ho * E4;
        for (int i = tid; i < n4; i += MB_THREADS) {      // scale the band in place: s * x  (misc.py:260)
            const int px = fdiv(i, a.fd_e4), e4 = i - px * E4;
            float4* pv = reinterpret_cast<float4*>(dwo + (size_t)px * LDE + e4 * 4);
            float4 v = *pv;
            const float4 s4 = reinterpret_cast<const float4*>(scale)[e4];
            v.x *= s4.x; v.y *= s4.y; v.z *= s4.z; v.w *= s4.w;
            *pv = v;
        }
        __syncthreads();
    }

    // ---- 3. project 1x1 (+ folded BN) (+ residual) -> global
    const int npx = (a.skip & 8) ? 0 : (oy1 - oy0) * Ho;
    if (npx > 0) {
        float* outb = a.out + ((size_t)b * Ho + oy0) * Ho * a.cout;
        const float* resb = a.res ? a.in + ((size_t)b * H + oy0) * H * a.cin : nullptr;   // res => stride 1, cin == cout
        const int cout = a.cout, cin = a.cin;
        if constexpr (MMA) {
            gemm_mma(dwo, LDE, npx, E, a.wf_proj, cout, a.b_proj, [&](int m, int n, float x0, float x1) {
                if (resb) {
                    const float2 q = __ldg(reinterpret_cast<const float2*>(resb + (size_t)m * cin + n));
                    x0 += q.x; x1 += q.y;
                }
                *reinterpret_cast<float2*>(outb + (size_t)m * cout + n) = make_float2(x0, x1);
            });
        } else {
            gemm_dispatch(a.p_proj, dwo, LDE, npx, E, a.w_proj, cout, a.b_proj, [&](int m, int n4i, float4 v) {
                if (resb) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(resb + (size_t)m * cin) + n4i);
                    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                }
                reinterpret_cast<float4*>(outb + (size_t)m * cout)[n4i] = v;
            });
        }
    }
    if (a.se) cluster.sync();                             // keep this CTA's shared memory alive until all peers read `pool`
}

static int pick_p(int M, int N4) {                        // largest pixel tile that still gives >= 2 items per thread
    for (int p : {4, 2}) if ((int64_t)((M + p - 1) / p) * N4 >= 2 * MB_THREADS) return p;
    return 1;
}

// Host launcher: picks the cluster size so that the band + ring fit in shared memory and the grid fills the GPU.
int run_mbconv(const kpd_plan* p, const MBBlock& m, const float* in, float* out, int B, int H, int strict_fp32, cudaStream_t s) {
    MBArgs a{};
    a.in = in; a.out = out;
    a.w_exp = m.w_exp; a.b_exp = m.b_exp; a.w_dw = m.w_dw; a.b_dw = m.b_dw; a.w_fc1 = m.w_fc1; a.b_fc1 = m.b_fc1;
    a.w_fc2 = m.w_fc2; a.b_fc2 = m.b_fc2; a.w_proj = m.w_proj; a.b_proj = m.b_proj;
    static const bool use_fma = getenv("KPD_MBCONV_FMA") != nullptr;     // debugging aid: plain fp32 FMA GEMMs
    a.wf_exp = (use_fma || strict_fp32) ? nullptr : m.wf_exp; a.wf_proj = (use_fma || strict_fp32) ? nullptr : m.wf_proj;
    a.cin = m.cin; a.exp = m.exp; a.cout = m.cout; a.k = m.k; a.stride = m.stride; a.se = m.se; a.sq = m.sq; a.act = m.act; a.res = m.res;
    a.H = H; a.Ho = H / m.stride;
    static const int skip = getenv("KPD_MB_SKIP") ? atoi(getenv("KPD_MB_SKIP")) : 0;
    a.skip = skip;
    if ((m.cin & 7) || (m.exp & 7) || (m.cout & 3) || (m.se && ((m.sq & 1) || (m.exp & 3)))) { set_error("run_mbconv: unsupported channel counts"); return KPD_ERR_UNSUPPORTED; }
    // tensor-core GEMMs make halo re-expansion cheap -> small bands, two CTAs per SM (phases of one CTA overlap the
    // other's latencies); the FMA variant keeps big bands (one CTA per SM) because its expand GEMM dominates.
    const bool mma = a.wf_proj != nullptr;
    static const int limit_kb = getenv("KPD_MB_LIMIT_KB") ? atoi(getenv("KPD_MB_LIMIT_KB")) : 200;
    const size_t limit = (size_t)limit_kb * 1024; (void)mma;
    auto carve = [&](int c) {
        const int rpc = (a.Ho + c - 1) / c;
        const int rows_needed = (rpc - 1) * m.stride + m.k;
        int off = 0;
        auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
        a.lde = m.exp + 4; a.ldx = m.cin + 4;           // +4: bank-conflict-free mma fragment loads, keeps 16-byte alignment
        a.off_dwo = take(rpc * a.Ho * a.lde);
        a.off_pool = take(m.exp); a.off_scale = take(m.exp); a.off_hid = take(m.sq > 0 ? m.sq : 4);
        // x staging: up to 16 KB, at least one row
        a.chunk_rows = m.w_exp ? std::max(1, std::min(rows_needed, (16 * 1024) / (H * m.cin * 4))) : 0;
        a.off_x = take(a.chunk_rows * H * a.ldx);
        // ring: everything the band needs if it fits in the remaining budget, else as many rows as fit (>= k)
        const int64_t left = (int64_t)(limit / 4) - off;
        int R = (int)std::min<int64_t>(rows_needed, left / ((int64_t)H * a.lde));
        a.R = R; a.off_ring = take(std::max(R, 0) * H * a.lde);
        a.smem_floats = off; a.C = c; a.rows_per_cta = rpc;
        return R >= m.k;
    };
    // Cluster size: the band must fit; prefer the configuration that needs the fewest expand/depthwise passes (a ring
    // that holds every row the band reads => one fat pass instead of many latency-bound thin ones), then enough CTAs
    // to fill the machine, then the smaller cluster (less halo re-expansion).
    int C = 0; int best_passes = 1 << 30;
    for (int c : {1, 2, 4, 8}) {
        if (c > a.Ho) break;
        if (!carve(c)) continue;
        const int rows_needed = (a.rows_per_cta - 1) * m.stride + m.k;
        int passes = a.R >= rows_needed ? 1 : 1 + (rows_needed - a.R + std::max(1, a.R - (m.k - m.stride)) - 1) / std::max(1, a.R - (m.k - m.stride));
        if ((int64_t)B * c < p->sm_count) passes += 2;                    // too few CTAs to fill the GPU
        if (passes < best_passes) { best_passes = passes; C = c; }
    }
    if (C == 0) { set_error("run_mbconv: block does not fit in shared memory at H=%d (exp=%d)", H, m.exp); return KPD_ERR_UNSUPPORTED; }
    carve(C);
    a.fd_h = make_fastdiv(H); a.fd_ho = make_fastdiv(a.Ho); a.fd_e4 = make_fastdiv(m.exp / 4);
    a.fd_c4 = make_fastdiv(m.cin / 4); a.fd_r = make_fastdiv(a.R);
    a.p_exp = m.w_exp ? pick_p(std::min(a.chunk_rows, (a.rows_per_cta - 1) * m.stride + m.k) * H, m.exp / 4) : 1;
    a.p_proj = pick_p(a.rows_per_cta * a.Ho, m.cout / 4);
    const size_t bytes = sizeof(float) * (size_t)a.smem_floats;
    static uint64_t attr_t = 0, attr_f = 0;
    KPD_CUDA(ensure_smem_attr(mbconv_kernel<true>, 200 * 1024, attr_t));
    KPD_CUDA(ensure_smem_attr(mbconv_kernel<false>, 200 * 1024, attr_f));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B * C)); cfg.blockDim = dim3(MB_THREADS); cfg.dynamicSmemBytes = bytes; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (mma) KPD_CUDA(cudaLaunchKernelEx(&cfg, mbconv_kernel<true>, a));
    else KPD_CUDA(cudaLaunchKernelEx(&cfg, mbconv_kernel<false>, a));
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
