// MobileNetV3-Small body (torchvision features[0:13]) and FPN laterals, NHWC fp32.
//   reference: dll/models/backbone.py:247-264 (MobileNetV3Wrapper), :29-37 (LightweightFPN laterals + top-down),
//              tv:models/mobilenetv3.py:53-116 (InvertedResidual), tv:ops/misc.py:225-261 (SqueezeExcitation)
// Kernels: stem 3x3/s2 direct conv, register-tiled SIMT GEMM for every 1x1 conv (expand / project / f12 /
// laterals, with SE-scale on the A operand and residual / nearest-upsample add in the epilogue), depthwise
// kxk with fused SE squeeze (per-CTA reduction + one atomicAdd per channel), SE excitation MLP.
#include "kpd_common.cuh"
#include <stdlib.h>

namespace kpd {

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

int depthwise_batched(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// stem: conv3x3 s2 p1 (Cin -> 16) + folded BN + hardswish.  in NCHW fp32, out NHWC fp32.
// 4 threads per output pixel, each 4 output channels -> one coalesced float4 store per thread.
// One thread per output pixel, all 16 channels: 9 Cin input loads (neighbouring lanes read neighbouring pixels, the three kx taps
// overlap in L1), the 16-wide weight rows as shared-memory broadcasts (4 LDS.128 per tap against 16 FMAs), 64 contiguous bytes
// stored per thread and copy (a warp store covers 2 KB of consecutive pixels).  The kernel is a stream: 4 B in + 2 x 64 B out per
// pixel against 144 FMAs; the four-threads-per-pixel predecessor spent 3x the instructions on index arithmetic and bounds checks
// (147 us at B = 256 against a 72 us HBM floor).
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                   float* __restrict__ out_tf32, int B, int Cin, int S) {
    __shared__ __align__(16) float ws[27 * 16];
    __shared__ __align__(16) float bs[16];
    __shared__ __align__(16) float xpose[8 * 32 * 20];      // per-warp [32 pixels][16 + 4 pad] store-transposition tiles
    for (int i = threadIdx.x; i < 9 * Cin * 16; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < 16) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const uint32_t Ho = S / 2, Hc = Ho >> 2, NBi = ((uint32_t)B + 1u) & ~1u;
    const uint32_t total = (uint32_t)B * Ho * Ho;          // 32-bit index arithmetic (host-checked)
    const uint32_t plane_px = S * S;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {      // block-uniform: whole warps stay together
        const uint32_t pix = base + threadIdx.x;
        const uint32_t pc = min(pix, total - 1);            // lanes past the end compute the last pixel and store nothing
        const uint32_t rowi = pc / Ho, ox = pc - rowi * Ho, b = rowi / Ho, oy = rowi - b * Ho;
        float4 acc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) acc[q] = reinterpret_cast<const float4*>(bs)[q];
        const int iy0 = (int)oy * 2 - 1, ix0 = (int)ox * 2 - 1;   // only the top / left taps can fall outside (S is even)
        for (int c = 0; c < Cin; c++) {
            const float* ip = in + ((size_t)b * Cin + c) * plane_px;
            float v[9];
#pragma unroll
            for (int ky = 0; ky < 3; ky++)
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const int iy = iy0 + ky, ix = ix0 + kx;
                    v[ky * 3 + kx] = (iy >= 0 && ix >= 0) ? __ldg(ip + (uint32_t)iy * (uint32_t)S + (uint32_t)ix) : 0.f;
                }
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const float4* wp = reinterpret_cast<const float4*>(ws + (t * Cin + c) * 16);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 w4 = wp[q];
                    acc[q].x = fmaf(v[t], w4.x, acc[q].x); acc[q].y = fmaf(v[t], w4.y, acc[q].y);
                    acc[q].z = fmaf(v[t], w4.z, acc[q].z); acc[q].w = fmaf(v[t], w4.w, acc[q].w);
                }
            }
        }
        // A thread holds its pixel's 64 bytes; stored as is, every instruction would write 32 half-empty sectors (16 bytes at a 64-byte
        // lane stride).  The warp's 32 x 64 bytes go through a padded shared-memory tile instead and leave as four instructions of 512
        // consecutive bytes (f0) / eight pixels' complete 64-byte runs (the sub-position planes).
        float* tile = xpose + (threadIdx.x >> 5) * (32 * 20);
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float4 r;
            r.x = apply_act(acc[q].x, ACT_HSWISH); r.y = apply_act(acc[q].y, ACT_HSWISH);
            r.z = apply_act(acc[q].z, ACT_HSWISH); r.w = apply_act(acc[q].w, ACT_HSWISH);
            *reinterpret_cast<float4*>(tile + lane * 20 + 4 * q) = r;
        }
        __syncwarp();
        const uint32_t pix0 = pix - (uint32_t)lane;           // first pixel of this warp's run (blockDim % 32 == 0: runs are warp-aligned)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int pl = i * 8 + (lane >> 2), q = lane & 3;  // pixel of the run, float4 of the pixel
            const uint32_t pp = pix0 + (uint32_t)pl;
            if (pp < total) {
                float4 r = *reinterpret_cast<const float4*>(tile + pl * 20 + 4 * q);
                reinterpret_cast<float4*>(out)[(size_t)pp * 4 + q] = r;
                if (out_tf32) {
                    // second copy, rounded to tf32 and stored as 16 sub-position planes [oy&3][ox&3][image][Ho/4][Ho/4][16]: the A operand of
                    // the composed FPN-0 conv (kpd_fpn0.cu)
                    const uint32_t rw = pp / Ho, x = pp - rw * Ho, bb = rw / Ho, y = rw - bb * Ho;
                    const uint32_t plane = (y & 3u) * 4u + (x & 3u);
                    r.x = tf32_rna(r.x); r.y = tf32_rna(r.y); r.z = tf32_rna(r.z); r.w = tf32_rna(r.w);
                    reinterpret_cast<float4*>(out_tf32)[((((size_t)plane * NBi + bb) * Hc + (y >> 2)) * Hc + (x >> 2)) * 4 + q] = r;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Pointwise (1x1) convolution as a SIMT GEMM:  out[p][n] = act( sum_k in[p][k]*se[b(p)][k] * w[k][n] + bias[n] ) (+res)
// Block tile 128 pixels x 64 couts, K chunk 16, 256 threads, thread tile 8x4.
struct PwArgs {
    const float* in; const float* w; const float* bias; float* out;
    int64_t pixels; int cin, cout, act;
    const float* se_scale; int hw;                 // optional per-(image,k) scale on the input
    const float* res; int res_mode;                // 0 none, 1 same-shape add, 2 nearest-upsampled add
    int H, W, up;                                  // res_mode 2: output map H x W, coarse map (H/up) x (W/up)
    int out_nchw;                                  // 1: write out as [B][cout][hw] (API outputs), hw = H*W
    int round_tf32;                                // 1: round the result to tf32 (it feeds a kind::tf32 MMA)
};

__global__ void __launch_bounds__(256) pw_gemm_kernel(PwArgs a) {
    __shared__ float xs[16][128 + 4];
    __shared__ float wsm[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;                 // tx -> 4 couts, ty -> 8 pixels
    const int64_t p0 = (int64_t)blockIdx.x * 128;
    const int n0 = blockIdx.y * 64;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.cin; k0 += 16) {
        // A tile: 128 pixels x 16 k  (each thread loads 8 elements; k fastest for coalescing over NHWC rows)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = tid + i * 256;
            const int kk = e & 15, pp = e >> 4;
            const int64_t p = p0 + pp; const int k = k0 + kk;
            float v = 0.f;
            if (p < a.pixels && k < a.cin) {
                v = __ldg(a.in + p * a.cin + k);
                if (a.se_scale) v *= __ldg(a.se_scale + (p / a.hw) * a.cin + k);
            }
            xs[kk][pp] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + i * 256;
            const int nn = e & 63, kk = e >> 6;
            const int k = k0 + kk, n = n0 + nn;
            wsm[kk][nn] = (k < a.cin && n < a.cout) ? __ldg(a.w + (int64_t)k * a.cout + n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float xv[8], wv[4];
#pragma unroll
            for (int i = 0; i < 8; i++) xv[i] = xs[kk][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) wv[j] = wsm[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t p = p0 + ty * 8 + i;
        if (p >= a.pixels) continue;
        int64_t rbase = 0;
        if (a.res_mode == 1) rbase = p * a.cout;
        else if (a.res_mode == 2) {
            const int x = (int)(p % a.W), y = (int)((p / a.W) % a.H); const int64_t b = p / ((int64_t)a.W * a.H);
            const int Wc = a.W / a.up, Hc = a.H / a.up;
            rbase = ((b * Hc + y / a.up) * Wc + x / a.up) * a.cout;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.cout) continue;
            float v = acc[i][j] + (a.bias ? __ldg(a.bias + n) : 0.f);
            v = apply_act(v, a.act);
            if (a.res_mode) v += __ldg(a.res + rbase + n);
            if (a.round_tf32) v = tf32_rna(v);
            if (a.out_nchw) { const int64_t b = p / a.hw, q = p % a.hw; a.out[(b * a.cout + n) * a.hw + q] = v; }
            else a.out[p * a.cout + n] = v;
        }
    }
}

static int launch_pw(const PwArgs& a, cudaStream_t s) {
    dim3 grid((unsigned)((a.pixels + 127) / 128), (unsigned)((a.cout + 63) / 64));
    pw_gemm_kernel<<<grid, 256, 0, s>>>(a);
    KPD_LAUNCH_CHECK();
    return 0;
}

int pointwise(const float* in, const float* w, const float* bias, float* out, int64_t pixels, int cin, int cout,
              int act, const float* residual, const float* se_scale, int hw_per_image, cudaStream_t s) {
    PwArgs a{}; a.in = in; a.w = w; a.bias = bias; a.out = out; a.pixels = pixels; a.cin = cin; a.cout = cout; a.act = act;
    a.se_scale = se_scale; a.hw = hw_per_image > 0 ? hw_per_image : 1; a.res = residual; a.res_mode = residual ? 1 : 0;
    return launch_pw(a, s);
}

// SE excitation: scale[b][c] = hardsigmoid( W2 relu(W1 mean + b1) + b2 )   (tv:ops/misc.py:243-261).
// One thread-block CLUSTER of SE_CS CTAs per SE_IG images.  The two weight matrices (2 x 332 KB at C = 576) are what the layer
// costs; a cluster reads each of them exactly once, one slice per CTA, so every SM pulls <= 1/8 of a matrix from L2 and the
// whole chip works on one layer:
//   phase 0  CTA r owns a channel slice (C/8 rounded up to whole quads): squeeze means of its channels (row partials added in order)
//   phase 1  fc1 split along K: hid_r[img][j] = sum over the slice of mean[c] W1[c][j]  (W1 rows of the slice: one contiguous block)
//   exchange cluster barrier; every CTA adds the 8 partials in rank order through distributed shared memory (+ b1, ReLU)
//   phase 2  fc2 for the outputs of the slice: W2[:, slice] (+ b2, hardsigmoid) -> scale[img][slice]
// All sums run in a fixed order that depends only on (C, SQ): the trunk stays bit-identical across batch sizes.
// 9 launches took 178 us as one-CTA-per-4-images kernels (latency-bound: 64 CTAs, each streaming both matrices); this form is
// bound by the cluster barrier and two short dependent load chains.
constexpr int SE_CS = 8, SE_IG = 8, SE_THREADS = 256, SE_MAXC = 576, SE_MAXSQ = 144, SE_MAXSLICE = SE_MAXC / SE_CS, SE_MAXPARTS = 8;
static_assert(SE_IG == 8, "the per-k image vectors are read as two float4");

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }

__global__ void __cluster_dims__(SE_CS, 1, 1) __launch_bounds__(SE_THREADS)
se_fc_kernel(const float* __restrict__ pool, int parts, const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale, int B, int C, int SQ, float inv_hw) {
    extern __shared__ __align__(16) float se_dyn[];            // W1 slice [cn][SQ] | W2 slice [SQ][cn]
    __shared__ __align__(16) float m_s[SE_MAXSLICE][SE_IG];    // squeeze means of this CTA's channel slice, [channel][image]
    __shared__ float part_s[SE_THREADS * SE_IG];               // per-thread-part partial sums of a phase
    __shared__ float hp_s[SE_IG][SE_MAXSQ];                    // this CTA's fc1 partial (read by the whole cluster)
    __shared__ __align__(16) float hid_s[SE_MAXSQ][SE_IG];     // full hidden layer, [unit][image]
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    const int b0 = (blockIdx.x / SE_CS) * SE_IG;
    // slices are whole channel quads, so that both weight slices move as 16-byte copies (4-byte cp.async made the kernel L2-request bound)
    const int slice = 4 * ((C + 4 * SE_CS - 1) / (4 * SE_CS)), c0 = (int)rank * slice, cn = max(0, min(slice, C - c0));
    float* w1s = se_dyn; float* w2s = se_dyn + ((slice * SQ + 3) & ~3);
    // ---- the CTA's two weight slices start towards shared memory right away: everything the layer reads is in flight at once
    {
        const int n4 = (cn * SQ) >> 2;                         // W1 rows of the slice are one contiguous block (SQ % 4 == 0: 16-byte copies)
        const float* src = w1 + (size_t)c0 * SQ;
        const uint32_t d1 = smem_u32_generic(w1s);
        for (int i = tid; i < n4; i += SE_THREADS) cp_async16(d1 + 16u * i, src + 4 * i);
        const uint32_t d2 = smem_u32_generic(w2s);
        const int q4 = cn >> 2;
        for (int i = tid; i < SQ * q4; i += SE_THREADS) { const int k = i / q4, c = (i - k * q4) << 2; cp_async16(d2 + 4u * (uint32_t)(k * cn + c), w2 + (size_t)k * C + c0 + c); }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- phase 0: means of the slice (partials of an image's squeeze added top to bottom)
    for (int i = tid; i < SE_IG * cn; i += SE_THREADS) {
        const int img = i / cn, c = i - img * cn;
        float sum = 0.f;
        if (b0 + img < B) {
            const float* pp = pool + (size_t)(b0 + img) * parts * C + c0 + c;
            int t = 0;
            for (; t + 8 <= parts; t += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = __ldg(pp + (size_t)(t + u) * C);
#pragma unroll
                for (int u = 0; u < 8; u++) sum += v[u];
            }
            for (; t < parts; t++) sum += __ldg(pp + (size_t)t * C);
        }
        m_s[c][img] = sum * inv_hw;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- phase 1: fc1 partial over the slice.  thread = (hidden unit j, part of the slice); parts are added in order
    {
        const int np = max(1, min(SE_MAXPARTS, SE_THREADS / SQ)), j = tid % SQ, pt = tid / SQ;
        if (pt < np) {
            float acc[SE_IG];
#pragma unroll
            for (int g = 0; g < SE_IG; g++) acc[g] = 0.f;
            const int per = (cn + np - 1) / np, k0 = pt * per, k1 = min(cn, k0 + per);
#pragma unroll 4
            for (int k = k0; k < k1; k++) {
                const float w = w1s[k * SQ + j];
                const float4 ma = *reinterpret_cast<const float4*>(&m_s[k][0]), mb = *reinterpret_cast<const float4*>(&m_s[k][4]);
                acc[0] = fmaf(ma.x, w, acc[0]); acc[1] = fmaf(ma.y, w, acc[1]); acc[2] = fmaf(ma.z, w, acc[2]); acc[3] = fmaf(ma.w, w, acc[3]);
                acc[4] = fmaf(mb.x, w, acc[4]); acc[5] = fmaf(mb.y, w, acc[5]); acc[6] = fmaf(mb.z, w, acc[6]); acc[7] = fmaf(mb.w, w, acc[7]);
            }
#pragma unroll
            for (int g = 0; g < SE_IG; g++) part_s[(pt * SE_IG + g) * SQ + j] = acc[g];
        }
        __syncthreads();
        for (int i = tid; i < SE_IG * SQ; i += SE_THREADS) {
            const int g = i / SQ, jj = i - g * SQ;
            float h = 0.f;
            for (int q = 0; q < np; q++) h += part_s[(q * SE_IG + g) * SQ + jj];
            hp_s[g][jj] = h;
        }
    }
    // ---- exchange: all partials of the cluster are complete; add them in rank order
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    {
        const uint32_t local = smem_u32_generic(&hp_s[0][0]);
        for (int i = tid; i < SE_IG * SQ; i += SE_THREADS) {
            const int g = i / SQ, jj = i - g * SQ;
            const uint32_t off = local + (uint32_t)(g * SE_MAXSQ + jj) * 4u;
            float v[SE_CS];
#pragma unroll
            for (int r = 0; r < SE_CS; r++) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(off), "r"(r));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v[r]) : "r"(remote) : "memory");
            }
            float h = __ldg(b1 + jj);
#pragma unroll
            for (int r = 0; r < SE_CS; r++) h += v[r];
            hid_s[jj][g] = fmaxf(h, 0.f);
        }
    }
    // no CTA may exit (or overwrite hp_s) while a peer still reads its partial
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- phase 2: fc2 + hardsigmoid for the channels of the slice.  thread = (channel, part of the hidden units)
    if (cn > 0) {
        const int np = max(1, min(SE_MAXPARTS, SE_THREADS / cn)), c = tid % cn, pt = tid / cn;
        if (pt < np) {
            float acc[SE_IG];
#pragma unroll
            for (int g = 0; g < SE_IG; g++) acc[g] = 0.f;
            const int per = (SQ + np - 1) / np, k0 = pt * per, k1 = min(SQ, k0 + per);
#pragma unroll 4
            for (int k = k0; k < k1; k++) {
                const float w = w2s[k * cn + c];
                const float4 ha = *reinterpret_cast<const float4*>(&hid_s[k][0]), hb = *reinterpret_cast<const float4*>(&hid_s[k][4]);
                acc[0] = fmaf(ha.x, w, acc[0]); acc[1] = fmaf(ha.y, w, acc[1]); acc[2] = fmaf(ha.z, w, acc[2]); acc[3] = fmaf(ha.w, w, acc[3]);
                acc[4] = fmaf(hb.x, w, acc[4]); acc[5] = fmaf(hb.y, w, acc[5]); acc[6] = fmaf(hb.z, w, acc[6]); acc[7] = fmaf(hb.w, w, acc[7]);
            }
#pragma unroll
            for (int g = 0; g < SE_IG; g++) part_s[(pt * SE_IG + g) * cn + c] = acc[g];
        }
        __syncthreads();
        for (int i = tid; i < SE_IG * cn; i += SE_THREADS) {
            const int g = i / cn, cc = i - g * cn;
            if (b0 + g >= B) continue;
            float a = __ldg(b2 + c0 + cc);
            for (int q = 0; q < np; q++) a += part_s[(q * SE_IG + g) * cn + cc];
            scale[(size_t)(b0 + g) * C + c0 + cc] = apply_act(a, ACT_HSIGMOID);
        }
    }
}

static int launch_se_fc(const MBBlock& m, const float* pool, int parts, float* scale, int B, float inv_hw, cudaStream_t s) {
    if (m.exp > SE_MAXC || (m.exp & 3) || m.sq > SE_MAXSQ || m.sq < 4 || (m.sq & 3) || m.sq > SE_THREADS) { set_error("se_fc: unsupported SE shape %d -> %d", m.exp, m.sq); return KPD_ERR_UNSUPPORTED; }
    const int clusters = (B + SE_IG - 1) / SE_IG;
    const int slice = 4 * ((m.exp + 4 * SE_CS - 1) / (4 * SE_CS));
    const size_t smem = sizeof(float) * (size_t)(((slice * m.sq + 3) & ~3) + slice * m.sq + 4);
    static uint64_t attr = 0;
    KPD_CUDA(ensure_smem_attr(se_fc_kernel, 100 * 1024, attr));
    se_fc_kernel<<<clusters * SE_CS, SE_THREADS, smem, s>>>(pool, parts, m.w_fc1, m.b_fc1, m.w_fc2, m.b_fc2, scale, B, m.exp, m.sq, inv_hw);
    KPD_LAUNCH_CHECK();
    return 0;
}

// Batch-wide trunk.  Per inverted-residual block: [expand 1x1 + depthwise k x k + squeeze partials] as ONE kernel with the
// expanded tensor on chip (kpd_ed.cu; fp32 / bf16 modes), the SE excitation MLP, and the projection GEMM with the excitation
// applied to its A operand and the residual added in its epilogue (tcgen05 3xTF32, kpd_pwu.cu, or exact CUDA-core fp32,
// kpd_pwf.cu).  Strict mode and bands that do not fit an SM take the separate expand GEMM + TMA-staged depthwise (kpd_dw.cu).
static int run_body_batched(const kpd_plan* p, const float* images, int B, int S, int strict_fp32, int want_f0_tf32, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    char* scratch = ws + L.scratch;
    const ScratchMap sm = scratch_map(B, S, 1);
    float* f0_tf32 = want_f0_tf32 ? reinterpret_cast<float*>(scratch + sm.f0_tf32) : nullptr;
    float* feat[4]; for (int i = 0; i < 4; i++) feat[i] = reinterpret_cast<float*>(ws + L.feat[i]);
    float* expand = reinterpret_cast<float*>(scratch + sm.expand);
    float* dwout = reinterpret_cast<float*>(scratch + sm.dwout);
    float* ping = reinterpret_cast<float*>(scratch + sm.ping);
    float* pong = reinterpret_cast<float*>(scratch + sm.pong);
    float* se_pool = reinterpret_cast<float*>(scratch + sm.se_pool);
    float* se_scale = reinterpret_cast<float*>(scratch + sm.se_scale);
    {
        const int64_t total = (int64_t)B * (S / 2) * (S / 2);
        int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
        ProfScope ps("body.stem", s);
        stem_kernel<<<grid, 256, 0, s>>>(images, p->w_stem, p->b_stem, feat[0], f0_tf32, B, p->in_channels, S);
        KPD_LAUNCH_CHECK();
    }
    const float* cur = feat[0];
    int H = S / 2;
    for (int i = 0; i < 11; i++) {
        const MBBlock& m = p->blocks[i];
        const int Ho = H / m.stride;
        const float* x = cur;
        int pool_parts = 1;
        bool fused = false;
        if (m.w_exp && !strict_fp32) {
            // expand + depthwise (+ squeeze partials) in one kernel, the expanded tensor stays on chip (kpd_ed.cu); layers whose
            // bands do not fit an SM (wide maps of large inputs) fall through to the two-kernel form below
            ProfScope ps("body.expand_dw", s);
            static const char* tags[11] = {"ed.f1", "ed.f2", "ed.f3", "ed.f4", "ed.f5", "ed.f6", "ed.f7", "ed.f8", "ed.f9", "ed.f10", "ed.f11"};
            ProfScope ps2(tags[i], s);
            const int r = expand_depthwise_fused(m, cur, dwout, m.se ? se_pool : nullptr, sm.se_pool_floats, &pool_parts, B, H, p->sm_count, s);
            if (r == 0) fused = true; else if (r != KPD_ERR_UNSUPPORTED) return r;
        }
        if (!fused) {
            if (m.w_exp) {
                PwMma a{}; a.A = cur; a.Wraw = m.w_exp; a.Wu = m.wu_exp; a.bias = m.b_exp; a.out = expand; a.M = (int64_t)B * H * H; a.K = m.cin; a.N = m.exp; a.act = m.act; a.exact = strict_fp32;
                ProfScope ps("body.pointwise", s);
                KPD_TRY(pointwise_gemm(a, p->sm_count, s));
                x = expand;
            }
            ProfScope ps("body.depthwise", s);
            KPD_TRY(depthwise_batched(m, x, dwout, m.se ? se_pool : nullptr, sm.se_pool_floats, &pool_parts, B, H, p->sm_count, s));
        }
        if (m.se) {
            ProfScope ps("body.se", s);
            KPD_TRY(launch_se_fc(m, se_pool, pool_parts, se_scale, B, 1.f / (float)(Ho * Ho), s));
        }
        float* dst = (i == 2) ? feat[1] : (i == 7) ? feat[2] : (cur == ping ? pong : ping);
        PwMma a{}; a.A = dwout; a.Wraw = m.w_proj; a.Wu = m.wu_proj; a.bias = m.b_proj; a.out = dst; a.M = (int64_t)B * Ho * Ho; a.K = m.exp; a.N = m.cout; a.act = ACT_NONE; a.exact = strict_fp32;
        a.se_scale = m.se ? se_scale : nullptr; a.hw = Ho * Ho; a.res = m.res ? cur : nullptr; a.res_mode = m.res ? 1 : 0;
        { ProfScope ps("body.pointwise", s); KPD_TRY(pointwise_gemm(a, p->sm_count, s)); }
        cur = dst; H = Ho;
    }
    PwMma a{}; a.A = cur; a.Wraw = p->w_last; a.Wu = p->wu_last; a.bias = p->b_last; a.out = feat[3]; a.M = (int64_t)B * H * H; a.K = 96; a.N = 576; a.act = ACT_HSWISH; a.exact = strict_fp32;
    ProfScope ps("body.pointwise", s);
    return pointwise_gemm(a, p->sm_count, s);
}

int run_body(const kpd_plan* p, const float* images, int B, int S, int strict_fp32, int want_f0_tf32, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    ProfScope ps("body", s);
    return run_body_batched(p, images, B, S, strict_fp32, want_f0_tf32, ws, L, s);
}

// lateral_convs[0] (1x1, 16 -> 128, no bias) + nearest x`up` upsample-add of the coarser lateral: a pure streaming
// kernel (64 B in, 512 B out per pixel).  One warp per pixel strip; lane l owns output channels 4l..4l+3 and keeps its
// 16x4 weight slice in registers; the 16 input channels of a pixel are a warp-uniform (broadcast) 64-byte load.
__global__ void __launch_bounds__(256) lateral0_kernel(const float* __restrict__ feat0, const float* __restrict__ w,
                                                       const float* __restrict__ coarse, float* __restrict__ out,
                                                       int64_t pixels, int H, int W, int up, int round_tf32) {
    __shared__ float4 ws4[16][32];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) ws4[i >> 5][i & 31] = __ldg(reinterpret_cast<const float4*>(w) + i);
    __syncthreads();
    // 32-bit index arithmetic throughout (the host checks pixels < 2^31): 64-bit div/mod costs more than the 64 FMAs per pixel
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t Wc = W / up, Hc = H / up, npix = (uint32_t)pixels;
    constexpr int PX = 4;                                  // pixels in flight per warp iteration
    const bool quad = (W % PX == 0) && (up % PX == 0);     // the 4 pixels share one coarse (top-down) pixel
    for (uint32_t p0 = warp * PX; p0 < npix; p0 += nwarps * PX) {
        float4 xv[PX][4], acc[PX];
#pragma unroll
        for (int j = 0; j < PX; j++) {
            const uint32_t p = min(p0 + j, npix - 1);
            const float4* ip = reinterpret_cast<const float4*>(feat0 + (size_t)p * 16);
            xv[j][0] = __ldg(ip); xv[j][1] = __ldg(ip + 1); xv[j][2] = __ldg(ip + 2); xv[j][3] = __ldg(ip + 3);
        }
        if (quad) {
            const uint32_t row = p0 / (uint32_t)W, x = p0 - row * W, b = row / (uint32_t)H, y = row - b * H;
            const float4 cv = __ldg(reinterpret_cast<const float4*>(coarse + ((size_t)(b * Hc + y / up) * Wc + x / up) * 128) + lane);
#pragma unroll
            for (int j = 0; j < PX; j++) acc[j] = cv;
        } else {
#pragma unroll
            for (int j = 0; j < PX; j++) {
                const uint32_t p = min(p0 + j, npix - 1);
                const uint32_t row = p / (uint32_t)W, x = p - row * W, b = row / (uint32_t)H, y = row - b * H;
                acc[j] = __ldg(reinterpret_cast<const float4*>(coarse + ((size_t)(b * Hc + y / up) * Wc + x / up) * 128) + lane);
            }
        }
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
            const float4 w0 = ws4[k4 * 4 + 0][lane], w1 = ws4[k4 * 4 + 1][lane], w2 = ws4[k4 * 4 + 2][lane], w3 = ws4[k4 * 4 + 3][lane];
#pragma unroll
            for (int j = 0; j < PX; j++) {
                const float4 xx = xv[j][k4];
                acc[j].x = fmaf(xx.x, w0.x, acc[j].x); acc[j].y = fmaf(xx.x, w0.y, acc[j].y); acc[j].z = fmaf(xx.x, w0.z, acc[j].z); acc[j].w = fmaf(xx.x, w0.w, acc[j].w);
                acc[j].x = fmaf(xx.y, w1.x, acc[j].x); acc[j].y = fmaf(xx.y, w1.y, acc[j].y); acc[j].z = fmaf(xx.y, w1.z, acc[j].z); acc[j].w = fmaf(xx.y, w1.w, acc[j].w);
                acc[j].x = fmaf(xx.z, w2.x, acc[j].x); acc[j].y = fmaf(xx.z, w2.y, acc[j].y); acc[j].z = fmaf(xx.z, w2.z, acc[j].z); acc[j].w = fmaf(xx.z, w2.w, acc[j].w);
                acc[j].x = fmaf(xx.w, w3.x, acc[j].x); acc[j].y = fmaf(xx.w, w3.y, acc[j].y); acc[j].z = fmaf(xx.w, w3.z, acc[j].z); acc[j].w = fmaf(xx.w, w3.w, acc[j].w);
            }
        }
#pragma unroll
        for (int j = 0; j < PX; j++) {
            if (p0 + j >= npix) break;
            float4 a = acc[j];
            if (round_tf32) { a.x = tf32_rna(a.x); a.y = tf32_rna(a.y); a.z = tf32_rna(a.z); a.w = tf32_rna(a.w); }
            reinterpret_cast<float4*>(out + (size_t)(p0 + j) * 128)[lane] = a;
        }
    }
}

// lateral 0, quad form (W % 4 == 0 and up % 4 == 0: the 4 pixels of an aligned quad share one top-down pixel).  A lane keeps
// its 4 output channels' 16 x 4 weights in registers for the whole kernel; a warp iteration streams one quad: the quad's
// 4 x 16 inputs are fetched as ONE coalesced 256-byte row (8 B per lane) and re-read as shared-memory broadcasts, the next
// quad's input and top-down vectors are prefetched into registers before the 256 FMAs of the current one, and the result
// leaves as four 512-byte rows.  The 1.64 GB fp32 output (B = 256) makes this the trunk's largest HBM stream.
__global__ void __launch_bounds__(256, 2) lateral0_quad_kernel(const float* __restrict__ feat0, const float* __restrict__ w,
                                                               const float* __restrict__ coarse, float* __restrict__ out,
                                                               uint32_t nquads, int H, int W, int up, int round_tf32) {
    __shared__ __align__(16) float xs[8][2][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4 wr[16];
#pragma unroll
    for (int k = 0; k < 16; k++) wr[k] = __ldg(reinterpret_cast<const float4*>(w) + k * 32 + lane);
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t Wc = W / up, Hc = H / up;
    auto fetch = [&](uint32_t q, float2& x, float4& c) {
        x = __ldg(reinterpret_cast<const float2*>(feat0 + (size_t)q * 64) + lane);
        const uint32_t p0 = q * 4, row = p0 / (uint32_t)W, xq = p0 - row * W, b = row / (uint32_t)H, y = row - b * H;
        c = __ldg(reinterpret_cast<const float4*>(coarse + ((size_t)(b * Hc + y / up) * Wc + xq / up) * 128) + lane);
    };
    uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float2 xn = make_float2(0.f, 0.f); float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nquads) fetch(q, xn, cn);
    int buf = 0;
    for (; q < nquads; q += nw) {
        const float4 cc = cn;
        *reinterpret_cast<float2*>(&xs[wid][buf][lane * 2]) = xn;
        __syncwarp();                                       // also orders this write after every lane's reads two iterations back
        if (q + nw < nquads) fetch(q + nw, xn, cn);
        float4 acc[4] = {cc, cc, cc, cc};
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 xx = *reinterpret_cast<const float4*>(&xs[wid][buf][j * 16 + k4 * 4]);      // broadcast
                const float4 w0 = wr[k4 * 4 + 0], w1 = wr[k4 * 4 + 1], w2 = wr[k4 * 4 + 2], w3 = wr[k4 * 4 + 3];
                acc[j].x = fmaf(xx.x, w0.x, acc[j].x); acc[j].y = fmaf(xx.x, w0.y, acc[j].y); acc[j].z = fmaf(xx.x, w0.z, acc[j].z); acc[j].w = fmaf(xx.x, w0.w, acc[j].w);
                acc[j].x = fmaf(xx.y, w1.x, acc[j].x); acc[j].y = fmaf(xx.y, w1.y, acc[j].y); acc[j].z = fmaf(xx.y, w1.z, acc[j].z); acc[j].w = fmaf(xx.y, w1.w, acc[j].w);
                acc[j].x = fmaf(xx.z, w2.x, acc[j].x); acc[j].y = fmaf(xx.z, w2.y, acc[j].y); acc[j].z = fmaf(xx.z, w2.z, acc[j].z); acc[j].w = fmaf(xx.z, w2.w, acc[j].w);
                acc[j].x = fmaf(xx.w, w3.x, acc[j].x); acc[j].y = fmaf(xx.w, w3.y, acc[j].y); acc[j].z = fmaf(xx.w, w3.z, acc[j].z); acc[j].w = fmaf(xx.w, w3.w, acc[j].w);
            }
        }
        float4* op = reinterpret_cast<float4*>(out + (size_t)q * 512) + lane;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float4 a = acc[j];
            if (round_tf32) { a.x = tf32_rna(a.x); a.y = tf32_rna(a.y); a.z = tf32_rna(a.z); a.w = tf32_rna(a.w); }
            op[j * 32] = a;
        }
        buf ^= 1;
    }
}

static int launch_lateral0(const float* feat0, const float* w, const float* coarse, float* out, int64_t pixels, int H, int W, int up,
                           int round_tf32, int sm_count, cudaStream_t s) {
    if ((W % 4 == 0) && (up % 4 == 0)) {
        lateral0_quad_kernel<<<sm_count * 2, 256, 0, s>>>(feat0, w, coarse, out, (uint32_t)(pixels / 4), H, W, up, round_tf32);
    } else {
        const int grid = (int)std::min<int64_t>((pixels + 31) / 32, (int64_t)sm_count * 8);
        lateral0_kernel<<<grid, 256, 0, s>>>(feat0, w, coarse, out, pixels, H, W, up, round_tf32);
    }
    KPD_LAUNCH_CHECK();
    return 0;
}

// FPN laterals + top-down pathway (backbone.py:33-37): lat3 = L3(f3); lat2 = L2(f2)+up2(lat3); lat1 = L1(f1)+up2(lat2);
// lateral0 = L0(f0) + up4(lat1)  -> written to the workspace as the input of fpn_convs[0].
// composed != 0: FPN-0 runs in composed form (kpd_fpn0.cu) -- lateral 0 is not materialised and lateral 1 is rounded to tf32
int run_fpn_laterals(const kpd_plan* p, int B, int S, int round_tf32, int strict_fp32, int composed, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    ProfScope ps("fpn_laterals", s);
    const ScratchMap sm = scratch_map(B, S, 1);
    char* scratch = ws + L.scratch;
    float* lat[4] = {reinterpret_cast<float*>(ws + L.lateral0), reinterpret_cast<float*>(scratch + sm.lat1),
                     reinterpret_cast<float*>(scratch + sm.lat2), reinterpret_cast<float*>(scratch + sm.lat3)};
    const int hs[4] = {S / 2, S / 8, S / 16, S / 32};
    const int cin[4] = {16, 24, 48, 576};
    if ((int64_t)B * hs[0] * hs[0] >= (int64_t)1 << 31) { set_error("fpn laterals: B*H*W = %lld pixels exceeds the 32-bit pixel index", (long long)B * hs[0] * hs[0]); return KPD_ERR_UNSUPPORTED; }
    (void)strict_fp32;                                  // every lateral is exact fp32 on CUDA cores; the modes differ only in round_tf32
    for (int i = 3; i >= 1; i--) {                       // laterals 3..1: fp32 GEMM + nearest-upsample add in the epilogue
        PwMma a{}; a.A = reinterpret_cast<const float*>(ws + L.feat[i]); a.Wraw = p->w_lat[i]; a.Wu = p->wu_lat[i]; a.out = lat[i];
        a.M = (int64_t)B * hs[i] * hs[i]; a.K = cin[i]; a.N = 128; a.act = ACT_NONE; a.hw = hs[i] * hs[i]; a.exact = strict_fp32;
        if (i < 3) { a.res = lat[i + 1]; a.res_mode = 2; a.H = hs[i]; a.W = hs[i]; a.up = hs[i] / hs[i + 1]; }
        a.round_tf32 = (composed && i == 1) ? 1 : 0;
        KPD_TRY(pointwise_gemm(a, p->sm_count, s));
    }
    if (composed) return 0;
    // strict mode only: lateral 0 (1x1, 16 -> 128) + x4 nearest top-down add, materialised by a CUDA-core streaming kernel
    ProfScope ps0("fpn_laterals.lat0", s);
    KPD_TRY(launch_lateral0(reinterpret_cast<const float*>(ws + L.feat[0]), p->w_lat[0], lat[1], lat[0], (int64_t)B * hs[0] * hs[0],
                            hs[0], hs[0], hs[0] / hs[1], round_tf32, p->sm_count, s));
    return 0;
}

// standalone access for the person heads / final 1x1 (NHWC in; NHWC or NCHW out)
int pointwise_ex(const float* in, const float* w, const float* bias, float* out, int64_t pixels, int cin, int cout,
                 int act, int hw, int out_nchw, cudaStream_t s) {
    PwArgs a{}; a.in = in; a.w = w; a.bias = bias; a.out = out; a.pixels = pixels; a.cin = cin; a.cout = cout; a.act = act;
    a.hw = hw > 0 ? hw : 1; a.out_nchw = out_nchw;
    return launch_pw(a, s);
}

}  // namespace kpd
