// FPN level-0 pooling, channel-attention scores + top-64 selection, and the strict-fp32 SIMT dense convolution.
//   reference: dll/models/keypoint_model.py:18-44 (ChannelAttention), :653-661 (select_top_k_channels),
//              dll/models/backbone.py:39 (fpn_convs) / dll/models/heatmap_head.py:49-68 (head convs) for the SIMT conv.
#include "kpd_common.cuh"

namespace kpd {

// ------------------------------------------------------------------------------------------------
// Global avg / max pool of fpn0 [B, HW, 128] (NHWC fp32): deterministic two-level reduction.
// grid (POOL_PARTS, B), block 256 = 2 pixel rows x 128 channels; partial [B][POOL_PARTS][2][128].
__global__ void __launch_bounds__(256) pool_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int HW) {
    const int b = blockIdx.y, part_id = blockIdx.x;
    const int c = threadIdx.x & 127, r = threadIdx.x >> 7;
    const int per = (HW + POOL_PARTS - 1) / POOL_PARTS;
    const int p0 = part_id * per, p1 = min(HW, p0 + per);
    float s = 0.f, m = -INFINITY;
    const float* xb = x + (int64_t)b * HW * 128;
    for (int p = p0 + r; p < p1; p += 2) {
        const float v = __ldg(xb + (int64_t)p * 128 + c);
        s += v; m = fmaxf(m, v);
    }
    __shared__ float ss[2][128], sm[2][128];
    ss[r][c] = s; sm[r][c] = m;
    __syncthreads();
    if (r == 0) {
        float* o = part + (((int64_t)b * POOL_PARTS + part_id) * 2) * 128;
        o[c] = ss[0][c] + ss[1][c];
        o[128 + c] = fmaxf(sm[0][c], sm[1][c]);
    }
}

__global__ void __launch_bounds__(128) pool_final_kernel(const float* __restrict__ part, float* __restrict__ avg,
                                                         float* __restrict__ mx, int HW, int parts) {
    const int b = blockIdx.x, c = threadIdx.x;
    float s = 0.f, m = -INFINITY;
    const float* o = part + ((int64_t)b * parts * 2) * 128 + c;
    int i = 0;
    for (; i + 8 <= parts; i += 8) {               // 16 loads in flight, then the sums in the canonical order (same bits as a serial loop)
        float sv[8], mv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { sv[j] = __ldg(o + (size_t)(i + j) * 256); mv[j] = __ldg(o + (size_t)(i + j) * 256 + 128); }
#pragma unroll
        for (int j = 0; j < 8; j++) { s += sv[j]; m = fmaxf(m, mv[j]); }
    }
    for (; i < parts; i++) { s += __ldg(o + (size_t)i * 256); m = fmaxf(m, __ldg(o + (size_t)i * 256 + 128)); }
    avg[(int64_t)b * 128 + c] = s / (float)HW;
    mx[(int64_t)b * 128 + c] = m;
}

int run_pools(const float* fpn0, int B, int HW, float* avg, float* mx, float* scratch, cudaStream_t s) {
    ProfScope ps("pool", s);
    pool_partial_kernel<<<dim3(POOL_PARTS, B), 256, 0, s>>>(fpn0, scratch, HW);
    KPD_LAUNCH_CHECK();
    pool_final_kernel<<<B, 128, 0, s>>>(scratch, avg, mx, HW, POOL_PARTS);
    KPD_LAUNCH_CHECK();
    return 0;
}

// second level of the pooling fused into the FPN-0 conv epilogue: `parts` per-tile partials per image
int run_pool_final(const float* partials, int B, int parts, int HW, float* avg, float* mx, cudaStream_t s) {
    ProfScope ps("pool", s);
    pool_final_kernel<<<B, 128, 0, s>>>(partials, avg, mx, HW, parts);
    KPD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// scores = sigmoid(MLP(avg) + MLP(max)), MLP = Linear(128,8)+ReLU+Linear(8,128); then top-64 by rank counting
// (descending score, ties -> lower channel index first).  One 128-thread block per image.
__global__ void __launch_bounds__(128) scores_topk_kernel(const float* __restrict__ avg, const float* __restrict__ mx,
                                                          const float* __restrict__ w0, const float* __restrict__ b0,
                                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                                          float* __restrict__ scores, int32_t* __restrict__ topk) {
    __shared__ float va[128], vm[128], ha[8], hm[8], sc[128];
    const int b = blockIdx.x, c = threadIdx.x, lane = c & 31, warp = c >> 5;
    va[c] = avg[(int64_t)b * 128 + c]; vm[c] = mx[(int64_t)b * 128 + c];
    __syncthreads();
    for (int j = warp; j < 8; j += 4) {          // hidden unit j: dot over 128 inputs, one warp each
        float a = 0.f, m = 0.f;
        for (int k = lane; k < 128; k += 32) { const float w = __ldg(w0 + j * 128 + k); a = fmaf(va[k], w, a); m = fmaf(vm[k], w, m); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); m += __shfl_xor_sync(0xffffffffu, m, o); }
        if (lane == 0) { ha[j] = fmaxf(a + b0[j], 0.f); hm[j] = fmaxf(m + b0[j], 0.f); }
    }
    __syncthreads();
    float oa = b2[c], om = b2[c];
#pragma unroll
    for (int j = 0; j < 8; j++) { const float w = __ldg(w2 + c * 8 + j); oa = fmaf(ha[j], w, oa); om = fmaf(hm[j], w, om); }
    const float sv = 1.f / (1.f + expf(-(oa + om)));
    sc[c] = sv; scores[(int64_t)b * 128 + c] = sv;
    __syncthreads();
    // rank on order-preserving integer keys: a strict total order even when a score is NaN (a NaN ranks above every number),
    // so the 64 slots are always all written with distinct channels
    auto key = [](float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
    const uint32_t kv = key(sv);
    int rank = 0;
    for (int j = 0; j < 128; j++) { const uint32_t o = key(sc[j]); rank += (o > kv) || (o == kv && j < c); }
    if (rank < KPD_TOPK) topk[(int64_t)b * KPD_TOPK + rank] = c;
}

int run_scores_topk(const kpd_plan* p, int B, const float* avg, const float* mx, float* scores, int32_t* topk, cudaStream_t s) {
    ProfScope ps("scores_topk", s);
    scores_topk_kernel<<<B, 128, 0, s>>>(avg, mx, p->ca_w0, p->ca_b0, p->ca_w2, p->ca_b2, scores, topk);
    KPD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Strict-fp32 dense convolution on CUDA cores (KPD_PREC_FP32_SIMT): parity debugging reference for the tcgen05 path.
// taps = 9 (3x3 pad 1) or 1.  in NHWC [N,H,W,cin], w [taps][cin][cout], out NHWC [N,H,W,cout].
// Block: 64 pixels (8x8 patch) x 64 couts, 256 threads, thread tile 4 pixels x 4 couts; K chunk = 16 channels of one tap.
__global__ void __launch_bounds__(256) conv_simt_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int N, int H, int W, int cin, int cout, int taps, int act,
                                                        const int32_t* __restrict__ n_valid) {
    const int n = blockIdx.z;
    if (n_valid && n >= *n_valid) return;
    __shared__ float xs[16][64 + 4];
    __shared__ float wsm[16][64 + 4];
    const int tiles_x = (W + 7) / 8;
    const int ty0 = (blockIdx.x / tiles_x) * 8, tx0 = (blockIdx.x % tiles_x) * 8;
    const int n0 = blockIdx.y * 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;     // tx -> 4 couts, ty -> 4 pixels
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    const int ksz = taps == 9 ? 3 : 1, pad = taps == 9 ? 1 : 0;
    for (int t = 0; t < taps; t++) {
        const int dy = t / ksz - pad, dx = t % ksz - pad;
        for (int c0 = 0; c0 < cin; c0 += 16) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int e = tid + i * 256;
                const int kk = e & 15, pp = e >> 4;
                const int y = ty0 + (pp >> 3) + dy, x = tx0 + (pp & 7) + dx;
                float v = 0.f;
                if (y >= 0 && y < H && x >= 0 && x < W && c0 + kk < cin)
                    v = __ldg(in + (((int64_t)n * H + y) * W + x) * cin + c0 + kk);
                xs[kk][pp] = v;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int e = tid + i * 256;
                const int nn = e & 63, kk = e >> 6;
                wsm[kk][nn] = (c0 + kk < cin && n0 + nn < cout) ? __ldg(w + ((int64_t)t * cin + c0 + kk) * cout + n0 + nn) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                float xv[4], wv[4];
#pragma unroll
                for (int i = 0; i < 4; i++) xv[i] = xs[kk][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; j++) wv[j] = wsm[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int pp = ty * 4 + i;
        const int y = ty0 + (pp >> 3), x = tx0 + (pp & 7);
        if (y >= H || x >= W) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = n0 + tx * 4 + j;
            if (o < cout) out[(((int64_t)n * H + y) * W + x) * cout + o] = apply_act(acc[i][j] + __ldg(bias + o), act);
        }
    }
}

int conv_dense_simt(const DenseConv& c, const float* in, float* out, int N, int H, int W, int act,
                    const int32_t* n_valid, cudaStream_t s) {
    if (N <= 0) return 0;
    dim3 grid(((H + 7) / 8) * ((W + 7) / 8), (c.cout + 63) / 64, N);
    conv_simt_kernel<<<grid, 256, 0, s>>>(in, c.w_simt, c.bias, out, N, H, W, c.cin, c.cout, c.taps, act, n_valid);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd

// ------------------------------------------------------------------------------------------------
// Leaf form of ChannelAttention.forward + select_top_k_channels on a caller's NCHW tensor (keypoint_model.py:33-44, :653-661).
namespace kpd {
// global avg / max of one (image, channel) plane per CTA; fixed-order tree => deterministic
__global__ void __launch_bounds__(256) plane_pool_kernel(const float* __restrict__ x, float* __restrict__ avg, float* __restrict__ mx, int HW) {
    const int plane = blockIdx.y * gridDim.x + blockIdx.x;
    const float* p = x + (int64_t)plane * HW;
    float s = 0.f, m = -INFINITY;
    for (int i = threadIdx.x; i < HW; i += 256) { const float v = __ldg(p + i); s += v; m = fmaxf(m, v); }
    __shared__ float ss[256], sm[256];
    ss[threadIdx.x] = s; sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { avg[plane] = ss[0] / (float)HW; mx[plane] = sm[0]; }
}
// out[b][j] = x[b][topk[b][j]] (planes of HW floats)
__global__ void __launch_bounds__(256) gather_planes_kernel(const float* __restrict__ x, const int32_t* __restrict__ topk,
                                                            float* __restrict__ out, int HW) {
    const int b = blockIdx.y, j = blockIdx.x;
    const int c = min(max(topk[b * KPD_TOPK + j], 0), KPD_FPN_CH - 1);
    const float* src = x + ((int64_t)b * KPD_FPN_CH + c) * HW;
    float* dst = out + ((int64_t)b * KPD_TOPK + j) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) dst[i] = __ldg(src + i);
}
}  // namespace kpd

extern "C" int kpd_channel_attention(const kpd_plan* plan, const float* d_feat_nchw, int batch, int hw, float* d_scores,
                                     int32_t* d_topk, float* d_selected, void* d_scratch, void* stream) {
    using namespace kpd;
    KPD_CHECK_ARG(plan && d_feat_nchw && d_scratch, "kpd_channel_attention: null argument");
    KPD_CHECK_ARG(batch > 0 && batch <= 65535 && hw > 0, "kpd_channel_attention: bad batch / plane size %d / %d", batch, hw);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* avg = static_cast<float*>(d_scratch);
    float* mx = avg + (size_t)batch * KPD_FPN_CH;
    float* scores = d_scores ? d_scores : mx + (size_t)batch * KPD_FPN_CH;
    int32_t* topk = d_topk ? d_topk : reinterpret_cast<int32_t*>(mx + 2 * (size_t)batch * KPD_FPN_CH);
    plane_pool_kernel<<<dim3(KPD_FPN_CH, batch), 256, 0, s>>>(d_feat_nchw, avg, mx, hw);
    KPD_LAUNCH_CHECK();
    KPD_TRY(run_scores_topk(plan, batch, avg, mx, scores, topk, s));
    if (d_selected) {
        gather_planes_kernel<<<dim3(KPD_TOPK, batch), 256, 0, s>>>(d_feat_nchw, topk, d_selected, hw);
        KPD_LAUNCH_CHECK();
    }
    return 0;
}
