// ROI stage: box compaction, batched legacy ROI-align over the 64 selected FPN-0 channels, and the heat-map head's
// CBAM attention, all without the reference's per-box Python loop.
//   reference: dll/models/keypoint_model.py:137-170 (loop, zero-box skip), :212-228 (extract_roi_features),
//              :630-638 (box_center_to_corners), tv:ops/roi_align.py (aligned=False, sampling_ratio=-1),
//              dll/models/heatmap_head.py:94-101,115-151 (ChannelAttention, SpatialAttention)
#include "kpd_common.cuh"
#include <stdlib.h>

namespace kpd {

// ------------------------------------------------------------------------------------------------
// Order-preserving compaction of the non-zero boxes.  One block; images are walked in chunks of 1024 (thread = image) with a
// block-wide exclusive scan of the per-image counts and a running carry, so any batch size works without dynamic shared memory.
__global__ void __launch_bounds__(1024) build_rois_kernel(const float* __restrict__ boxes, int B, int P,
                                                          float* __restrict__ roi_box, int32_t* __restrict__ roi_batch,
                                                          int32_t* __restrict__ roi_slot, int32_t* __restrict__ img_count,
                                                          int32_t* __restrict__ roi_count) {
    __shared__ int32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int b = b0 + tid;
        int c = 0;
        if (b < B) {
            for (int p = 0; p < P; p++) {
                const float4 v = reinterpret_cast<const float4*>(boxes)[(int64_t)b * P + p];
                c += !(v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f);
            }
            img_count[b] = c;
        }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        int n = carry + incl - c + (warp > 0 ? warp_tot[warp - 1] : 0);
        carry += warp_tot[31];
        if (b < B) {
            int rank = 0;
            for (int p = 0; p < P; p++) {
                const float4 v = reinterpret_cast<const float4*>(boxes)[(int64_t)b * P + p];
                if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
                reinterpret_cast<float4*>(roi_box)[n] = v;
                roi_batch[n] = b; roi_slot[n] = b * P + rank;
                n++; rank++;
            }
        }
        __syncthreads();                                   // warp_tot is rewritten by the next chunk
    }
    if (tid == 0) *roi_count = carry;
}

// ------------------------------------------------------------------------------------------------
// ROI-align (legacy, adaptive sampling) of channels topk[b][0..63] of fpn0 [B,H,W,128] -> roi_feat [N,56,56,64] fp32.
// grid (ROI_PARTS, Nmax): each block handles 2 output rows; a warp handles one output pixel at a time with
// lanes over channels (2 per lane) so the gather of one bilinear tap touches the 4 cache lines of one NHWC pixel.
// Also emits per-block partial sum/max per channel for the head's channel attention.
__global__ void __launch_bounds__(256) roi_align_kernel(const float* __restrict__ fpn0, const int32_t* __restrict__ topk,
                                                        const float* __restrict__ roi_box, const int32_t* __restrict__ roi_batch,
                                                        const int32_t* __restrict__ roi_count, float* __restrict__ roi_feat,
                                                        float* __restrict__ part, int H, int W) {
    const int n = blockIdx.y;
    if (roi_count && n >= *roi_count) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = roi_batch[n];
    const float4 bx = reinterpret_cast<const float4*>(roi_box)[n];
    // box_center_to_corners + scaling to feature pixels (keypoint_model.py:224-226, :630-638)
    const float x1 = fminf(fmaxf(bx.x - bx.z / 2.f, 0.f), 1.f) * (float)W;
    const float y1 = fminf(fmaxf(bx.y - bx.w / 2.f, 0.f), 1.f) * (float)H;
    const float x2 = fminf(fmaxf(bx.x + bx.z / 2.f, 0.f), 1.f) * (float)W;
    const float y2 = fminf(fmaxf(bx.y + bx.w / 2.f, 0.f), 1.f) * (float)H;
    const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
    const float bin_h = rh / 56.f, bin_w = rw / 56.f;
    const int gh = (int)ceilf(rh / 56.f), gw = (int)ceilf(rw / 56.f);
    const float count = (float)max(gh * gw, 1);
    const int c0 = min(max(topk[b * KPD_TOPK + lane], 0), KPD_FPN_CH - 1), c1 = min(max(topk[b * KPD_TOPK + lane + 32], 0), KPD_FPN_CH - 1);   // (a caller-supplied override cannot index out of the pixel)
    const float* fb = fpn0 + (int64_t)b * H * W * 128;
    float s0 = 0.f, s1 = 0.f, m0 = -INFINITY, m1 = -INFINITY;
    const int row0 = blockIdx.x * 2;
    for (int q = warp; q < 2 * 56; q += 8) {
        const int ph = row0 + q / 56, pw = q % 56;
        float a0 = 0.f, a1 = 0.f;
        auto sample = [&](int iy, int ix) {
            float y = y1 + (float)ph * bin_h + ((float)iy + 0.5f) * bin_h / (float)gh;
            float x = x1 + (float)pw * bin_w + ((float)ix + 0.5f) * bin_w / (float)gw;
            if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) return;
            y = fmaxf(y, 0.f); x = fmaxf(x, 0.f);
            int yl = (int)y, yh, xl = (int)x, xh;
            if (yl >= H - 1) { yl = yh = H - 1; y = (float)yl; } else yh = yl + 1;
            if (xl >= W - 1) { xl = xh = W - 1; x = (float)xl; } else xh = xl + 1;
            const float ly = y - (float)yl, hy = 1.f - ly, lx = x - (float)xl, hx = 1.f - lx;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            const float* p1 = fb + ((int64_t)yl * W + xl) * 128; const float* p2 = fb + ((int64_t)yl * W + xh) * 128;
            const float* p3 = fb + ((int64_t)yh * W + xl) * 128; const float* p4 = fb + ((int64_t)yh * W + xh) * 128;
            a0 += w1 * __ldg(p1 + c0) + w2 * __ldg(p2 + c0) + w3 * __ldg(p3 + c0) + w4 * __ldg(p4 + c0);
            a1 += w1 * __ldg(p1 + c1) + w2 * __ldg(p2 + c1) + w3 * __ldg(p3 + c1) + w4 * __ldg(p4 + c1);
        };
        if (gh <= 2 && gw <= 2) {          // S=224 (ROI <= 112 px): fully unrolled so all 32 gathers of a pixel are in flight
#pragma unroll
            for (int iy = 0; iy < 2; iy++)
#pragma unroll
                for (int ix = 0; ix < 2; ix++)
                    if (iy < gh && ix < gw) sample(iy, ix);
        } else {
            for (int iy = 0; iy < gh; iy++)
                for (int ix = 0; ix < gw; ix++) sample(iy, ix);
        }
        a0 /= count; a1 /= count;
        float* o = roi_feat + (((int64_t)n * 56 + ph) * 56 + pw) * 64;
        o[lane] = a0; o[lane + 32] = a1;
        s0 += a0; s1 += a1; m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1);
    }
    __shared__ float ss[8][64], sm[8][64];
    ss[warp][lane] = s0; ss[warp][lane + 32] = s1; sm[warp][lane] = m0; sm[warp][lane + 32] = m1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f, m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; i++) { s += ss[i][threadIdx.x]; m = fmaxf(m, sm[i][threadIdx.x]); }
        float* o = part + (((int64_t)n * ROI_PARTS + blockIdx.x) * 2) * 64;
        o[threadIdx.x] = s; o[64 + threadIdx.x] = m;
    }
}

// ------------------------------------------------------------------------------------------------
// Separable form of the same ROI-align.  The sampling grid of an output pixel is a product grid and every bilinear weight
// is a product (row factor) x (column factor), so
//     out(ph, pw) = 1/count * sum_x Wx[pw](x) * ( sum_y Wy[ph](y) * f(y, x) ),
// where Wy[ph](y) accumulates the row factors hy / ly of the valid samples of output row ph (and likewise Wx).  A CTA
// (two output rows of one ROI) first forms the inner sum for every feature column the ROI spans -- each feature pixel of
// the <= g+1 contributing rows is gathered ONCE per output row instead of once per sample corner (2.6x fewer gathers at
// S = 224) -- and keeps it in shared memory; the outer sum then runs over <= g+1 shared-memory columns per output pixel.
// Results differ from the sample-by-sample order only by fp32 re-association.  Needs g <= RA_MAXG (S <= 784); larger
// maps use roi_align_kernel.
// (Measured alternative, not kept: reading the whole 128-channel pixel as one coalesced row per lane-quad and selecting the 64
//  channels in shared memory -- the 64 selected channels touch every 32-byte sector of a pixel anyway -- doubled the FMAs and
//  turned phase 2 into a conflicted shared-memory gather: 0.65 ms against 0.33 ms for this lane-per-channel gather at C2.)
constexpr int RA_MAXG = 7, RA_TAPS = RA_MAXG + 1;

// Axis footprint of one output bin: first index touched, number of indices, and the accumulated weight of index first+j.
//   start = roi_start + bin_index * bin; samples at start + (i + 0.5) * bin / g, i < g  (tv:ops/roi_align.py:128-144)
__device__ __forceinline__ void ra_axis(float start, float bin, int g, int size, int j, int& first, int& count, float& wj) {
    first = 0; count = 0; wj = 0.f;
    int lo = 0, hi = -1; bool any = false;
    for (int i = 0; i < g; i++) {
        float y = start + ((float)i + 0.5f) * bin / (float)g;
        if (y < -1.f || y > (float)size) continue;
        y = fmaxf(y, 0.f);
        int yl = (int)y, yh;
        if (yl >= size - 1) { yl = yh = size - 1; y = (float)yl; } else yh = yl + 1;
        if (!any) { lo = yl; any = true; }
        hi = yh;
    }
    if (!any) return;
    first = lo; count = hi - lo + 1;
    const int target = lo + j;
    for (int i = 0; i < g; i++) {
        float y = start + ((float)i + 0.5f) * bin / (float)g;
        if (y < -1.f || y > (float)size) continue;
        y = fmaxf(y, 0.f);
        int yl = (int)y, yh;
        if (yl >= size - 1) { yl = yh = size - 1; y = (float)yl; } else yh = yl + 1;
        const float ly = y - (float)yl, hy = 1.f - ly;
        if (yl == target) wj += hy;
        if (yh == target) wj += ly;
    }
}

template <int NR>                                        // NR >= rows / columns a bin can touch (g + 1)
__global__ void __launch_bounds__(256) roi_align_sep_kernel(const float* __restrict__ fpn0, const int32_t* __restrict__ topk,
                                                            const float* __restrict__ roi_box, const int32_t* __restrict__ roi_batch,
                                                            const int32_t* __restrict__ roi_count, float* __restrict__ roi_feat,
                                                            float* __restrict__ part, int H, int W) {
    const int n = blockIdx.y;
    if (roi_count && n >= *roi_count) return;
    extern __shared__ __align__(16) float ra_v[];          // [span <= W][64]: column-wise partial interpolation of one output row
    __shared__ float wxs[56][RA_TAPS];
    __shared__ int xfirst[56], xcount[56];
    __shared__ float wys[2][RA_TAPS];
    __shared__ int yfirst[2], ycount[2], span[2];
    __shared__ float ss[8][64], sm[8][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = roi_batch[n];
    const float4 bx = reinterpret_cast<const float4*>(roi_box)[n];
    // box_center_to_corners + scaling to feature pixels (keypoint_model.py:224-226, :630-638)
    const float x1 = fminf(fmaxf(bx.x - bx.z / 2.f, 0.f), 1.f) * (float)W;
    const float y1 = fminf(fmaxf(bx.y - bx.w / 2.f, 0.f), 1.f) * (float)H;
    const float x2 = fminf(fmaxf(bx.x + bx.z / 2.f, 0.f), 1.f) * (float)W;
    const float y2 = fminf(fmaxf(bx.y + bx.w / 2.f, 0.f), 1.f) * (float)H;
    const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
    const float bin_h = rh / 56.f, bin_w = rw / 56.f;
    const int gh = (int)ceilf(rh / 56.f), gw = (int)ceilf(rw / 56.f);
    const float inv_count = 1.f;                            // (division by count kept as a division below, as in the reference)
    (void)inv_count;
    const float count = (float)max(gh * gw, 1);
    const int row0 = blockIdx.x * 2;
    for (int t = tid; t < 56 * RA_TAPS + 2 * RA_TAPS; t += 256) {
        int first, cnt; float wj;
        if (t < 56 * RA_TAPS) {
            const int pw = t / RA_TAPS, j = t - pw * RA_TAPS;
            ra_axis(x1 + (float)pw * bin_w, bin_w, gw, W, j, first, cnt, wj);
            wxs[pw][j] = wj;
            if (j == 0) { xfirst[pw] = first; xcount[pw] = cnt; }
        } else {
            const int u = t - 56 * RA_TAPS, r = u / RA_TAPS, j = u - r * RA_TAPS;
            ra_axis(y1 + (float)(row0 + r) * bin_h, bin_h, gh, H, j, first, cnt, wj);
            wys[r][j] = wj;
            if (j == 0) { yfirst[r] = first; ycount[r] = cnt; }
        }
    }
    __syncthreads();
    if (tid == 0) {                                         // feature columns the ROI touches (bins are monotonic in pw)
        int lo = W, hi = -1;
        for (int pw = 0; pw < 56; pw++)
            if (xcount[pw] > 0) { lo = min(lo, xfirst[pw]); hi = max(hi, xfirst[pw] + xcount[pw] - 1); }
        span[0] = lo; span[1] = hi;
    }
    __syncthreads();
    const int xs0 = span[0], ncol = span[1] - span[0] + 1;
    const int c0 = min(max(topk[b * KPD_TOPK + lane], 0), KPD_FPN_CH - 1), c1 = min(max(topk[b * KPD_TOPK + lane + 32], 0), KPD_FPN_CH - 1);   // (a caller-supplied override cannot index out of the pixel)
    const float* fb = fpn0 + (int64_t)b * H * W * 128;
    float s0 = 0.f, s1 = 0.f, m0 = -INFINITY, m1 = -INFINITY;
    for (int r = 0; r < 2; r++) {
        const int ph = row0 + r, yf = yfirst[r], nr = ycount[r];
        float wy[NR];
#pragma unroll
        for (int j = 0; j < NR; j++) wy[j] = wys[r][j];
        // phase 1: one warp per feature column (CU columns in flight), lanes over the 64 selected channels
        constexpr int CU = 16 / NR;                        // 32 gathers in flight per lane
        for (int cb = warp * CU; cb < ncol; cb += 8 * CU) {
            float v0[CU][NR], v1[CU][NR];
#pragma unroll
            for (int u = 0; u < CU; u++) {
                const float* px = fb + ((int64_t)yf * W + xs0 + min(cb + u, ncol - 1)) * 128;
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    if (j < nr) { v0[u][j] = __ldg(px + (int64_t)j * W * 128 + c0); v1[u][j] = __ldg(px + (int64_t)j * W * 128 + c1); }
                    else { v0[u][j] = 0.f; v1[u][j] = 0.f; }
                }
            }
#pragma unroll
            for (int u = 0; u < CU; u++) {
                if (cb + u >= ncol) break;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int j = 0; j < NR; j++) { a0 = fmaf(wy[j], v0[u][j], a0); a1 = fmaf(wy[j], v1[u][j], a1); }
                ra_v[(cb + u) * 64 + lane] = a0; ra_v[(cb + u) * 64 + lane + 32] = a1;
            }
        }
        __syncthreads();
        // phase 2: one warp per output pixel
        for (int pw = warp; pw < 56; pw += 8) {
            const int xf = xfirst[pw] - xs0, nc = xcount[pw];
            float a0 = 0.f, a1 = 0.f;
            for (int j = 0; j < nc; j++) {
                const float w = wxs[pw][j];
                a0 = fmaf(w, ra_v[(xf + j) * 64 + lane], a0); a1 = fmaf(w, ra_v[(xf + j) * 64 + lane + 32], a1);
            }
            a0 /= count; a1 /= count;
            float* o = roi_feat + (((int64_t)n * 56 + ph) * 56 + pw) * 64;
            o[lane] = a0; o[lane + 32] = a1;
            s0 += a0; s1 += a1; m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1);
        }
        __syncthreads();
    }
    ss[warp][lane] = s0; ss[warp][lane + 32] = s1; sm[warp][lane] = m0; sm[warp][lane + 32] = m1;
    __syncthreads();
    if (tid < 64) {
        float s = 0.f, m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; i++) { s += ss[i][tid]; m = fmaxf(m, sm[i][tid]); }
        float* o = part + (((int64_t)n * ROI_PARTS + blockIdx.x) * 2) * 64;
        o[tid] = s; o[64 + tid] = m;
    }
}

static int launch_roi_align(const float* fpn0, const int32_t* topk, const float* roi_box, const int32_t* roi_batch, const int32_t* roi_count,
                            float* roi_feat, float* part, int n, int H, int W, cudaStream_t s) {
    const size_t dyn = sizeof(float) * (size_t)W * 64;
    const int gmax = std::max((H + 55) / 56, (W + 55) / 56);
    if (gmax <= RA_MAXG && dyn <= 160 * 1024) {       // else (maps beyond 392 pixels a side): the sample-by-sample kernel
        static uint64_t attr4 = 0, attr8 = 0;
        if (gmax + 1 <= 4) {
            KPD_CUDA(ensure_smem_attr(roi_align_sep_kernel<4>, 160 * 1024, attr4));    // set once per device: the maximum any call may ask for
            roi_align_sep_kernel<4><<<dim3(ROI_PARTS, n), 256, dyn, s>>>(fpn0, topk, roi_box, roi_batch, roi_count, roi_feat, part, H, W);
        } else {
            KPD_CUDA(ensure_smem_attr(roi_align_sep_kernel<8>, 160 * 1024, attr8));
            roi_align_sep_kernel<8><<<dim3(ROI_PARTS, n), 256, dyn, s>>>(fpn0, topk, roi_box, roi_batch, roi_count, roi_feat, part, H, W);
        }
    } else {
        roi_align_kernel<<<dim3(ROI_PARTS, n), 256, 0, s>>>(fpn0, topk, roi_box, roi_batch, roi_count, roi_feat, part, H, W);
    }
    KPD_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Head attention (heatmap_head.py:115-151): cw = sigmoid(MLP(avg)+MLP(max)); x*=cw; sw = sigmoid(conv7x7([mean_c,max_c])+b);
// x*=sw; the result is the input of deconv_layers[0] (fp32/tf32-rounded or bf16 NHWC).
// Two kernels, each over (8-row band, ROI) CTAs so that a ROI is spread over 7 CTAs and every access is a 16-byte vector:
//   head_att_maps_kernel : channel weights (tiny MLP, recomputed per band) and the per-pixel mean/max over channels of x*cw
//   head_att_apply_kernel: 7x7 spatial attention on the band (+3 halo rows of the maps, zero-padded in shared memory) and
//                          the scaled output.
constexpr int HA_ROWS = 8, HA_BANDS = 56 / HA_ROWS, HA_PIX = HA_ROWS * 56;

__global__ void __launch_bounds__(256) head_att_maps_kernel(const float* __restrict__ roi_feat, const float* __restrict__ part,
                                                            const int32_t* __restrict__ roi_count,
                                                            const float* __restrict__ w0, const float* __restrict__ b0,
                                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                                            float* __restrict__ out_cw, float* __restrict__ maps) {
    const int n = blockIdx.y, band = blockIdx.x;
    if (n >= *roi_count) return;
    constexpr int HW = 56 * 56;
    __shared__ float avg[64], mx[64], hid[8];
    __shared__ __align__(16) float cw[64];
    const int tid = threadIdx.x;
    if (tid < 64) {
        float s = 0.f, m = -INFINITY;
        for (int i = 0; i < ROI_PARTS; i++) {
            const float* o = part + (((int64_t)n * ROI_PARTS + i) * 2) * 64;
            s += o[tid]; m = fmaxf(m, o[64 + tid]);
        }
        avg[tid] = s / (float)HW; mx[tid] = m;
    }
    __syncthreads();
    if (tid < 8) {       // hidden units: 0..3 avg branch, 4..7 max branch
        const int j = tid & 3; const float* v = tid < 4 ? avg : mx;
        float a = b0[j];
        for (int k = 0; k < 64; k++) a = fmaf(v[k], w0[j * 64 + k], a);
        hid[tid] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (tid < 64) {
        float oa = b2[tid], om = b2[tid];
#pragma unroll
        for (int j = 0; j < 4; j++) { oa = fmaf(hid[j], w2[tid * 4 + j], oa); om = fmaf(hid[4 + j], w2[tid * 4 + j], om); }
        const float c = 1.f / (1.f + expf(-(oa + om)));
        cw[tid] = c;
        if (band == 0) out_cw[(int64_t)n * 64 + tid] = c;
    }
    __syncthreads();
    // 16 lanes per pixel (4 channels each), 16 pixels per pass, 4 passes in flight
    const int l16 = tid & 15, slot = tid >> 4;
    const float4 c4 = *reinterpret_cast<const float4*>(&cw[l16 * 4]);
    const float4* x = reinterpret_cast<const float4*>(roi_feat + ((int64_t)n * HW + (int64_t)band * HA_PIX) * 64) + l16;
    float* mo = maps + (int64_t)n * 2 * HW + band * HA_PIX;
    for (int p0 = slot; p0 < HA_PIX; p0 += 64) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldg(x + (int64_t)(p0 + 16 * u) * 16);      // HA_PIX = 448 = 7 * 64: no tail
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float a = v[u].x * c4.x, b = v[u].y * c4.y, c = v[u].z * c4.z, d = v[u].w * c4.w;
            float s = (a + b) + (c + d), m = fmaxf(fmaxf(a, b), fmaxf(c, d));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
            if (l16 == 0) { mo[p0 + 16 * u] = s / 64.f; mo[HW + p0 + 16 * u] = m; }
        }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(256, 4) head_att_apply_kernel(const float* __restrict__ roi_feat, const float* __restrict__ maps,
                                                             const int32_t* __restrict__ roi_count, const float* __restrict__ cw_g,
                                                             const float* __restrict__ saw, const float* __restrict__ sab,
                                                             float* __restrict__ out_sw, OutT* __restrict__ out, int round_tf32) {
    const int n = blockIdx.y, band = blockIdx.x;
    if (n >= *roi_count) return;
    constexpr int HW = 56 * 56, MR = HA_ROWS + 6, MW = 56 + 6;
    __shared__ float mmap[2][MR][MW];                       // mean / max maps of the band with a zero-padded 3-pixel halo
    __shared__ float wk[98], sw[HA_PIX];
    __shared__ __align__(16) float cw[64];
    const int tid = threadIdx.x;
    if (tid < 98) wk[tid] = saw[tid];
    if (tid < 64) cw[tid] = cw_g[(int64_t)n * 64 + tid];
    const float* mg = maps + (int64_t)n * 2 * HW;
    const int r0 = band * HA_ROWS - 3;
    for (int i = tid; i < 2 * MR * MW; i += 256) {
        const int ch = i / (MR * MW), rem = i - ch * (MR * MW), ry = rem / MW, rx = rem - ry * MW;
        const int y = r0 + ry, xx = rx - 3;
        (&mmap[0][0][0])[i] = (y >= 0 && y < 56 && xx >= 0 && xx < 56) ? mg[ch * HW + y * 56 + xx] : 0.f;
    }
    __syncthreads();
    const float bias = sab[0];
    for (int p = tid; p < HA_PIX; p += 256) {
        const int y = p / 56, xx = p - y * 56;
        float a = bias;
#pragma unroll
        for (int ky = 0; ky < 7; ky++)
#pragma unroll
            for (int kx = 0; kx < 7; kx++) {
                a = fmaf(mmap[0][y + ky][xx + kx], wk[ky * 7 + kx], a);
                a = fmaf(mmap[1][y + ky][xx + kx], wk[49 + ky * 7 + kx], a);
            }
        const float sv = 1.f / (1.f + expf(-a));
        sw[p] = sv; out_sw[(int64_t)n * HW + band * HA_PIX + p] = sv;
    }
    __syncthreads();
    const int l16 = tid & 15, slot = tid >> 4;
    const float4 c4 = *reinterpret_cast<const float4*>(&cw[l16 * 4]);
    const int64_t base = ((int64_t)n * HW + (int64_t)band * HA_PIX) * 64;
    const float4* x = reinterpret_cast<const float4*>(roi_feat + base) + l16;
    for (int p0 = slot; p0 < HA_PIX; p0 += 64) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldg(x + (int64_t)(p0 + 16 * u) * 16);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + 16 * u;
            const float sv = sw[p];
            float4 r;
            r.x = (v[u].x * c4.x) * sv; r.y = (v[u].y * c4.y) * sv; r.z = (v[u].z * c4.z) * sv; r.w = (v[u].w * c4.w) * sv;
            if constexpr (sizeof(OutT) == 2) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
                uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(out + base + (int64_t)p * 64 + l16 * 4) = pk;
            } else {
                if (round_tf32) { r.x = tf32_rna(r.x); r.y = tf32_rna(r.y); r.z = tf32_rna(r.z); r.w = tf32_rna(r.w); }
                *reinterpret_cast<float4*>(out + base + (int64_t)p * 64 + l16 * 4) = r;
            }
        }
    }
}

int run_build_rois(const float* boxes, int B, int P, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    build_rois_kernel<<<1, 1024, 0, s>>>(boxes, B, P, reinterpret_cast<float*>(ws + L.roi_box),
        reinterpret_cast<int32_t*>(ws + L.roi_batch), reinterpret_cast<int32_t*>(ws + L.roi_slot),
        reinterpret_cast<int32_t*>(ws + L.img_count), reinterpret_cast<int32_t*>(ws + L.roi_count));
    KPD_LAUNCH_CHECK();
    return 0;
}

int run_roi_align(int B, int S, int n_max, const int32_t* topk, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    ProfScope ps("roi_align", s);
    const ScratchMap sm = scratch_map(B, S, n_max + 1);
    return launch_roi_align(reinterpret_cast<const float*>(ws + L.fpn0), topk, reinterpret_cast<const float*>(ws + L.roi_box),
        reinterpret_cast<const int32_t*>(ws + L.roi_batch), reinterpret_cast<const int32_t*>(ws + L.roi_count),
        reinterpret_cast<float*>(ws + L.roi_feat), reinterpret_cast<float*>(ws + L.scratch + sm.roi_part), n_max, S / 2, S / 2, s);
}

int run_head_attention(const kpd_plan* p, int B, int S, int n_max, int out_bf16, int round_tf32, char* ws, const kpd_workspace_layout& L, cudaStream_t s) {
    ProfScope ps("head_attention", s);
    const ScratchMap sm = scratch_map(B, S, n_max + 1);
    const float* part = reinterpret_cast<const float*>(ws + L.scratch + sm.roi_part);
    float* maps = reinterpret_cast<float*>(ws + L.scratch + sm.att_maps);
    const float* feat = reinterpret_cast<const float*>(ws + L.roi_feat);
    const int32_t* cnt = reinterpret_cast<const int32_t*>(ws + L.roi_count);
    float* cw = reinterpret_cast<float*>(ws + L.head_cw);
    const dim3 grid(HA_BANDS, n_max);
    head_att_maps_kernel<<<grid, 256, 0, s>>>(feat, part, cnt, p->hca_w0, p->hca_b0, p->hca_w2, p->hca_b2, cw, maps);
    KPD_LAUNCH_CHECK();
    if (out_bf16)
        head_att_apply_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(feat, maps, cnt, cw, p->sa_w, p->sa_b, reinterpret_cast<float*>(ws + L.head_sw),
                                                                  reinterpret_cast<__nv_bfloat16*>(ws + L.head_in), 0);
    else
        head_att_apply_kernel<float><<<grid, 256, 0, s>>>(feat, maps, cnt, cw, p->sa_w, p->sa_b, reinterpret_cast<float*>(ws + L.head_sw),
                                                          reinterpret_cast<float*>(ws + L.head_in), round_tf32);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd

extern "C" int kpd_roi_align(const float* d_feat, const int32_t* d_channels, const float* d_boxes, const int32_t* d_roi_batch,
                             int n, int height, int width, float* d_out, float* d_partials, void* stream) {
    KPD_CHECK_ARG(n >= 0 && n <= 65535 && height > 0 && width > 0, "kpd_roi_align: bad sizes");
    if (n == 0) return 0;
    KPD_CHECK_ARG(d_feat && d_channels && d_boxes && d_roi_batch && d_out && d_partials, "kpd_roi_align: null pointer");
    return kpd::launch_roi_align(d_feat, d_channels, d_boxes, d_roi_batch, nullptr, d_out, d_partials, n, height, width, static_cast<cudaStream_t>(stream));
}
