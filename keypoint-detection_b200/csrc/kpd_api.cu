// extern "C" forward entry points: orchestration of the hot path on the caller's stream.
//   reference: dll/models/keypoint_model.py:73-206 (MultiPersonKeypointModel.forward with bboxes supplied).
#include "kpd_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace kpd {

struct DecodeArgs {
    const void* hm; int64_t maps; int K, H, W; float inv_temp;
    int32_t* argmax; float* maxval; float* kp_plain; float* kp_subpix; float* score_subpix; float* kp_soft; float* vis;
    const float* roi_box; const int32_t* roi_slot; const int32_t* roi_count;
    int map_stride_is_slot;
};
int run_decode(const DecodeArgs& a, int dtype, int sm_count, cudaStream_t s);
int run_dummy_person(const int32_t* img_count, float* vis, int B, int P, int K, cudaStream_t s);
int run_build_rois(const float* boxes, int B, int P, char* ws, const kpd_workspace_layout& L, cudaStream_t s);
int run_roi_align(int B, int S, int n_max, const int32_t* topk, char* ws, const kpd_workspace_layout& L, cudaStream_t s);
int run_head_attention(const kpd_plan* p, int B, int S, int n_max, int out_bf16, int round_tf32, char* ws, const kpd_workspace_layout& L, cudaStream_t s);
int conv_dense_umma_halo(const kpd_plan* plan, const DenseConv& c, int prec_bf16, const void* in, void* out, int out_bf16,
                         int out_cm, int N, int H, int W, int act, int round_out_tf32, const int32_t* n_valid, cudaStream_t s);
int pointwise_ex(const float* in, const float* w, const float* bias, float* out, int64_t pixels, int cin, int cout,
                 int act, int hw, int out_nchw, cudaStream_t s);

// final_layer[3] (1x1 conv 64->17 + bias) + sigmoid (heatmap_head.py:40-44,111), written as NCHW fp32 straight into the
// padded model output at the ROI's slot.  thread = pixel.
template <typename T>
__global__ void __launch_bounds__(128) head_final_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const int32_t* __restrict__ roi_slot,
                                                         const int32_t* __restrict__ roi_count, float* __restrict__ out) {
    constexpr int HW = 56 * 56;
    const int n = blockIdx.y;
    if (n >= *roi_count) return;
    __shared__ float ws[64 * 17];
    __shared__ float bs[17];
    for (int i = threadIdx.x; i < 64 * 17; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < 17) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    float acc[17];
#pragma unroll
    for (int k = 0; k < 17; k++) acc[k] = bs[k];
    const T* x = in + ((int64_t)n * HW + pix) * 64;
#pragma unroll 4
    for (int c8 = 0; c8 < 8; c8++) {
        float v[8];
        if constexpr (sizeof(T) == 2) {
            const uint4 u = reinterpret_cast<const uint4*>(x)[c8];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
        } else {
            const float4 a = reinterpret_cast<const float4*>(x)[c8 * 2], b = reinterpret_cast<const float4*>(x)[c8 * 2 + 1];
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int k = 0; k < 17; k++) acc[k] = fmaf(v[i], ws[(c8 * 8 + i) * 17 + k], acc[k]);
    }
    float* o = out + (int64_t)roi_slot[n] * 17 * HW + pix;
#pragma unroll
    for (int k = 0; k < 17; k++) o[(int64_t)k * HW] = 1.f / (1.f + expf(-acc[k]));
}

// round_out: the fp32 result feeds another kind::tf32 MMA -> store it rounded-to-nearest tf32
static int dense_conv(const kpd_plan* p, const DenseConv& c, int precision, int use_bf16, const void* in, void* out,
                      int N, int H, int W, int round_out, const int32_t* n_valid, cudaStream_t s) {
    if (precision == KPD_PREC_FP32_SIMT)
        return conv_dense_simt(c, static_cast<const float*>(in), static_cast<float*>(out), N, H, W, ACT_RELU, n_valid, s);
    return conv_dense_umma(p, c, use_bf16, in, out, use_bf16, N, H, W, ACT_RELU, (round_out && !use_bf16) ? 1 : 0, n_valid, s);
}

static int check_layout(const kpd_plan* plan, const void* ws, const kpd_workspace_layout* L) {
    KPD_CHECK_ARG(plan != nullptr, "null plan");
    KPD_CHECK_ARG(L != nullptr && L->total_bytes > 0, "null / empty workspace layout");
    KPD_CHECK_ARG(ws != nullptr, "null workspace");
    KPD_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
    return 0;
}

// A stage call must use the (batch, size, ROI capacity, precision) its workspace layout was computed for: the offsets are
// only valid for those, and a mismatch would let one stage tensor overrun the next.  n_rois / precision < 0 = not checked.
static int check_stage(const kpd_workspace_layout& L, int B, int S, int n_rois, int precision, const char* who) {
    KPD_CHECK_ARG(B == L.batch, "%s: batch %d != %d the workspace layout was built for", who, B, L.batch);
    KPD_CHECK_ARG(S < 0 || S == L.size, "%s: size %d != %d the workspace layout was built for", who, S, L.size);
    KPD_CHECK_ARG(n_rois < 0 || n_rois <= L.max_rois, "%s: %d ROIs exceed the %d the workspace layout was built for", who, n_rois, L.max_rois);
    KPD_CHECK_ARG(precision < 0 || precision == L.precision, "%s: precision %d != %d the workspace layout was built for", who, precision, L.precision);
    return 0;
}

// NCHW fp32 [n,64,56,56] -> NHWC roi_feat + channel-attention pool partials (HeatmapHead.forward standalone).
__global__ void __launch_bounds__(256) nchw_to_roi_kernel(const float* __restrict__ x, float* __restrict__ roi_feat,
                                                          float* __restrict__ part, int32_t* __restrict__ roi_slot,
                                                          int32_t* __restrict__ roi_count, int n_total) {
    constexpr int HW = 56 * 56;
    const int n = blockIdx.y, c = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int p0 = blockIdx.x * 112;
    float s = 0.f, m = -INFINITY;
    for (int p = p0 + q; p < p0 + 112; p += 4) {
        const float v = x[((int64_t)n * 64 + c) * HW + p];
        roi_feat[((int64_t)n * HW + p) * 64 + c] = v;
        s += v; m = fmaxf(m, v);
    }
    __shared__ float ss[4][64], sm[4][64];
    ss[q][c] = s; sm[q][c] = m;
    __syncthreads();
    if (q == 0) {
        float* o = part + (((int64_t)n * ROI_PARTS + blockIdx.x) * 2) * 64;
        o[c] = (ss[0][c] + ss[1][c]) + (ss[2][c] + ss[3][c]);
        o[64 + c] = fmaxf(fmaxf(sm[0][c], sm[1][c]), fmaxf(sm[2][c], sm[3][c]));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { roi_slot[n] = n; if (n == 0) *roi_count = n_total; }
}

static int run_head_stack(const kpd_plan* plan, int precision, int B_for_scratch, int S, int n_max, char* ws,
                          const kpd_workspace_layout& L, float* d_heatmap, cudaStream_t s) {
    const int use_bf16 = precision == KPD_PREC_BF16;
    const int32_t* n_valid = reinterpret_cast<const int32_t*>(ws + L.roi_count);
    KPD_TRY(run_head_attention(plan, B_for_scratch, S, n_max, use_bf16, precision == KPD_PREC_TF32, ws, L, s));
    {
        ProfScope ps("head_conv1", s);
        KPD_TRY(dense_conv(plan, plan->head_conv[0], precision, use_bf16, ws + L.head_in, ws + L.head_a, n_max, 56, 56, 1, n_valid, s));
    }
    {
        ProfScope ps("head_conv2", s);
        KPD_TRY(dense_conv(plan, plan->head_conv[1], precision, use_bf16, ws + L.head_a, ws + L.head_b, n_max, 56, 56, 1, n_valid, s));
    }
    const int32_t* slot = reinterpret_cast<const int32_t*>(ws + L.roi_slot);
    if (precision != KPD_PREC_FP32_SIMT) {
        // conv 256 -> 64 with final_layer[3] (1x1 64 -> 17) + sigmoid fused into its epilogue: the heat maps are computed from
        // the fp32 accumulators and written straight into the padded output; head_c is never materialised
        ProfScope ps("head_conv3", s);
        const FinalFuse fin{plan->w_final, plan->b_final, d_heatmap, slot};
        KPD_TRY(conv_dense_umma(plan, plan->head_conv[2], use_bf16, ws + L.head_b, ws + L.head_c, use_bf16, n_max, 56, 56, ACT_RELU, 0, n_valid, s,
                                nullptr, nullptr, 1, &fin));
        return 0;
    }
    {
        ProfScope ps("head_conv3", s);
        KPD_TRY(dense_conv(plan, plan->head_conv[2], precision, use_bf16, ws + L.head_b, ws + L.head_c, n_max, 56, 56, 0, n_valid, s));
    }
    {
        ProfScope ps("head_final", s);
        dim3 grid((56 * 56 + 127) / 128, n_max);
        head_final_kernel<float><<<grid, 128, 0, s>>>(reinterpret_cast<const float*>(ws + L.head_c), plan->w_final, plan->b_final, slot, n_valid, d_heatmap);
        KPD_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace kpd

using namespace kpd;

extern "C" int kpd_backbone_fpn0(const kpd_plan* plan, int precision, const float* d_images, int B, int S,
                                 void* d_workspace, const kpd_workspace_layout* layout, void* stream) {
    KPD_TRY(check_layout(plan, d_workspace, layout));
    KPD_CHECK_ARG(d_images != nullptr, "kpd_backbone_fpn0: null images");
    KPD_CHECK_ARG(B > 0 && B <= 65535 && S > 0 && S % 32 == 0, "kpd_backbone_fpn0: bad batch/size %d/%d", B, S);
    KPD_CHECK_ARG((int64_t)B * (S / 2) * (S / 2) < ((int64_t)1 << 30), "kpd_backbone_fpn0: B*(S/2)^2 = %lld exceeds the 32-bit pixel index of the trunk kernels", (long long)B * (S / 2) * (S / 2));
    KPD_CHECK_ARG(precision >= 0 && precision <= 2, "unknown precision %d", precision);
    const kpd_workspace_layout& L = *layout;
    KPD_TRY(check_stage(L, B, S, -1, precision, "kpd_backbone_fpn0"));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(d_workspace);
    // FPN level 0: composed tcgen05 form (kpd_fpn0.cu; lateral 0 never materialised) in the fp32 / bf16 modes; the strict
    // mode materialises lateral 0 and runs the fp32 CUDA-core conv.  The map stays fp32 in every mode: storing it as bf16 was
    // measured (profiles/r02_error_attribution*.json) to add as much heat-map error as a whole bf16 head conv for 1% of the step.
    const bool strict = precision == KPD_PREC_FP32_SIMT;
    const bool composed = !strict;
    KPD_TRY(run_body(plan, d_images, B, S, strict, composed, ws, L, s));
    KPD_TRY(run_fpn_laterals(plan, B, S, !strict, strict, composed, ws, L, s));
    const ScratchMap sm = scratch_map(B, S, 1);
    float* pool_part = reinterpret_cast<float*>(ws + L.scratch + sm.pool_part);
    const int H0 = S / 2;
    if (composed) {
        {
            ProfScope ps("fpn0_conv", s);
            KPD_TRY(fpn0_fused(plan, reinterpret_cast<const float*>(ws + L.scratch + sm.f0_tf32), reinterpret_cast<const float*>(ws + L.scratch + sm.lat1),
                               reinterpret_cast<float*>(ws + L.fpn0), pool_part, B, S, s));
        }
        KPD_TRY(run_pool_final(pool_part, B, fpn0_fused_parts(S), H0 * H0, reinterpret_cast<float*>(ws + L.pool_avg), reinterpret_cast<float*>(ws + L.pool_max), s));
        return 0;
    }
    {
        ProfScope ps("fpn0_conv", s);
        KPD_TRY(dense_conv(plan, plan->fpn_conv[0], precision, 0, ws + L.lateral0, ws + L.fpn0, B, H0, H0, 0, nullptr, s));
    }
    KPD_TRY(run_pools(reinterpret_cast<const float*>(ws + L.fpn0), B, H0 * H0, reinterpret_cast<float*>(ws + L.pool_avg),
                      reinterpret_cast<float*>(ws + L.pool_max), pool_part, s));
    return 0;
}

extern "C" int kpd_channel_topk(const kpd_plan* plan, int B, void* d_workspace, const kpd_workspace_layout* layout, void* stream) {
    KPD_TRY(check_layout(plan, d_workspace, layout));
    KPD_CHECK_ARG(B > 0, "kpd_channel_topk: batch must be > 0");
    KPD_TRY(check_stage(*layout, B, -1, -1, -1, "kpd_channel_topk"));
    char* ws = static_cast<char*>(d_workspace);
    return run_scores_topk(plan, B, reinterpret_cast<const float*>(ws + layout->pool_avg), reinterpret_cast<const float*>(ws + layout->pool_max),
                           reinterpret_cast<float*>(ws + layout->scores), reinterpret_cast<int32_t*>(ws + layout->topk),
                           static_cast<cudaStream_t>(stream));
}

extern "C" int kpd_roi_head(const kpd_plan* plan, int precision, int B, int S, const float* d_boxes, int P,
                            const int32_t* d_topk_override, float* d_heatmap, float* d_keypoints, float* d_visibilities,
                            void* d_workspace, const kpd_workspace_layout* layout, void* stream) {
    KPD_TRY(check_layout(plan, d_workspace, layout));
    const int n_max = B * P;
    KPD_CHECK_ARG(B > 0 && P > 0, "kpd_roi_head: batch and persons must be > 0");
    KPD_CHECK_ARG(d_boxes && d_heatmap && d_keypoints && d_visibilities, "kpd_roi_head: null pointer");
    KPD_CHECK_ARG(n_max <= 65535, "kpd_roi_head: at most 65535 ROIs per call (got %d)", n_max);
    KPD_CHECK_ARG(precision >= 0 && precision <= 2, "unknown precision %d", precision);
    KPD_CHECK_ARG(layout->elt_bytes == (size_t)(precision == KPD_PREC_BF16 ? 2 : 4), "workspace layout was built for another precision");
    const kpd_workspace_layout& L = *layout;
    KPD_TRY(check_stage(L, B, S, n_max, precision, "kpd_roi_head"));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(d_workspace);
    const int K = plan->num_keypoints;
    const int32_t* n_valid = reinterpret_cast<const int32_t*>(ws + L.roi_count);
    const int32_t* topk = d_topk_override ? d_topk_override : reinterpret_cast<const int32_t*>(ws + L.topk);

    // padded outputs start as zeros (keypoint_model.py:640-651 pad_to_length)
    KPD_CUDA(cudaMemsetAsync(d_heatmap, 0, sizeof(float) * (size_t)n_max * K * 56 * 56, s));
    KPD_CUDA(cudaMemsetAsync(d_keypoints, 0, sizeof(float) * (size_t)n_max * K * 2, s));
    KPD_CUDA(cudaMemsetAsync(d_visibilities, 0, sizeof(float) * (size_t)n_max * K * 3, s));
    KPD_TRY(run_build_rois(d_boxes, B, P, ws, L, s));
    KPD_TRY(run_roi_align(B, S, n_max, topk, ws, L, s));
    KPD_TRY(run_head_stack(plan, precision, B, S, n_max, ws, L, d_heatmap, s));
    DecodeArgs a{};
    a.hm = d_heatmap; a.maps = (int64_t)n_max * K; a.K = K; a.H = 56; a.W = 56; a.inv_temp = 1.f;
    a.kp_soft = d_keypoints; a.vis = d_visibilities;
    a.roi_box = reinterpret_cast<const float*>(ws + L.roi_box); a.roi_slot = reinterpret_cast<const int32_t*>(ws + L.roi_slot);
    a.roi_count = n_valid; a.map_stride_is_slot = 1;
    KPD_TRY(run_decode(a, 0, plan->sm_count, s));
    KPD_TRY(run_dummy_person(reinterpret_cast<const int32_t*>(ws + L.img_count), d_visibilities, B, P, K, s));
    return 0;
}

extern "C" int kpd_heatmap_head(const kpd_plan* plan, int precision, const float* d_x, int n, float* d_heatmaps,
                                float* d_cw, float* d_sw, void* d_workspace, const kpd_workspace_layout* layout, void* stream) {
    KPD_TRY(check_layout(plan, d_workspace, layout));
    KPD_CHECK_ARG(d_x && d_heatmaps && n > 0 && n <= 65535, "kpd_heatmap_head: bad arguments");
    KPD_CHECK_ARG(layout->elt_bytes == (size_t)(precision == KPD_PREC_BF16 ? 2 : 4), "workspace layout was built for another precision");
    const kpd_workspace_layout& L = *layout;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(d_workspace);
    KPD_CHECK_ARG(n <= layout->max_rois, "kpd_heatmap_head: workspace laid out for %d ROIs, got %d", layout->max_rois, n);
    KPD_CHECK_ARG(precision == layout->precision, "kpd_heatmap_head: precision %d != %d the workspace layout was built for", precision, layout->precision);
    const ScratchMap sm = scratch_map(L.batch, L.size, L.max_rois + 1);
    nchw_to_roi_kernel<<<dim3(ROI_PARTS, n), 256, 0, s>>>(d_x, reinterpret_cast<float*>(ws + L.roi_feat),
        reinterpret_cast<float*>(ws + L.scratch + sm.roi_part), reinterpret_cast<int32_t*>(ws + L.roi_slot),
        reinterpret_cast<int32_t*>(ws + L.roi_count), n);
    KPD_LAUNCH_CHECK();
    KPD_TRY(run_head_stack(plan, precision, L.batch, L.size, n, ws, L, d_heatmaps, s));
    if (d_cw) KPD_CUDA(cudaMemcpyAsync(d_cw, ws + L.head_cw, sizeof(float) * (size_t)n * 64, cudaMemcpyDeviceToDevice, s));
    if (d_sw) KPD_CUDA(cudaMemcpyAsync(d_sw, ws + L.head_sw, sizeof(float) * (size_t)n * 56 * 56, cudaMemcpyDeviceToDevice, s));
    return 0;
}

extern "C" int kpd_forward(const kpd_plan* plan, int precision, const float* d_images, int B, int S,
                           const float* d_boxes, int P, const int32_t* d_topk_override,
                           float* d_heatmap, float* d_keypoints, float* d_visibilities,
                           void* d_workspace, size_t workspace_bytes, void* stream) {
    kpd_workspace_layout L;
    KPD_CHECK_ARG(plan != nullptr, "null plan");
    KPD_CHECK_ARG(P > 0, "kpd_forward: persons must be > 0 (the no-box case is handled by the host wrapper)");
    KPD_TRY(kpd_plan_workspace_layout(plan, B, S, B * P, precision, &L));
    if (workspace_bytes < L.total_bytes) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, L.total_bytes); return KPD_ERR_WORKSPACE; }
    KPD_TRY(kpd_backbone_fpn0(plan, precision, d_images, B, S, d_workspace, &L, stream));
    KPD_TRY(kpd_channel_topk(plan, B, d_workspace, &L, stream));
    return kpd_roi_head(plan, precision, B, S, d_boxes, P, d_topk_override, d_heatmap, d_keypoints, d_visibilities,
                        d_workspace, &L, stream);
}

// backbone.py:39 levels 1..3 (dead in the model; computed only for a standalone MobileNetV3Wrapper call).
extern "C" int kpd_fpn_levels123(const kpd_plan* plan, int B, int S, void* d_workspace, const kpd_workspace_layout* layout,
                                 float* const* d_out, void* stream) {
    KPD_TRY(check_layout(plan, d_workspace, layout));
    KPD_CHECK_ARG(d_out != nullptr, "kpd_fpn_levels123: null output table");
    const kpd_workspace_layout& L = *layout;
    KPD_TRY(check_stage(L, B, S, -1, -1, "kpd_fpn_levels123"));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(d_workspace);
    const ScratchMap sm = scratch_map(B, S, 1);
    const size_t lat_off[4] = {0, sm.lat1, sm.lat2, sm.lat3};
    const int hs[4] = {S / 2, S / 8, S / 16, S / 32};
    for (int i = 1; i < 4; i++) {
        if (!d_out[i]) continue;
        const float* lat = reinterpret_cast<const float*>(ws + L.scratch + lat_off[i]);
        if (i == 1) {
            // the trunk keeps lateral 1 rounded to tf32 (it is an operand of the composed FPN-0 MMA); level 1 of the reference is
            // plain fp32, so the lateral is recomputed unrounded here (exact CUDA-core GEMM + x2 top-down add) into the
            // lateral-0 region, which nothing reads after kpd_backbone_fpn0 has finished
            float* exact = reinterpret_cast<float*>(ws + L.lateral0);
            PwMma a{}; a.A = reinterpret_cast<const float*>(ws + L.feat[1]); a.Wraw = plan->w_lat[1]; a.out = exact;
            a.M = (int64_t)B * hs[1] * hs[1]; a.K = 24; a.N = 128; a.act = ACT_NONE; a.hw = hs[1] * hs[1]; a.exact = 1;
            a.res = reinterpret_cast<const float*>(ws + L.scratch + lat_off[2]); a.res_mode = 2; a.H = hs[1]; a.W = hs[1]; a.up = 2;
            KPD_TRY(pointwise_ffma(a, plan->sm_count, s));
            lat = exact;
        }
        KPD_TRY(conv_dense_simt(plan->fpn_conv[i], lat, d_out[i], B, hs[i], hs[i], ACT_RELU, nullptr, s));
    }
    return 0;
}

// Leaf access to the dense 3x3 convolutions (unit tests / micro-benchmarks of the tensor-core kernels).
// conv: 0 = fpn_convs[0], 1..3 = head convs.  impl: 0 = fp32 CUDA cores (NHWC fp32 in/out), 1 = tcgen05 tf32 (NHWC fp32),
// 2 = tcgen05 bf16 (NHWC bf16), 3 = tap-reuse tcgen05 tf32 (chunk-major fp32 in, NHWC fp32 out), 4 = tap-reuse bf16
// (chunk-major bf16 in, NHWC bf16 out).  All with bias + ReLU.
extern "C" int kpd_dense_conv(const kpd_plan* plan, int conv, int impl, const void* d_in, void* d_out, int n, int h, int w, void* stream) {
    KPD_CHECK_ARG(plan && d_in && d_out, "kpd_dense_conv: null argument");
    KPD_CHECK_ARG(conv >= 0 && conv <= 3 && impl >= 0 && impl <= 4 && n > 0 && h > 0 && w > 0, "kpd_dense_conv: bad argument");
    const DenseConv& c = conv == 0 ? plan->fpn_conv[0] : plan->head_conv[conv - 1];
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (impl) {
        case 0: return conv_dense_simt(c, static_cast<const float*>(d_in), static_cast<float*>(d_out), n, h, w, ACT_RELU, nullptr, s);
        case 1: return conv_dense_umma(plan, c, 0, d_in, d_out, 0, n, h, w, ACT_RELU, 0, nullptr, s);
        case 2: return conv_dense_umma(plan, c, 1, d_in, d_out, 1, n, h, w, ACT_RELU, 0, nullptr, s);
        case 3: return conv_dense_umma_halo(plan, c, 0, d_in, d_out, 0, 0, n, h, w, ACT_RELU, 0, nullptr, s);
        default: return conv_dense_umma_halo(plan, c, 1, d_in, d_out, 1, 0, n, h, w, ACT_RELU, 0, nullptr, s);
    }
}
