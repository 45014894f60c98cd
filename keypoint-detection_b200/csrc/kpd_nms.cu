// Person head: box IoU (cx,cy,w,h), greedy NMS, 1x1 heads.
//   reference: dll/models/person_head.py:54-94 (box_iou), :96-139 (non_max_suppression), :27-34 (heads).
// NMS: one CTA per image. Scores are ranked by counting (descending, ties -> lower index; the reference's sort is
// unstable so ties are implementation-defined there), then a sequential sweep in rank order where every thread tests
// its candidates against the newly kept box -- IoU arithmetic is IEEE fp32 in the reference's operation order, so the
// keep list is bit-identical for tie-free inputs.
#include "kpd_common.cuh"

namespace kpd {

int pointwise_ex(const float* in, const float* w, const float* bias, float* out, int64_t pixels, int cin, int cout,
                 int act, int hw, int out_nchw, cudaStream_t s);

__device__ __forceinline__ float iou_cxcywh(const float4 a, const float4 b) {
    const float ax1 = a.x - a.z / 2.f, ay1 = a.y - a.w / 2.f, ax2 = a.x + a.z / 2.f, ay2 = a.y + a.w / 2.f;
    const float bx1 = b.x - b.z / 2.f, by1 = b.y - b.w / 2.f, bx2 = b.x + b.z / 2.f, by2 = b.y + b.w / 2.f;
    const float iw = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float a1 = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    const float a2 = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
    const float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
    return __fdiv_rn(inter, __fadd_rn(uni, 1e-16f));
}

__global__ void box_iou_kernel(const float* __restrict__ b1, int n1, const float* __restrict__ b2, int n2, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 * n2) return;
    out[i] = iou_cxcywh(reinterpret_cast<const float4*>(b1)[i / n2], reinterpret_cast<const float4*>(b2)[i % n2]);
}

constexpr int NMS_MAX = 1024;

__global__ void __launch_bounds__(256) nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int M,
                                                  float thr, int max_out, int32_t* __restrict__ keep, int32_t* __restrict__ keep_count) {
    __shared__ float4 sb[NMS_MAX];       // boxes in rank order
    __shared__ float ssc[NMS_MAX];
    __shared__ int order[NMS_MAX];
    __shared__ unsigned char dead[NMS_MAX];
    __shared__ int s_cur, s_kept;
    const int img = blockIdx.x, tid = threadIdx.x;
    const float* sc = scores + (int64_t)img * M;
    for (int i = tid; i < M; i += blockDim.x) ssc[i] = sc[i];
    __syncthreads();
    for (int i = tid; i < M; i += blockDim.x) {
        const float v = ssc[i]; int r = 0;
        for (int j = 0; j < M; j++) { const float o = ssc[j]; r += (o > v) || (o == v && j < i); }
        order[r] = i;
    }
    __syncthreads();
    for (int r = tid; r < M; r += blockDim.x) { sb[r] = reinterpret_cast<const float4*>(boxes)[(int64_t)img * M + order[r]]; dead[r] = 0; }
    if (tid == 0) { s_cur = 0; s_kept = 0; }
    __syncthreads();
    int32_t* kp = keep + (int64_t)img * M;
    while (true) {
        if (tid == 0) {                 // advance to the next live candidate in rank order
            int c = s_cur;
            while (c < M && dead[c]) c++;
            s_cur = c;
            if (c < M) { kp[s_kept] = order[c]; s_kept++; }
        }
        __syncthreads();
        const int cur = s_cur, kept = s_kept;
        if (cur >= M || (max_out > 0 && kept >= max_out)) break;
        const float4 top = sb[cur];
        for (int r = cur + 1 + tid; r < M; r += blockDim.x)
            if (!dead[r] && !(iou_cxcywh(top, sb[r]) <= thr)) dead[r] = 1;
        __syncthreads();
        if (tid == 0) s_cur = cur + 1;
        __syncthreads();
    }
    __syncthreads();
    const int kept = s_kept;
    for (int i = kept + tid; i < M; i += blockDim.x) kp[i] = -1;
    if (tid == 0) keep_count[img] = kept;
}

// in-library launch (kpd_detect.cu)
int nms_launch(const float* boxes, const float* scores, int batch, int m, float iou_thr, int max_out, int32_t* keep, int32_t* keep_count, cudaStream_t s) {
    if (m > NMS_MAX) { set_error("nms: at most %d candidates per image (got %d)", NMS_MAX, m); return KPD_ERR_ARG; }
    nms_kernel<<<batch, 256, 0, s>>>(boxes, scores, m, iou_thr, max_out, keep, keep_count);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd

using namespace kpd;

extern "C" int kpd_box_iou(const float* d_boxes1, int n1, const float* d_boxes2, int n2, float* d_iou, void* stream) {
    KPD_CHECK_ARG(n1 >= 0 && n2 >= 0, "kpd_box_iou: negative size");
    if (n1 == 0 || n2 == 0) return 0;
    KPD_CHECK_ARG(d_boxes1 && d_boxes2 && d_iou, "kpd_box_iou: null pointer");
    box_iou_kernel<<<(n1 * n2 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_boxes1, n1, d_boxes2, n2, d_iou);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_nms(const float* d_boxes, const float* d_scores, int batch, int m, float iou_thr, int max_out,
                       int32_t* d_keep, int32_t* d_keep_count, void* stream) {
    KPD_CHECK_ARG(batch >= 0 && m >= 0, "kpd_nms: negative size");
    KPD_CHECK_ARG(m <= NMS_MAX, "kpd_nms: at most %d candidates per image (got %d)", NMS_MAX, m);
    if (batch == 0) return 0;
    KPD_CHECK_ARG(d_keep && d_keep_count, "kpd_nms: null output");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (m == 0) { KPD_CUDA(cudaMemsetAsync(d_keep_count, 0, sizeof(int32_t) * batch, s)); return 0; }
    KPD_CHECK_ARG(d_boxes && d_scores, "kpd_nms: null input");
    ProfScope ps("nms", s);
    nms_kernel<<<batch, 256, 0, s>>>(d_boxes, d_scores, m, iou_thr, max_out, d_keep, d_keep_count);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_person_heads(const kpd_plan* plan, int level, const float* d_feat_nhwc, int64_t pixels,
                                float* d_box_out, float* d_cls_out, void* stream) {
    KPD_CHECK_ARG(plan && d_feat_nhwc, "kpd_person_heads: null argument");
    KPD_CHECK_ARG(level >= 0 && level < 4, "kpd_person_heads: level must be 0..3");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ProfScope ps("person_heads", s);
    KPD_CHECK_ARG(pixels > 0, "kpd_person_heads: pixels must be > 0");
    if (d_box_out && d_cls_out && pixels < ((int64_t)1 << 31) - 128) {
        // both heads as one tcgen05 GEMM [pixels x 128] x [128 x 48] with the on-chip 3xTF32 split (fp32-level accuracy);
        // the epilogue routes the 36 box and 9 class columns to their own tensors, the feature map is read once
        PwMma a{};
        a.A = d_feat_nhwc; a.Wu = plan->wu_heads[level]; a.bias = plan->b_heads[level]; a.out = d_box_out; a.out2 = d_cls_out; a.split = 36; a.n2 = 9;
        a.M = pixels; a.K = 128; a.N = 48; a.act = ACT_NONE; a.hw = 1;
        return pointwise_umma(a, plan->sm_count, s);
    }
    if (d_box_out) KPD_TRY(pointwise_ex(d_feat_nhwc, plan->w_box[level], plan->b_box[level], d_box_out, pixels, 128, 36, ACT_NONE, 1, 0, s));
    if (d_cls_out) KPD_TRY(pointwise_ex(d_feat_nhwc, plan->w_cls[level], plan->b_cls[level], d_cls_out, pixels, 128, 9, ACT_NONE, 1, 0, s));
    return 0;
}
