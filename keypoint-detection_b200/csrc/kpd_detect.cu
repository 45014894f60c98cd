// Person boxes from the level-0 heads on FPN-0: the bare-tensor path of MultiPersonKeypointModel.forward.
//
// NEW BEHAVIOUR -- there is no reference semantics to reproduce.  Upstream routes a batch without 'bboxes' to
// PERSON_HEAD.forward (dll/models/keypoint_model.py:114-119), which returns the raw box-head output of the last feature level
// (dll/models/person_head.py:158-166: no decode, no NMS) and then fails on the 64-channel input; its anchors buffer mixes
// normalised centres with pixel sizes on a fixed 56 x 56 grid and is never read (person_head.py:39-52).  SURVEY.md 8f N3 asks for a
// defined replacement, labelled as such.  What this file defines, using only the reference's own ingredients:
//   candidates  9 anchors per FPN-0 pixel, in the order of person_head.py:45-51: for size in anchor_sizes, for ratio in
//               (0.5, 1, 2): anchor = (cx, cy, size * ratio / S, size / ratio / S) with (cx, cy) the pixel centre, normalised;
//               box head channel 4a + d = delta d of anchor a, cls head channel a = person logit of anchor a
//   decode      cx = acx + dx * aw, cy = acy + dy * ah, w = aw * exp(clamp(dw, -4, 4)), h = ah * exp(clamp(dh, -4, 4)),
//               centres clamped to [0, 1], sizes to (0, 1];  score = sigmoid(logit)
//   selection   per image the top_m candidates by score (ties: lower candidate index), those with score >= conf_threshold
//               (PersonDetectionConfig.conf_threshold), greedy NMS at nms_iou_threshold (the reference's own non_max_suppression,
//               kpd_nms.cu), at most max_persons boxes in descending score order, zero-padded (an all-zero box is the padding
//               convention of the box path, keypoint_model.py:149)
// Kernels: detect_select_kernel (one CTA per image: exact top-m by a 3-pass radix select on the logits' ordered keys, rank sort,
// decode), nms_kernel (kpd_nms.cu), detect_assemble_kernel.  Integer/selection results are exact functions of the head outputs
// (no floating-point atomics, ties resolved by index).
#include "kpd_common.cuh"

namespace kpd {

int nms_launch(const float* boxes, const float* scores, int batch, int m, float iou_thr, int max_out, int32_t* keep, int32_t* keep_count, cudaStream_t s);

constexpr int DET_THREADS = 1024, DET_MAXM = 128, DET_BINS = 2048;

// order-preserving map float -> uint32 (larger float = larger key; NaN sorts lowest)
__device__ __forceinline__ uint32_t det_key(float x) {
    if (x != x) return 0u;
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct DetArgs {
    const float* box; const float* cls;      // [B][HW][36], [B][HW][9]
    int H, W, M;
    float inv_size, a0, a1, a2, conf;
    float* cand_box; float* cand_score; int32_t* n_valid;   // [B][M][4], [B][M], [B]
};

__global__ void __launch_bounds__(DET_THREADS) detect_select_kernel(DetArgs a) {
    __shared__ int hist[DET_BINS];
    __shared__ int warp_tot[32];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_need, s_ngt_slot, s_running;
    __shared__ uint32_t sel_key[DET_MAXM];
    __shared__ int sel_idx[DET_MAXM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, HW = a.H * a.W, ncand = HW * 9;
    const float* cls = a.cls + (size_t)b * ncand;
    const int M = min(a.M, ncand);
    if (tid == 0) { s_prefix = 0u; s_mask = 0u; s_need = M; }
    __syncthreads();
    // ---- exact M-th largest key: radix select over 11 + 11 + 10 bits, most significant first
    for (int pass = 0; pass < 3; pass++) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0), bits = pass == 2 ? 10 : 11, nb = 1 << bits;
        for (int i = tid; i < DET_BINS; i += DET_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int c = tid; c < ncand; c += DET_THREADS) {
            const uint32_t k = det_key(__ldg(cls + c));
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & (uint32_t)(nb - 1)], 1);
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns bins [l * per, (l + 1) * per); suffix sums over the lanes find the lane, then the bin inside it
            const int per = nb / 32;
            int mine = 0;
            for (int i = 0; i < per; i++) mine += hist[lane * per + i];
            int above = 0;                                  // candidates in the bins of higher lanes
            for (int o = 1; o < 32; o++) { const int v = __shfl_down_sync(0xffffffffu, mine, o); if (lane + o < 32) above += v; }
            const int need = s_need;
            const bool here = above < need && above + mine >= need;
            const uint32_t ball = __ballot_sync(0xffffffffu, here);
            if (ball && lane == (int)__ffs(ball) - 1) {
                int cum = above, chosen = lane * per;
                for (int i = per - 1; i >= 0; i--) {
                    const int hcount = hist[lane * per + i];
                    if (cum + hcount >= need) { chosen = lane * per + i; break; }
                    cum += hcount;
                }
                s_need = need - cum;                        // still to take from inside the chosen bin
                s_prefix = prefix | ((uint32_t)chosen << shift);
                s_mask = mask | ((uint32_t)(nb - 1) << shift);
            }
        }
        __syncthreads();
    }
    const uint32_t kstar = s_prefix;                        // key of the M-th best candidate
    const int need_eq = s_need, n_gt = M - need_eq;         // take every key > kstar and the need_eq lowest-index candidates with key == kstar
    if (tid == 0) { s_ngt_slot = 0; s_running = 0; }
    __syncthreads();
    // ---- collection in candidate-index order (block scan over chunks for the ties; the order of the > kstar part is irrelevant,
    // the list is sorted next)
    for (int base = 0; base < ncand; base += DET_THREADS) {
        const int c = base + tid;
        uint32_t k = 0u; bool gt = false, eq = false;
        if (c < ncand) { k = det_key(__ldg(cls + c)); gt = k > kstar; eq = k == kstar; }
        if (gt) { const int slot = atomicAdd(&s_ngt_slot, 1); sel_key[slot] = k; sel_idx[slot] = c; }
        const uint32_t eb = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) warp_tot[warp] = __popc(eb);
        __syncthreads();
        int before = __popc(eb & ((1u << lane) - 1u));
        for (int w2 = 0; w2 < warp; w2++) before += warp_tot[w2];
        const int run0 = s_running;
        if (eq && run0 + before < need_eq) { sel_key[n_gt + run0 + before] = k; sel_idx[n_gt + run0 + before] = c; }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w2 = 0; w2 < DET_THREADS / 32; w2++) t += warp_tot[w2]; s_running = run0 + t; }
        __syncthreads();
    }
    // ---- rank sort (score descending, candidate index ascending), decode, write
    int nv = 0;
    if (tid < M) {
        const uint32_t k = sel_key[tid]; const int c = sel_idx[tid];
        int rank = 0;
        for (int j = 0; j < M; j++) { const uint32_t kj = sel_key[j]; rank += (kj > k || (kj == k && sel_idx[j] < c)) ? 1 : 0; }
        const int pixel = c / 9, an = c - pixel * 9, y = pixel / a.W, x = pixel - y * a.W;
        const int si = an / 3, ri = an - si * 3;
        const float size = si == 0 ? a.a0 : (si == 1 ? a.a1 : a.a2), ratio = ri == 0 ? 0.5f : (ri == 1 ? 1.f : 2.f);
        const float aw = size * ratio * a.inv_size, ah = size / ratio * a.inv_size;
        const float acx = ((float)x + 0.5f) / (float)a.W, acy = ((float)y + 0.5f) / (float)a.H;
        const float* d = a.box + ((size_t)b * HW + pixel) * 36 + an * 4;
        const float cx = fminf(fmaxf(acx + __ldg(d + 0) * aw, 0.f), 1.f), cy = fminf(fmaxf(acy + __ldg(d + 1) * ah, 0.f), 1.f);
        const float w = fminf(aw * expf(fminf(fmaxf(__ldg(d + 2), -4.f), 4.f)), 1.f), h = fminf(ah * expf(fminf(fmaxf(__ldg(d + 3), -4.f), 4.f)), 1.f);
        const float score = 1.f / (1.f + expf(-__ldg(cls + c)));
        float* ob = a.cand_box + ((size_t)b * a.M + rank) * 4;
        ob[0] = cx; ob[1] = cy; ob[2] = w; ob[3] = h;
        a.cand_score[(size_t)b * a.M + rank] = score;
        nv = score >= a.conf ? 1 : 0;
    } else if (tid < a.M) {                                 // fewer candidates than top_m: empty slots sort last
        float* ob = a.cand_box + ((size_t)b * a.M + tid) * 4;
        ob[0] = ob[1] = ob[2] = ob[3] = 0.f;
        a.cand_score[(size_t)b * a.M + tid] = -1.f;
    }
    nv = __syncthreads_count(nv);
    if (tid == 0) a.n_valid[b] = nv;
}

// out[b][p] = p-th kept candidate of image b (NMS order = descending score) while it passed the confidence threshold, else zeros
__global__ void detect_assemble_kernel(const float* __restrict__ cand_box, const int32_t* __restrict__ keep, const int32_t* __restrict__ keep_count,
                                       const int32_t* __restrict__ n_valid, int B, int M, int P, float* __restrict__ out, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * P) return;
    const int b = i / P, p = i - b * P;
    const int kc = min(keep_count[b], P), nv = n_valid[b];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = false;
    if (p < kc) { const int k = keep[(size_t)b * M + p]; if (k >= 0 && k < nv) { v = *reinterpret_cast<const float4*>(cand_box + ((size_t)b * M + k) * 4); ok = true; } }
    *reinterpret_cast<float4*>(out + (size_t)i * 4) = v;
    if (p == 0 && count) {                                  // kept boxes are a prefix (candidates are sorted by score): count them
        int n = 0;
        for (int q = 0; q < kc; q++) { const int k = keep[(size_t)b * M + q]; n += (k >= 0 && k < nv) ? 1 : 0; }
        count[b] = n;
    }
    (void)ok;
}

}  // namespace kpd

using namespace kpd;

extern "C" int kpd_detect_persons(const float* d_box, const float* d_cls, int batch, int height, int width, int size,
                                  const float* anchor_sizes3, float conf_threshold, float iou_threshold, int top_m, int max_persons,
                                  float* d_boxes_out, int32_t* d_count, void* d_scratch, void* stream) {
    KPD_CHECK_ARG(d_box && d_cls && d_boxes_out && d_scratch && anchor_sizes3, "kpd_detect_persons: null pointer");
    KPD_CHECK_ARG(batch >= 1 && height >= 1 && width >= 1 && size >= 1, "kpd_detect_persons: bad shape");
    KPD_CHECK_ARG(top_m >= 1 && top_m <= DET_MAXM && max_persons >= 1 && max_persons <= top_m, "kpd_detect_persons: need 1 <= max_persons <= top_m <= %d", DET_MAXM);
    KPD_CHECK_ARG((int64_t)height * width * 9 < ((int64_t)1 << 31), "kpd_detect_persons: map too large");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // scratch: cand_box [B][M][4] | cand_score [B][M] | keep [B][M] | keep_count [B] | n_valid [B]   (16-byte aligned pieces)
    char* sp = static_cast<char*>(d_scratch);
    const size_t M = (size_t)top_m;
    float* cand_box = reinterpret_cast<float*>(sp); sp += ((size_t)batch * M * 16 + 15) & ~size_t(15);
    float* cand_score = reinterpret_cast<float*>(sp); sp += ((size_t)batch * M * 4 + 15) & ~size_t(15);
    int32_t* keep = reinterpret_cast<int32_t*>(sp); sp += ((size_t)batch * M * 4 + 15) & ~size_t(15);
    int32_t* keep_count = reinterpret_cast<int32_t*>(sp); sp += ((size_t)batch * 4 + 15) & ~size_t(15);
    int32_t* n_valid = reinterpret_cast<int32_t*>(sp);
    DetArgs a{};
    a.box = d_box; a.cls = d_cls; a.H = height; a.W = width; a.M = top_m; a.inv_size = 1.f / (float)size;
    a.a0 = anchor_sizes3[0]; a.a1 = anchor_sizes3[1]; a.a2 = anchor_sizes3[2]; a.conf = conf_threshold;
    a.cand_box = cand_box; a.cand_score = cand_score; a.n_valid = n_valid;
    detect_select_kernel<<<batch, DET_THREADS, 0, s>>>(a);
    KPD_LAUNCH_CHECK();
    KPD_TRY(nms_launch(cand_box, cand_score, batch, top_m, iou_threshold, max_persons, keep, keep_count, s));
    const int total = batch * max_persons;
    detect_assemble_kernel<<<(total + 127) / 128, 128, 0, s>>>(cand_box, keep, keep_count, n_valid, batch, top_m, max_persons, d_boxes_out, d_count);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t kpd_detect_scratch_bytes(int batch, int top_m) {
    const size_t M = (size_t)top_m, B = (size_t)batch;
    return (((B * M * 16 + 15) & ~size_t(15)) + 2 * ((B * M * 4 + 15) & ~size_t(15)) + 2 * ((B * 4 + 15) & ~size_t(15))) + 64;
}
