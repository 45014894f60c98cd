// Heat-map -> keypoint decoding, one warp per (instance, joint) map, single pass over HBM.
//   reference: dll/models/keypoint_model.py:250-313 (decode_heatmap/_soft_argmax: soft-argmax + sigmoid(max) 3-class
//              visibility), dll/models/heatmap_head.py:265-296 (argmax), :298-370 (3x3 sub-pixel centroid),
//              :372-413 (temperature soft-argmax); dll/models/keypoint_model.py:230-248 (convert_to_original_coords).
// Per map the warp keeps, per lane: running max + first-max index (strict '>' in increasing index order) and an
// online-softmax triple (m, sum e, sum x*e, sum y*e); lanes are merged with shuffles (ties -> lower index).
#include "kpd_common.cuh"

namespace kpd {

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

struct Soft { float m, s, sx, sy; };

__device__ __forceinline__ void soft_merge(Soft& a, const Soft& b) {
    const float m = fmaxf(a.m, b.m);
    const float ra = (a.m == -INFINITY) ? 0.f : __expf(a.m - m), rb = (b.m == -INFINITY) ? 0.f : __expf(b.m - m);
    a.s = a.s * ra + b.s * rb; a.sx = a.sx * ra + b.sx * rb; a.sy = a.sy * ra + b.sy * rb; a.m = m;
}

template <typename T> struct Elt;
template <> struct Elt<float> {
    static constexpr int VEC = 4;
    __device__ static void load(const float* p, float* v) { const float4 r = ld_stream(reinterpret_cast<const float4*>(p)); v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
    __device__ static float get(const float* p) { return __ldg(p); }
};
template <> struct Elt<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static void load(const __nv_bfloat16* p, float* v) {
        const float4 r = ld_stream(reinterpret_cast<const float4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    __device__ static float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};

struct DecodeArgs {
    const void* hm; int64_t maps; int K, H, W; float inv_temp;
    int32_t* argmax; float* maxval; float* kp_plain; float* kp_subpix; float* score_subpix; float* kp_soft; float* vis;
    // model-forward extras (all optional): ROI -> image coordinates and scatter into padded output slots
    const float* roi_box; const int32_t* roi_slot; const int32_t* roi_count;
    int map_stride_is_slot;      // 1: heat maps live in the padded output [slot][K][H][W], instance n -> slot roi_slot[n]
};

template <typename T>
__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int HW = a.H * a.W;
    int64_t total = a.maps;
    if (a.roi_count) total = (int64_t)(*a.roi_count) * a.K;
    for (int64_t mi = warp_global; mi < total; mi += nwarps) {
        const int64_t n = mi / a.K; const int k = (int)(mi % a.K);
        int64_t dst = n;                                    // output row (instance) index
        if (a.roi_slot) dst = a.roi_slot[n];
        const int64_t src_map = (a.map_stride_is_slot ? dst : n) * a.K + k;
        const T* h = reinterpret_cast<const T*>(a.hm) + src_map * HW;
        constexpr int VEC = Elt<T>::VEC;
        float best = -INFINITY; int best_i = 0x7fffffff;
        Soft sf{-INFINITY, 0.f, 0.f, 0.f};
        const bool want_soft = a.kp_soft != nullptr;
        const int nvec = HW / VEC;                       // the API guarantees HW % VEC == 0
        const int step = 32 * VEC, dxs = step % a.W, dys = step / a.W;
        int x0 = (lane * VEC) % a.W, y0 = (lane * VEC) / a.W;      // coordinates of element 0 of this lane's vector
        constexpr int UN = 4;                            // independent 16-byte loads in flight per lane
        for (int vb = lane; vb < nvec; vb += 32 * UN) {
            float v[UN][VEC];
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int v0 = vb + 32 * u;
                if (v0 < nvec) Elt<T>::load(h + (int64_t)v0 * VEC, v[u]);
                else {
#pragma unroll
                    for (int j = 0; j < VEC; j++) v[u][j] = -INFINITY;
                }
            }
#pragma unroll
            for (int u = 0; u < UN; u++) {
                const int v0 = vb + 32 * u;
#pragma unroll
                for (int j = 0; j < VEC; j++)
                    if (v[u][j] > best) { best = v[u][j]; best_i = v0 * VEC + j; }
                if (want_soft && v0 < nvec) {
                    float vm = v[u][0];
#pragma unroll
                    for (int j = 1; j < VEC; j++) vm = fmaxf(vm, v[u][j]);
                    vm *= a.inv_temp;
                    if (vm > sf.m) { const float r = __expf(sf.m - vm); sf.s *= r; sf.sx *= r; sf.sy *= r; sf.m = vm; }
#pragma unroll
                    for (int j = 0; j < VEC; j++) {
                        int xx = x0 + j, yy = y0;
                        if (xx >= a.W) { xx -= a.W; yy++; }
                        const float e = __expf(fmaf(v[u][j], a.inv_temp, -sf.m));
                        sf.s += e; sf.sx = fmaf(e, (float)xx, sf.sx); sf.sy = fmaf(e, (float)yy, sf.sy);
                    }
                    x0 += dxs; y0 += dys;
                    if (x0 >= a.W) { x0 -= a.W; y0++; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
            if (want_soft) {
                Soft t; t.m = __shfl_xor_sync(0xffffffffu, sf.m, o); t.s = __shfl_xor_sync(0xffffffffu, sf.s, o);
                t.sx = __shfl_xor_sync(0xffffffffu, sf.sx, o); t.sy = __shfl_xor_sync(0xffffffffu, sf.sy, o);
                soft_merge(sf, t);
            }
        }
        if (best_i == 0x7fffffff) best_i = 0;            // all-NaN / empty map guard
        if (lane == 0) {
            const int64_t o = dst * a.K + k;
            const int xr = best_i % a.W, yr = best_i / a.W;
            if (a.argmax) a.argmax[o] = best_i;
            if (a.maxval) a.maxval[o] = best;
            if (a.kp_plain) { a.kp_plain[o * 2] = (float)xr / (float)(a.W - 1); a.kp_plain[o * 2 + 1] = (float)yr / (float)(a.H - 1); }
            if (a.kp_subpix) {
                float tot = 0.f, sx = 0.f, sy = 0.f, wm = -INFINITY;
                for (int yy = max(0, yr - 1); yy < min(a.H, yr + 2); yy++)
                    for (int xx = max(0, xr - 1); xx < min(a.W, xr + 2); xx++) {
                        const float wv = Elt<T>::get(h + yy * a.W + xx);
                        tot += wv; sx = fmaf((float)xx, wv, sx); sy = fmaf((float)yy, wv, sy); wm = fmaxf(wm, wv);
                    }
                float kx = 0.f, ky = 0.f, sc = 0.f;
                if (tot > 0.f) { kx = sx / tot; ky = sy / tot; sc = wm; }
                a.kp_subpix[o * 2] = kx / (float)(a.W - 1); a.kp_subpix[o * 2 + 1] = ky / (float)(a.H - 1);
                if (a.score_subpix) a.score_subpix[o] = sc;
            }
            if (a.kp_soft) {
                float kx = (sf.sx / sf.s) / (float)(a.W - 1), ky = (sf.sy / sf.s) / (float)(a.H - 1);
                if (a.roi_box) {        // convert_to_original_coords: k*wh + (c - wh/2), clamp to [0,1]
                    const float4 bx = reinterpret_cast<const float4*>(a.roi_box)[n];
                    kx = fminf(fmaxf(kx * bx.z + (bx.x - bx.z / 2.f), 0.f), 1.f);
                    ky = fminf(fmaxf(ky * bx.w + (bx.y - bx.w / 2.f), 0.f), 1.f);
                }
                a.kp_soft[o * 2] = kx; a.kp_soft[o * 2 + 1] = ky;
            }
            if (a.vis) {
                const float conf = 1.f / (1.f + expf(-best));
                const int cls = conf < 0.3f ? 0 : (conf < 0.7f ? 1 : 2);
                a.vis[o * 3 + 0] = cls == 0 ? 1.f : 0.f; a.vis[o * 3 + 1] = cls == 1 ? 1.f : 0.f; a.vis[o * 3 + 2] = cls == 2 ? 1.f : 0.f;
            }
        }
    }
}

int run_decode(const DecodeArgs& a, int dtype, int sm_count, cudaStream_t s) {
    ProfScope ps("decode", s);
    const int64_t maps = a.maps;
    if (maps <= 0) return 0;
    const int64_t blocks_needed = (maps + 7) / 8;
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)sm_count * 8);
    if (dtype == 0) decode_kernel<float><<<grid, 256, 0, s>>>(a);
    else decode_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
    KPD_LAUNCH_CHECK();
    return 0;
}

// Dummy-person rule (keypoint_model.py:171-179): an image without a valid box gets vis = [1,0,0] in slot 0.
__global__ void dummy_person_kernel(const int32_t* __restrict__ img_count, float* __restrict__ vis, int B, int P, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K, k = i % K;
    if (img_count[b] == 0) vis[((int64_t)b * P * K + k) * 3] = 1.f;
}

int run_dummy_person(const int32_t* img_count, float* vis, int B, int P, int K, cudaStream_t s) {
    dummy_person_kernel<<<(B * K + 255) / 256, 256, 0, s>>>(img_count, vis, B, P, K);
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd

using namespace kpd;

extern "C" int kpd_decode(const void* d_heatmaps, int dtype, int64_t n, int k, int h, int w, float temperature,
                          int32_t* d_argmax, float* d_maxval, float* d_kp_plain, float* d_kp_subpix,
                          float* d_score_subpix, float* d_kp_soft, float* d_vis, void* stream) {
    KPD_CHECK_ARG(d_heatmaps != nullptr || n == 0, "kpd_decode: null heatmaps");
    KPD_CHECK_ARG(dtype == 0 || dtype == 1, "kpd_decode: dtype must be 0 (fp32) or 1 (bf16)");
    KPD_CHECK_ARG(n >= 0 && k > 0 && h > 1 && w > 1, "kpd_decode: bad shape n=%lld k=%d h=%d w=%d", (long long)n, k, h, w);
    KPD_CHECK_ARG(temperature > 0.f, "kpd_decode: temperature must be > 0");
    KPD_CHECK_ARG(((int64_t)h * w * (dtype ? 2 : 4)) % 16 == 0, "kpd_decode: H*W*elt must be a multiple of 16 bytes");
    KPD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_heatmaps) & 15) == 0, "kpd_decode: heatmaps must be 16-byte aligned");
    if (n == 0) return 0;
    int dev = 0, sms = 148;
    KPD_CUDA(cudaGetDevice(&dev));
    KPD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DecodeArgs a{};
    a.hm = d_heatmaps; a.maps = n * k; a.K = k; a.H = h; a.W = w; a.inv_temp = 1.f / temperature;
    a.argmax = d_argmax; a.maxval = d_maxval; a.kp_plain = d_kp_plain; a.kp_subpix = d_kp_subpix;
    a.score_subpix = d_score_subpix; a.kp_soft = d_kp_soft; a.vis = d_vis;
    return run_decode(a, dtype, sms, static_cast<cudaStream_t>(stream));
}
