// Generic fp32 layers.  The reference package exports two classes its model never instantiates -- the
// hand-written `BACKBONE` (dll/models/backbone.py:41-245) and the regression `KEYPOINT_HEAD` (dll/models/keypoint_head.py:9-90).
// Their layer shapes are configurable (width multiplier, branch widths), so they run on shape-generic CUDA-core kernels
// rather than on the packed plans of the hot path: NHWC activations, exact fp32 FMA, fixed summation orders (deterministic,
// batch-size independent).  BatchNorm is folded into the conv weights by the Python wrappers.
#include "kpd_common.cuh"

namespace kpd {
namespace {

__device__ __forceinline__ float layer_act(float x, int act) {
    if (act == 5) return fminf(fmaxf(x, 0.f), 6.f);              // ReLU6
    return apply_act(x, act);
}

// out[n][oy][ox][co] = act2( act( sum_{ky,kx,ci} in[n][iy][ix][g*cpg + ci] * scale[n][..] * w[ky][kx][ci][co] + bias[co] ) + res )
// One thread per (pixel, co): consecutive threads share the input address (broadcast) and read consecutive weights.
__global__ void __launch_bounds__(256)
conv2d_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ in_scale,
              const float* __restrict__ res, float* __restrict__ out, int N, int H, int W, int Cin, int Cout, int K, int stride, int pad,
              int groups, int OH, int OW, int act, int act2) {
    const int64_t total = (int64_t)N * OH * OW * Cout;
    const int cpg = Cin / groups, opg = Cout / groups;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(idx % Cout);
        const int64_t pix = idx / Cout;
        const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((int64_t)OW * OH));
        const int g = co / opg;
        const float* sc = in_scale ? in_scale + (size_t)n * Cin + g * cpg : nullptr;
        float acc = 0.f;
        for (int ky = 0; ky < K; ky++) {
            const int iy = oy * stride - pad + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < K; kx++) {
                const int ix = ox * stride - pad + kx;
                if (ix < 0 || ix >= W) continue;
                const float* ip = in + (((size_t)n * H + iy) * W + ix) * Cin + g * cpg;
                const float* wp = w + ((size_t)(ky * K + kx) * cpg) * Cout + co;
                if (sc) for (int ci = 0; ci < cpg; ci++) acc = fmaf(ip[ci] * sc[ci], wp[(size_t)ci * Cout], acc);
                else    for (int ci = 0; ci < cpg; ci++) acc = fmaf(ip[ci], wp[(size_t)ci * Cout], acc);
            }
        }
        if (bias) acc += bias[co];
        acc = layer_act(acc, act);
        if (res) acc += res[idx];
        out[idx] = layer_act(acc, act2);
    }
}

// torch adaptive_avg_pool2d: bin i covers [floor(i*H/OH), ceil((i+1)*H/OH)); sums in row-major order
__global__ void __launch_bounds__(256)
adaptive_avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C, int OH, int OW) {
    const int64_t total = (int64_t)N * OH * OW * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int64_t pix = idx / C;
        const int ox = (int)(pix % OW), oy = (int)((pix / OW) % OH), n = (int)(pix / ((int64_t)OW * OH));
        const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH, x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
        float s = 0.f;
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) s += in[(((size_t)n * H + y) * W + x) * C + c];
        out[idx] = s / (float)((y1 - y0) * (x1 - x0));
    }
}

// y[m][n] = act(x[m][:] . w[n][:] + b[n]); one warp per output, lanes stride over k, xor-shuffle tree
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y, int M, int Kd, int Nd, int act) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)M * Nd;
    for (int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < total; o += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int n = (int)(o % Nd), m = (int)(o / Nd);
        const float* xp = x + (size_t)m * Kd; const float* wp = w + (size_t)n * Kd;
        float acc = 0.f;
        for (int k = lane; k < Kd; k += 32) acc = fmaf(xp[k], wp[k], acc);
        for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) y[o] = layer_act(acc + (b ? b[n] : 0.f), act);
    }
}

// LayerNorm over the last dimension (biased variance, as torch), one warp per row
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ y, int M, int Nd, int act) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const float* xp = x + (size_t)m * Nd;
    float s = 0.f;
    for (int k = lane; k < Nd; k += 32) s += xp[k];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)Nd;
    float v = 0.f;
    for (int k = lane; k < Nd; k += 32) { const float d = xp[k] - mean; v = fmaf(d, d, v); }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)Nd + eps);
    for (int k = lane; k < Nd; k += 32) y[(size_t)m * Nd + k] = layer_act((xp[k] - mean) * rstd * (gamma ? gamma[k] : 1.f) + (beta ? beta[k] : 0.f), act);
}

// out[n][y][x][c] = x * map[n][y][x]  (spatial attention applied to every channel)
__global__ void __launch_bounds__(256)
mul_map_kernel(const float* __restrict__ x, const float* __restrict__ map, float* __restrict__ out, int64_t pixels, int C) {
    const int64_t total = pixels * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) out[idx] = x[idx] * map[idx / C];
}

// dst[n][y][x][c] += src[n][floor(y*hs/h)][floor(x*ws/w)][c]   (F.interpolate(mode='nearest', size=dst) + in-place add, backbone.py:35-37)
__global__ void __launch_bounds__(256)
upsample_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int N, int H, int W, int HS, int WS, int C) {
    const int64_t total = (int64_t)N * H * W * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int64_t pix = idx / C;
        const int x = (int)(pix % W), y = (int)((pix / W) % H), n = (int)(pix / ((int64_t)W * H));
        // torch nearest: source = min(floor(dst * float(in / out)), in - 1)
        const int sy = min((int)floorf((float)y * ((float)HS / (float)H)), HS - 1), sx = min((int)floorf((float)x * ((float)WS / (float)W)), WS - 1);
        dst[idx] += src[(((size_t)n * HS + sy) * WS + sx) * C + c];
    }
}

inline int grid_for(int64_t total, int per_block) { return (int)std::min<int64_t>((total + per_block - 1) / per_block, 148 * 32); }

}  // namespace
}  // namespace kpd

using namespace kpd;

extern "C" int kpd_conv2d(const float* d_in, const float* d_w, const float* d_bias, const float* d_in_scale, const float* d_residual,
                          float* d_out, int n, int h, int w, int cin, int cout, int k, int stride, int pad, int groups, int act, int act2,
                          void* stream) {
    KPD_CHECK_ARG(d_in && d_w && d_out, "kpd_conv2d: null pointer");
    KPD_CHECK_ARG(n > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && k > 0 && k <= 11 && stride > 0 && pad >= 0, "kpd_conv2d: bad shape");
    KPD_CHECK_ARG(groups > 0 && cin % groups == 0 && cout % groups == 0, "kpd_conv2d: groups %d must divide cin %d and cout %d", groups, cin, cout);
    KPD_CHECK_ARG(act >= 0 && act <= 5 && act2 >= 0 && act2 <= 5, "kpd_conv2d: unknown activation");
    const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
    KPD_CHECK_ARG(oh > 0 && ow > 0, "kpd_conv2d: empty output");
    const int64_t total = (int64_t)n * oh * ow * cout;
    conv2d_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_w, d_bias, d_in_scale, d_residual, d_out, n, h, w, cin, cout, k,
                                                                        stride, pad, groups, oh, ow, act, act2);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_adaptive_avgpool(const float* d_in, float* d_out, int n, int h, int w, int c, int oh, int ow, void* stream) {
    KPD_CHECK_ARG(d_in && d_out && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "kpd_adaptive_avgpool: bad arguments");
    adaptive_avgpool_kernel<<<grid_for((int64_t)n * oh * ow * c, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, h, w, c, oh, ow);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_linear(const float* d_x, const float* d_w, const float* d_b, float* d_y, int m, int k, int n, int act, void* stream) {
    KPD_CHECK_ARG(d_x && d_w && d_y && m > 0 && k > 0 && n > 0 && act >= 0 && act <= 5, "kpd_linear: bad arguments");
    linear_kernel<<<grid_for((int64_t)m * n, 8), 256, 0, (cudaStream_t)stream>>>(d_x, d_w, d_b, d_y, m, k, n, act);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_layernorm(const float* d_x, const float* d_gamma, const float* d_beta, float eps, float* d_y, int m, int n, int act, void* stream) {
    KPD_CHECK_ARG(d_x && d_y && m > 0 && n > 0 && act >= 0 && act <= 5, "kpd_layernorm: bad arguments");
    layernorm_kernel<<<(m + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_x, d_gamma, d_beta, eps, d_y, m, n, act);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_mul_map(const float* d_x, const float* d_map, float* d_out, int64_t pixels, int c, void* stream) {
    KPD_CHECK_ARG(d_x && d_map && d_out && pixels > 0 && c > 0, "kpd_mul_map: bad arguments");
    mul_map_kernel<<<grid_for(pixels * c, 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_map, d_out, pixels, c);
    KPD_LAUNCH_CHECK();
    return 0;
}

extern "C" int kpd_upsample_add(float* d_dst, const float* d_src, int n, int h, int w, int hs, int ws, int c, void* stream) {
    KPD_CHECK_ARG(d_dst && d_src && n > 0 && h > 0 && w > 0 && hs > 0 && ws > 0 && c > 0, "kpd_upsample_add: bad arguments");
    upsample_add_kernel<<<grid_for((int64_t)n * h * w * c, 256), 256, 0, (cudaStream_t)stream>>>(d_dst, d_src, n, h, w, hs, ws, c);
    KPD_LAUNCH_CHECK();
    return 0;
}
