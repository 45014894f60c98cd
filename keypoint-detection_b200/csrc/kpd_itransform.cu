// ITransform on the GPU: the reference's per-image OpenCV / Pillow preprocessing (dll/data/transforms.py:9-114)
// as batched byte kernels, bit-exact with the libraries (oracle/itransform.py states every stage and is pinned on them).
//
//   grayscale:  gray -> CLAHE -> { Gaussian 5x5 s1.5 -> median 5x5 -> Canny(100,200) -> dilate,erode,dilate,erode 3x3
//               -> Gaussian 3x3 -> rescale to the maximum } -> 0.7 CLAHE + 0.3 edges -> antialiased bilinear resize -> normalise
//   rgb:        per-channel CLAHE -> Gaussian 3x3 s0.5 -> resize -> ImageNet normalise
//
// All stages are HBM / L2-bound integer work on 1-byte pixels, so the design is about touching every plane once:
//   clahe_lut_kernel      one CTA per (plane, CLAHE tile): gray conversion + shared-memory histogram + clip / redistribute / scan -> LUT
//   clahe_canny_kernel    one CTA per 64x64 tile (+6 halo): CLAHE blend -> Gaussian -> median -> Sobel -> NMS, stage to stage in
//                         shared memory; writes the CLAHE plane and the Canny map (0 candidate / 1 none / 2 strong)
//   hysteresis_kernel     persistent cooperative grid; a warp holds a tile as 64-bit row masks (lane = row) and floods it with
//                         shifts and lane shuffles; dirty-tile rounds separated by a grid barrier
//   morph_kernel          edges -> dilate -> erode -> dilate -> erode on the same row masks -> Gaussian 3x3 + per-image maximum
//   blend_hresize_kernel  rescale LUT + addWeighted on the fly + horizontal pass of Pillow's fixed-point triangle filter
//   vresize_norm_kernel   vertical pass + ToTensor + Normalize -> float [B,C,S,S]
// Border rules differ per stage (reflect-101, replicate, ignore, zero); every shared-memory stage array therefore holds, at
// positions outside the image, the value its consumer's border rule prescribes ("computed at the mapped coordinate").
#include "kpd_common.cuh"
#include <algorithm>
#include <math.h>

namespace kpd {
namespace {

constexpr int IT_THREADS = 256;
constexpr int TH = 32, TW = 64;                 // output tile of the rgb stencil kernel
constexpr int HALO_C = 6;                       // gauss 2 + median 2 + sobel 1 + nms 1

struct ITParams {
    const uint8_t* src; int B, H, W, C;         // source [B][H][W][C] uint8, C in {1,3,4}
    int planes_per_image;                       // 1 (grayscale) or 3 (rgb: one plane per colour channel)
    int to_gray;                                // 1: plane = RGB2GRAY(src) (or the single channel); 0: plane c = src channel c
    int tiles_x, tiles_y, tw, th, clip;         // CLAHE geometry (tile size of the padded image), integer clip limit (0 = off)
    float lut_scale;
};

__device__ __forceinline__ int reflect101(int i, int n) { i = i < 0 ? -i : i; return i >= n ? 2 * (n - 1) - i : i; }
__device__ __forceinline__ int clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
__device__ __forceinline__ uint8_t sat_rint(float v) {                      // saturate_cast<uchar>(float): round half to even
    int r = __float2int_rn(v);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// plane value at (x, y) of image b, plane c: cv2.cvtColor(RGB2GRAY) = (R*9798 + G*19235 + B*3735 + 2^14) >> 15, or the raw channel
__device__ __forceinline__ uint32_t plane_px(const ITParams& p, int b, int c, int y, int x) {
    const uint8_t* s = p.src + ((size_t)((size_t)b * p.H + y) * p.W + x) * p.C;
    if (p.C == 1) return s[0];
    if (!p.to_gray) return s[c];
    return ((uint32_t)s[0] * 9798u + (uint32_t)s[1] * 19235u + (uint32_t)s[2] * 3735u + (1u << 14)) >> 15;
}

// ------------------------------------------------------------------------------------------------ CLAHE LUTs
// One CTA per (plane, tile).  Writes the plane (gray / channel) for the tile's in-image pixels and the tile's 256-entry LUT.
__global__ void __launch_bounds__(IT_THREADS)
clahe_lut_kernel(ITParams p, uint8_t* __restrict__ plane, uint8_t* __restrict__ lut) {
    __shared__ uint32_t hist[8][256];
    __shared__ uint32_t wsum[8];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int ntile = p.tiles_x * p.tiles_y;
    const int pl = blockIdx.x / ntile, tile = blockIdx.x - pl * ntile;
    const int b = pl / p.planes_per_image, c = pl - b * p.planes_per_image;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    for (int i = t; i < 8 * 256; i += IT_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    uint8_t* dst = plane + (size_t)pl * p.H * p.W;
    for (int yy = warp; yy < p.th; yy += IT_THREADS / 32) {            // a warp per tile row, lanes along x
        const int ye = ty * p.th + yy, y = reflect101(ye, p.H);       // copyMakeBorder(BORDER_REFLECT_101) on the bottom / right
        for (int xx = lane; xx < p.tw; xx += 32) {
            const int xe = tx * p.tw + xx, x = reflect101(xe, p.W);
            const uint32_t v = plane_px(p, b, c, y, x);
            if (ye < p.H && xe < p.W) dst[(size_t)y * p.W + x] = (uint8_t)v;
            atomicAdd(&hist[warp][v], 1u);
        }
    }
    __syncthreads();
    uint32_t h = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) h += hist[k][t];                       // thread t owns bin t
    if (p.clip > 0) {
        uint32_t over = h > (uint32_t)p.clip ? h - (uint32_t)p.clip : 0u;
        if (h > (uint32_t)p.clip) h = (uint32_t)p.clip;
        uint32_t s = over;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) wsum[warp] = s;
        __syncthreads();
        uint32_t clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) clipped += wsum[k];
        const uint32_t batch = clipped >> 8;
        uint32_t residual = clipped - (batch << 8);
        h += batch;
        if (residual != 0) {                                           // bins 0, step, 2 step, ... get one more until the residual is spent
            const uint32_t step = max(256u / residual, 1u);
            if (t % step == 0 && t / step < residual) h++;
        }
        __syncthreads();
    }
    // inclusive scan over the 256 bins
    uint32_t s = h;
    for (int o = 1; o < 32; o <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += n; }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < warp; k++) base += wsum[k];
    s += base;
    lut[((size_t)pl * ntile + tile) * 256 + t] = sat_rint(__fmul_rn((float)(int)s, p.lut_scale));
}

// CLAHE value of plane pixel (x, y): float32 bilinear blend of the four neighbouring tile LUTs, no FMA contraction
struct ClaheGeom { int tiles_x, tiles_y; float inv_tw, inv_th; };
__device__ __forceinline__ uint8_t clahe_px(const ClaheGeom& g, const uint8_t* __restrict__ lut, int x, int y, uint32_t v) {
    const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    const float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
    int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
    const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    int tx2 = min(tx1 + 1, g.tiles_x - 1), ty2 = min(ty1 + 1, g.tiles_y - 1);
    tx1 = max(tx1, 0); ty1 = max(ty1, 0);
    const uint8_t* r1 = lut + (size_t)ty1 * g.tiles_x * 256 + v;
    const uint8_t* r2 = lut + (size_t)ty2 * g.tiles_x * 256 + v;
    const float l11 = (float)__ldg(r1 + tx1 * 256), l12 = (float)__ldg(r1 + tx2 * 256);
    const float l21 = (float)__ldg(r2 + tx1 * 256), l22 = (float)__ldg(r2 + tx2 * 256);
    const float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    const float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    return sat_rint(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
}

// ------------------------------------------------------------------------------------------------ median of 25
// The classic 99-exchange selection network for the median of 25 (Devillard, "Fast median search", after Paeth); p[12] is the
// median.  Its correctness for every input follows from the zero-one principle, checked exhaustively over all 2^25 binary inputs
// (tools/check_median_network.py).  Exchanges whose outputs are never read again lose one half at compile time (174 live min / max).
// values are two pixels packed as u16x2: VIMNMX.U16x2 orders both at once
#define IT_CE(a, b) { const uint32_t _lo = __vminu2(a, b); b = __vmaxu2(a, b); a = _lo; }
__device__ __forceinline__ uint32_t median25(uint32_t (&p)[25]) {
    IT_CE(p[0], p[1]) IT_CE(p[3], p[4]) IT_CE(p[2], p[4]) IT_CE(p[2], p[3]) IT_CE(p[6], p[7]) IT_CE(p[5], p[7])
    IT_CE(p[5], p[6]) IT_CE(p[9], p[10]) IT_CE(p[8], p[10]) IT_CE(p[8], p[9]) IT_CE(p[12], p[13]) IT_CE(p[11], p[13])
    IT_CE(p[11], p[12]) IT_CE(p[15], p[16]) IT_CE(p[14], p[16]) IT_CE(p[14], p[15]) IT_CE(p[18], p[19]) IT_CE(p[17], p[19])
    IT_CE(p[17], p[18]) IT_CE(p[21], p[22]) IT_CE(p[20], p[22]) IT_CE(p[20], p[21]) IT_CE(p[23], p[24]) IT_CE(p[2], p[5])
    IT_CE(p[3], p[6]) IT_CE(p[0], p[6]) IT_CE(p[0], p[3]) IT_CE(p[4], p[7]) IT_CE(p[1], p[7]) IT_CE(p[1], p[4])
    IT_CE(p[11], p[14]) IT_CE(p[8], p[14]) IT_CE(p[8], p[11]) IT_CE(p[12], p[15]) IT_CE(p[9], p[15]) IT_CE(p[9], p[12])
    IT_CE(p[13], p[16]) IT_CE(p[10], p[16]) IT_CE(p[10], p[13]) IT_CE(p[20], p[23]) IT_CE(p[17], p[23]) IT_CE(p[17], p[20])
    IT_CE(p[21], p[24]) IT_CE(p[18], p[24]) IT_CE(p[18], p[21]) IT_CE(p[19], p[22]) IT_CE(p[8], p[17]) IT_CE(p[9], p[18])
    IT_CE(p[0], p[18]) IT_CE(p[0], p[9]) IT_CE(p[10], p[19]) IT_CE(p[1], p[19]) IT_CE(p[1], p[10]) IT_CE(p[11], p[20])
    IT_CE(p[2], p[20]) IT_CE(p[2], p[11]) IT_CE(p[12], p[21]) IT_CE(p[3], p[21]) IT_CE(p[3], p[12]) IT_CE(p[13], p[22])
    IT_CE(p[4], p[22]) IT_CE(p[4], p[13]) IT_CE(p[14], p[23]) IT_CE(p[5], p[23]) IT_CE(p[5], p[14]) IT_CE(p[15], p[24])
    IT_CE(p[6], p[24]) IT_CE(p[6], p[15]) IT_CE(p[7], p[16]) IT_CE(p[7], p[19]) IT_CE(p[13], p[21]) IT_CE(p[15], p[23])
    IT_CE(p[7], p[13]) IT_CE(p[7], p[15]) IT_CE(p[1], p[9]) IT_CE(p[3], p[11]) IT_CE(p[5], p[17]) IT_CE(p[11], p[17])
    IT_CE(p[9], p[17]) IT_CE(p[4], p[10]) IT_CE(p[6], p[12]) IT_CE(p[7], p[14]) IT_CE(p[4], p[6]) IT_CE(p[4], p[7])
    IT_CE(p[12], p[14]) IT_CE(p[10], p[14]) IT_CE(p[6], p[7]) IT_CE(p[10], p[12]) IT_CE(p[6], p[10]) IT_CE(p[6], p[17])
    IT_CE(p[12], p[17]) IT_CE(p[7], p[17]) IT_CE(p[7], p[10]) IT_CE(p[12], p[18]) IT_CE(p[7], p[12]) IT_CE(p[10], p[18])
    IT_CE(p[12], p[20]) IT_CE(p[10], p[20]) IT_CE(p[10], p[12])
    return p[12];
}

// ------------------------------------------------------------------------------------------------ CLAHE -> ... -> Canny map
constexpr int CT_H = 64, CT_W = 64;                                   // output tile; region = tile + 6 halo on each side
constexpr int RC_H = CT_H + 2 * HALO_C, RC_W = CT_W + 2 * HALO_C;    // 76 x 76
struct ClaheAxis { int pos, off1, off2; float a, a1; };               // per region column / row: source coordinate, LUT offsets of the two tiles, weights
__global__ void __launch_bounds__(IT_THREADS)
clahe_canny_kernel(const uint8_t* __restrict__ gray, const uint8_t* __restrict__ lut, ClaheGeom g, int H, int W, int low, int high,
                   uint8_t* __restrict__ clahe, uint8_t* __restrict__ pmap) {
    __shared__ uint8_t s_c[RC_H * RC_W];         // CLAHE, extended by reflect-101 (GaussianBlur's border)
    __shared__ uint8_t s_g[RC_H * RC_W];         // Gaussian, extended by replication (medianBlur's border)
    __shared__ uint8_t s_m[RC_H * RC_W];         // median, extended by replication (Canny's Sobel border)
    __shared__ uint16_t s_mag[RC_H * RC_W];      // horizontal Gaussian pass (Q8), later |dx| + |dy| (zero outside the image)
    __shared__ ClaheAxis s_ax[RC_W], s_ay[RC_H];
    const int t = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * CT_W - HALO_C, y0 = blockIdx.y * CT_H - HALO_C;
    const uint8_t* gp = gray + (size_t)b * H * W;
    const uint8_t* lp = lut + (size_t)b * g.tiles_x * g.tiles_y * 256;
    // CLAHE blend geometry is separable: one entry per region column and row (clahe.cpp computes the same tables per image)
    if (t < RC_W + RC_H) {
        const bool col = t < RC_W;
        const int i = col ? t : t - RC_W;
        const int n = col ? W : H, nt = col ? g.tiles_x : g.tiles_y;
        const int pos = clampi(reflect101((col ? x0 : y0) + i, n), n);          // far outside (never consumed): keep the address valid
        const float tf = __fsub_rn(__fmul_rn((float)pos, col ? g.inv_tw : g.inv_th), 0.5f);
        int t1 = (int)floorf(tf);
        const float a = __fsub_rn(tf, (float)t1);
        const int t2 = min(t1 + 1, nt - 1); t1 = max(t1, 0);
        const int stride = col ? 256 : g.tiles_x * 256;
        ClaheAxis e{pos, t1 * stride, t2 * stride, a, __fsub_rn(1.0f, a)};
        if (col) s_ax[i] = e; else s_ay[i] = e;
    }
    __syncthreads();
    for (int i = t; i < RC_H * RC_W; i += IT_THREADS) {
        const int v = i / RC_W, u = i - v * RC_W;
        const ClaheAxis ax = s_ax[u], ay = s_ay[v];
        const uint8_t* l = lp + gp[(size_t)ay.pos * W + ax.pos];
        const float l11 = (float)__ldg(l + ay.off1 + ax.off1), l12 = (float)__ldg(l + ay.off1 + ax.off2);
        const float l21 = (float)__ldg(l + ay.off2 + ax.off1), l22 = (float)__ldg(l + ay.off2 + ax.off2);
        const float top = __fadd_rn(__fmul_rn(l11, ax.a1), __fmul_rn(l12, ax.a));
        const float bot = __fadd_rn(__fmul_rn(l21, ax.a1), __fmul_rn(l22, ax.a));
        s_c[i] = sat_rint(__fadd_rn(__fmul_rn(top, ay.a1), __fmul_rn(bot, ay.a)));
    }
    __syncthreads();
    // interior of the CLAHE plane leaves here
    for (int i = t; i < CT_H * CT_W; i += IT_THREADS) {
        const int v = i / CT_W, u = i - v * CT_W, y = y0 + HALO_C + v, x = x0 + HALO_C + u;
        if (y < H && x < W) clahe[((size_t)b * H + y) * W + x] = s_c[(v + HALO_C) * RC_W + u + HALO_C];
    }
    // Gaussian 5x5 sigma 1.5 in fixed point: taps 31 60 74 60 31 (Q8) per axis, one rounding (x + 2^15) >> 16.
    // Columns / rows outside the image take the value at the clamped coordinate (what the median's replicate border sees).
    for (int i = t; i < RC_H * (RC_W - 4); i += IT_THREADS) {
        const int v = i / (RC_W - 4), u = i % (RC_W - 4) + 2;
        const uint8_t* r = s_c + v * RC_W + clampi(x0 + u, W) - x0;
        s_mag[v * RC_W + u] = (uint16_t)(31u * (r[-2] + r[2]) + 60u * (r[-1] + r[1]) + 74u * r[0]);
    }
    __syncthreads();
    for (int i = t; i < (RC_H - 4) * (RC_W - 4); i += IT_THREADS) {
        const int v = i / (RC_W - 4) + 2, u = i % (RC_W - 4) + 2;
        const uint16_t* r = s_mag + (clampi(y0 + v, H) - y0) * RC_W + u;
        const uint32_t acc = 31u * ((uint32_t)r[-2 * RC_W] + r[2 * RC_W]) + 60u * ((uint32_t)r[-RC_W] + r[RC_W]) + 74u * r[0];
        s_g[v * RC_W + u] = (uint8_t)min((acc + (1u << 15)) >> 16, 255u);
    }
    __syncthreads();
    // median 5x5, two horizontally adjacent pixels per thread (u16x2 lanes)
    constexpr int MPAIRS = (RC_W - 8) / 2;
    for (int i = t; i < (RC_H - 8) * MPAIRS; i += IT_THREADS) {
        const int v = i / MPAIRS + 4, u = (i % MPAIRS) * 2 + 4;
        const int ym = clampi(y0 + v, H) - y0;                                    // value Sobel's replicate border sees
        const int xa = clampi(x0 + u, W) - x0, xb = clampi(x0 + u + 1, W) - x0;
        uint32_t win[25];
#pragma unroll
        for (int k = 0; k < 25; k++) {
            const uint8_t* r = s_g + (ym + k / 5 - 2) * RC_W + k % 5 - 2;
            win[k] = (uint32_t)r[xa] | ((uint32_t)r[xb] << 16);
        }
        const uint32_t med = median25(win);
        s_m[v * RC_W + u] = (uint8_t)(med & 0xffu);
        s_m[v * RC_W + u + 1] = (uint8_t)(med >> 16);
    }
    __syncthreads();
    // Sobel magnitude |dx| + |dy| on the region shrunk by 5; zero outside the image
    for (int i = t; i < (RC_H - 10) * (RC_W - 10); i += IT_THREADS) {
        const int v = i / (RC_W - 10) + 5, u = i % (RC_W - 10) + 5;
        const int y = y0 + v, x = x0 + u;
        uint32_t mag = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const uint8_t* r = s_m + v * RC_W + u;
            const int a = r[-RC_W - 1], bb = r[-RC_W], c = r[-RC_W + 1], d = r[-1], f = r[1], gg = r[RC_W - 1], hh = r[RC_W], k = r[RC_W + 1];
            const int dx = (c + 2 * f + k) - (a + 2 * d + gg), dy = (gg + 2 * hh + k) - (a + 2 * bb + c);
            mag = (uint32_t)(abs(dx) + abs(dy));
        }
        s_mag[v * RC_W + u] = (uint16_t)mag;
    }
    __syncthreads();
    // non-maximum suppression with the integer tan(22.5 deg) sector test, thresholds: 2 strong, 0 candidate, 1 none
    for (int i = t; i < CT_H * CT_W; i += IT_THREADS) {
        const int v = i / CT_W + HALO_C, u = i % CT_W + HALO_C;
        const int y = y0 + v, x = x0 + u;
        if (y >= H || x >= W) continue;
        const uint16_t* mp = s_mag + v * RC_W + u;
        const int m = mp[0];
        uint8_t out = 1;
        if (m > low) {
            const uint8_t* r = s_m + v * RC_W + u;
            const int a = r[-RC_W - 1], bb = r[-RC_W], c = r[-RC_W + 1], d = r[-1], f = r[1], gg = r[RC_W - 1], hh = r[RC_W], k = r[RC_W + 1];
            const int dx = (c + 2 * f + k) - (a + 2 * d + gg), dy = (gg + 2 * hh + k) - (a + 2 * bb + c);
            const int ax = abs(dx), ay = abs(dy) << 15;
            const int tg22x = ax * 13573;
            bool keep;
            if (ay < tg22x) keep = m > mp[-1] && m >= mp[1];
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > mp[-RC_W] && m >= mp[RC_W];
                else { const int s = (dx ^ dy) < 0 ? -1 : 1; keep = m > mp[-RC_W - s] && m > mp[RC_W + s]; }
            }
            if (keep) out = m > high ? 2 : 0;
        }
        pmap[((size_t)b * H + y) * W + x] = out;
    }
}

// ------------------------------------------------------------------------------------------------ row bit masks
// Hysteresis and the morphology chain work on binary images: a warp owns a 64 x 32 window as one 64-bit mask per row
// (lane r = row r), so a 3x3 step over 2048 pixels is a handful of shifts, logic ops and two lane shuffles.
struct RowMasks { uint64_t strong, cand, valid; };      // map == 2, map == 0, inside the image
// Lane r reads its own row: 64 bytes as 17 aligned words re-aligned by funnel shifts; the map holds only 0 / 1 / 2, so "strong" is
// bit 1 of a byte and "candidate" is both low bits clear; four byte flags are gathered into a nibble by one multiply.
// (Words may start up to 3 bytes before / end up to 3 bytes after the window: still inside the scratch planes; masked by `valid`.)
__device__ __forceinline__ RowMasks load_rows(const uint8_t* __restrict__ gp, int H, int W, int x0, int y0, int lane) {
    RowMasks m{0, 0, 0};
    const int y = y0 + lane;
    const int lo = max(0, -x0), hi = min(64, W - x0);                      // window columns inside the image: [lo, hi)
    if (y < 0 || y >= H || hi <= lo) return m;
    m.valid = ((hi - lo) == 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull)) << lo;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(gp) + (size_t)y * W + x0;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(addr & 3) * 8u;
    uint32_t prev = __ldcg(wp);
    uint64_t strong = 0, cand = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t next = __ldcg(wp + j + 1);
        const uint32_t w = __funnelshift_r(prev, next, sh);               // window bytes 4j .. 4j+3
        prev = next;
        const uint32_t f2 = (w >> 1) & 0x01010101u, f0 = ~(w | (w >> 1)) & 0x01010101u;
        strong |= (uint64_t)((f2 * 0x01020408u) >> 24) << (4 * j);        // byte flags at bits 0, 8, 16, 24 -> bits 24..27
        cand |= (uint64_t)((f0 * 0x01020408u) >> 24) << (4 * j);
    }
    m.strong = strong & m.valid; m.cand = cand & m.valid;
    return m;
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int lane) { const uint64_t r = __shfl_up_sync(0xffffffffu, v, 1); return lane == 0 ? 0ull : r; }
__device__ __forceinline__ uint64_t shfl_dn64(uint64_t v, int lane) { const uint64_t r = __shfl_down_sync(0xffffffffu, v, 1); return lane == 31 ? 0ull : r; }

// ------------------------------------------------------------------------------------------------ hysteresis
// Persistent grid (cooperative launch: every CTA resident), one warp per 62 x 30 tile (+1 halo).  A round = every dirty tile
// floods its strong pixels into its candidates until nothing changes; a tile that changed publishes its new edge pixels
// and marks itself and its 8 neighbours dirty for the next round; rounds (grid barrier in between) repeat until none is
// dirty.  The fixed point is unique, so the result does not depend on the order.
struct HystCtl { unsigned int barrier; unsigned int flag[3]; };
constexpr int HY_IW = 62, HY_IH = 30;
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(IT_THREADS)
hysteresis_kernel(uint8_t* __restrict__ pmap, int B, int H, int W, int tiles_x, int tiles_y, uint8_t* __restrict__ dirty, HystCtl* ctl) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (IT_THREADS / 32) + (threadIdx.x >> 5), nw = gridDim.x * (IT_THREADS / 32);
    const int per_img = tiles_x * tiles_y, ntiles = B * per_img;
    const uint64_t interior = (lane >= 1 && lane <= HY_IH) ? 0x7ffffffffffffffeull : 0ull;
    for (unsigned int round = 0;; round++) {
        uint8_t* d_now = dirty + (size_t)(round & 1) * ntiles;
        uint8_t* d_next = dirty + (size_t)((round + 1) & 1) * ntiles;
        bool marked = false;
        for (int tile = gw; tile < ntiles; tile += nw) {
            if (round > 0) {
                if (!__ldcg(d_now + tile)) continue;
                __syncwarp();
                if (lane == 0) d_now[tile] = 0;
            }
            const int b = tile / per_img, r = tile - b * per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
            const int y0 = ty * HY_IH - 1, x0 = tx * HY_IW - 1;
            uint8_t* gp = pmap + (size_t)b * H * W;
            const RowMasks m = load_rows(gp, H, W, x0, y0, lane);
            const uint64_t C = m.cand & interior;
            uint64_t S = m.strong;
            if (!__any_sync(0xffffffffu, C != 0) || !__any_sync(0xffffffffu, S != 0)) continue;
            bool grew;
            do {
                const uint64_t n = S | (S << 1) | (S >> 1);
                const uint64_t add = C & ~S & (n | shfl_up64(n, lane) | shfl_dn64(n, lane));
                S |= add;
                grew = __any_sync(0xffffffffu, add != 0);
            } while (grew);
            const uint64_t fresh = S & ~m.strong;
            if (!__any_sync(0xffffffffu, fresh != 0)) continue;
            for (int rr = 1; rr <= HY_IH; rr++) {
                const uint64_t f = __shfl_sync(0xffffffffu, fresh, rr);
                if (f == 0) continue;
                uint8_t* row = gp + (size_t)(y0 + rr) * W + x0;
                if ((f >> lane) & 1ull) row[lane] = 2;
                if ((f >> (lane + 32)) & 1ull) row[lane + 32] = 2;
            }
            if (lane < 9) {                                                   // the 8 tiles around this one see new pixels in their halo
                const int ny = ty + lane / 3 - 1, nx = tx + lane % 3 - 1;
                if (ny >= 0 && ny < tiles_y && nx >= 0 && nx < tiles_x && lane != 4) d_next[b * per_img + ny * tiles_x + nx] = 1;
            }
            marked = true;
        }
        if (marked && lane == 0) atomicExch(&ctl->flag[round % 3], 1u);
        grid_barrier(&ctl->barrier, (round + 1) * gridDim.x);
        unsigned int f;
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(f) : "l"(&ctl->flag[round % 3]) : "memory");
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->flag[(round + 2) % 3] = 0;   // last read before this round's barrier, next written after the next one
        if (!f) break;
    }
}

// ------------------------------------------------------------------------------------------------ morphology + soft edges
// edges -> dilate -> erode -> dilate -> erode (3x3 box; pixels outside the image are ignored) on row masks, then the 3x3
// Gaussian (taps 1 2 1, reflect-101) of the 0 / 255 result and the per-image maximum.  One warp per 54 x 22 tile (+5 halo).
constexpr int MO_IW = 54, MO_IH = 22, MO_HALO = 5;
__device__ __forceinline__ uint64_t dilate_rows(uint64_t e, int lane) { const uint64_t h = e | (e << 1) | (e >> 1); return h | shfl_up64(h, lane) | shfl_dn64(h, lane); }
__device__ __forceinline__ uint64_t erode_rows(uint64_t e, int lane) {
    const uint64_t h = e & (e << 1) & (e >> 1);
    const uint64_t up = __shfl_up_sync(0xffffffffu, h, 1), dn = __shfl_down_sync(0xffffffffu, h, 1);
    return h & (lane == 0 ? 0ull : up) & (lane == 31 ? 0ull : dn);
}
__global__ void __launch_bounds__(IT_THREADS)
morph_kernel(const uint8_t* __restrict__ pmap, int H, int W, int tiles_x, int tiles_y, int ntiles, uint8_t* __restrict__ soft, unsigned int* __restrict__ maxv) {
    __shared__ uint64_t rows[IT_THREADS / 32][5][32];        // per warp: the 5 bit planes of the blurred rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * (IT_THREADS / 32) + warp;
    if (tile >= ntiles) return;
    const int per_img = tiles_x * tiles_y;
    const int b = tile / per_img, r = tile - b * per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
    const int y0 = ty * MO_IH - MO_HALO, x0 = tx * MO_IW - MO_HALO;
    const RowMasks m = load_rows(pmap + (size_t)b * H * W, H, W, x0, y0, lane);
    const uint64_t V = m.valid;
    uint64_t e = m.strong;                                   // edge pixels; 0 outside the image (dilate ignores it)
    e = dilate_rows(e, lane) | ~V;                           // outside = 255: ignored by the erode that follows
    e = erode_rows(e, lane) & V;
    e = dilate_rows(e, lane) | ~V;
    e = erode_rows(e, lane) & V;
    // 3x3 Gaussian (taps 1 2 1 per axis, reflect-101) of the 0 / 255 result as bit-sliced arithmetic on the row masks:
    // per row h = left + 2 centre + right (3 bit planes), then v = h_up + 2 h + h_down (5 bit planes, 0..16), 64 columns at a time.
    const int y = y0 + lane;
    uint64_t L = e << 1, R = e >> 1;                         // left / right neighbour of every column
    if (x0 <= 0 && -x0 < 64) { const int c = -x0; L = (L & ~(1ull << c)) | (((e >> (c + 1)) & 1ull) << c); }                // x = 0: its left neighbour is x = 1
    if (W - 1 - x0 >= 1 && W - 1 - x0 < 64) { const int c = W - 1 - x0; R = (R & ~(1ull << c)) | (((e >> (c - 1)) & 1ull) << c); }   // x = W-1: its right neighbour is x = W-2
    const uint64_t k = L & R, h0 = L ^ R, h1 = k ^ e, h2 = k & e;
    uint64_t u0 = __shfl_up_sync(0xffffffffu, h0, 1), u1 = __shfl_up_sync(0xffffffffu, h1, 1), u2 = __shfl_up_sync(0xffffffffu, h2, 1);
    uint64_t d0 = __shfl_down_sync(0xffffffffu, h0, 1), d1 = __shfl_down_sync(0xffffffffu, h1, 1), d2 = __shfl_down_sync(0xffffffffu, h2, 1);
    if (y == 0) { u0 = d0; u1 = d1; u2 = d2; }               // reflect-101: the row above row 0 is row 1
    if (y == H - 1) { d0 = u0; d1 = u1; d2 = u2; }           // the row below row H-1 is row H-2
    const uint64_t t0 = u0 ^ d0, c0 = u0 & d0;
    const uint64_t t1 = u1 ^ d1 ^ c0, c1 = (u1 & d1) | (c0 & (u1 ^ d1));
    const uint64_t t2 = u2 ^ d2 ^ c1, t3 = (u2 & d2) | (c1 & (u2 ^ d2));
    const uint64_t v1 = t1 ^ h0, e1 = t1 & h0;
    const uint64_t v2 = t2 ^ h1 ^ e1, e2 = (t2 & h1) | (e1 & (t2 ^ h1));
    const uint64_t v3 = t3 ^ h2 ^ e2, v4 = (t3 & h2) | (e2 & (t3 ^ h2));
    uint64_t* pr = &rows[warp][0][0];
    pr[0 * 32 + lane] = t0; pr[1 * 32 + lane] = v1; pr[2 * 32 + lane] = v2; pr[3 * 32 + lane] = v3; pr[4 * 32 + lane] = v4;
    __syncwarp();
    uint32_t mx = 0;
    uint8_t* sp = soft + (size_t)b * H * W;
    for (int rr = MO_HALO; rr < MO_HALO + MO_IH; rr++) {
        const int yr = y0 + rr;
        if (yr >= H) break;
        const uint64_t p0 = pr[rr], p1 = pr[32 + rr], p2 = pr[64 + rr], p3 = pr[96 + rr], p4 = pr[128 + rr];
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            const int c = lane + 32 * kk, x = x0 + c;
            if (c < MO_HALO || c >= MO_HALO + MO_IW || x >= W) continue;
            const uint32_t acc = (uint32_t)((p0 >> c) & 1) | ((uint32_t)((p1 >> c) & 1) << 1) | ((uint32_t)((p2 >> c) & 1) << 2) |
                                 ((uint32_t)((p3 >> c) & 1) << 3) | ((uint32_t)((p4 >> c) & 1) << 4);
            const uint32_t v = (255u * acc + 8u) >> 4;                     // Q8 taps 64 128 64 twice: (acc * 4096 + 2^15) >> 16
            sp[(size_t)yr * W + x] = (uint8_t)v;
            mx = max(mx, v);
        }
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(maxv + b, mx);
}

// rgb pipeline: CLAHE blend + Gaussian 3x3 sigma 0.5 (taps 27 202 27) per colour plane
constexpr int RR_H = TH + 2, RR_W = TW + 2;
__global__ void __launch_bounds__(IT_THREADS)
clahe_gauss3_kernel(const uint8_t* __restrict__ plane, const uint8_t* __restrict__ lut, ClaheGeom g, int H, int W, uint8_t* __restrict__ out) {
    __shared__ uint8_t s_c[RR_H * RR_W];
    const int t = threadIdx.x, pl = blockIdx.z;
    const int x0 = blockIdx.x * TW - 1, y0 = blockIdx.y * TH - 1;
    const uint8_t* gp = plane + (size_t)pl * H * W;
    const uint8_t* lp = lut + (size_t)pl * g.tiles_x * g.tiles_y * 256;
    for (int i = t; i < RR_H * RR_W; i += IT_THREADS) {
        const int v = i / RR_W, u = i - v * RR_W;
        const int yc = clampi(reflect101(y0 + v, H), H), xc = clampi(reflect101(x0 + u, W), W);
        s_c[i] = clahe_px(g, lp, xc, yc, gp[(size_t)yc * W + xc]);
    }
    __syncthreads();
    for (int i = t; i < TH * TW; i += IT_THREADS) {
        const int v = i / TW + 1, u = i % TW + 1, y = y0 + v, x = x0 + u;
        if (y >= H || x >= W) continue;
        const uint8_t* q = s_c + v * RR_W + u;
        uint32_t acc = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++) acc += (27u * (q[dy * RR_W - 1] + q[dy * RR_W + 1]) + 202u * q[dy * RR_W]) * (dy == 0 ? 202u : 27u);
        out[((size_t)pl * H + y) * W + x] = (uint8_t)min((acc + (1u << 15)) >> 16, 255u);
    }
}

// ------------------------------------------------------------------------------------------------ resize
// Pillow precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the whole axis: float64, C operation order.
__global__ void resize_coeff_kernel(int in_size, int out_size, int ksize, int32_t* __restrict__ bounds, int32_t* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double fscale = scale < 1.0 ? 1.0 : scale;
    const double support = fscale;                               // bilinear support 1.0 * filterscale
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    const double ss = __ddiv_rn(1.0, fscale);
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5)); if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5)); if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; x++) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        ww = __dadd_rn(ww, a < 1.0 ? __dsub_rn(1.0, a) : 0.0);
    }
    for (int x = 0; x < ksize; x++) {
        int32_t q = 0;
        if (x < xmax) {
            double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            q = w < 0.0 ? __double2int_rz(__dadd_rn(-0.5, __dmul_rn(w, 4194304.0))) : __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, 4194304.0)));
        }
        kk[(size_t)xx * ksize + x] = q;
    }
    bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
}

// horizontal pass; grayscale: the source row is blended on the fly = addWeighted(clahe, 0.7, rescale(soft), 0.3)
constexpr int HR_ROWS = 4;
static_assert(HR_ROWS * 64 == IT_THREADS, "blend_hresize_kernel: 64 threads per staged row");
__global__ void __launch_bounds__(IT_THREADS)
blend_hresize_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ soft, const unsigned int* __restrict__ maxv, int planes_per_image,
                     int H, int W, int S, int ksize, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, uint8_t* __restrict__ hp) {
    extern __shared__ uint8_t sm_dyn[];
    uint8_t* row = sm_dyn;                                        // [HR_ROWS][W]
    __shared__ uint8_t rescale[256];
    const int t = threadIdx.x, pl = blockIdx.y;
    const int y0 = blockIdx.x * HR_ROWS;
    const uint8_t* ap = a + (size_t)pl * H * W;
    const int rr = t >> 6, tx = t & 63, y = y0 + rr;                // 64 threads per staged row (HR_ROWS = 4)
    if (soft) {
        // (e.astype(float) / e.max() * 255).astype(uint8) (float64, truncation); unchanged when the maximum is 0
        const unsigned int mx = maxv[pl / planes_per_image];
        rescale[t] = mx ? (uint8_t)__double2int_rz(__dmul_rn(__ddiv_rn((double)t, (double)mx), 255.0)) : (uint8_t)t;
        __syncthreads();
        if (y < H) {
            const uint8_t* sp = soft + ((size_t)pl * H + y) * W;
            const uint8_t* cp = ap + (size_t)y * W;
            for (int x = tx; x < W; x += 64) {
                const float e = (float)rescale[sp[x]];
                // cv2.addWeighted: fma(a, 0.7f, round_f32(e * 0.3f)), round half to even, saturate
                row[rr * W + x] = sat_rint(__fmaf_rn((float)cp[x], 0.7f, __fmul_rn(e, 0.3f)));
            }
        }
    } else if (y < H) {
        const uint8_t* cp = ap + (size_t)y * W;
        for (int x = tx; x < W; x += 64) row[rr * W + x] = cp[x];
    }
    __syncthreads();
    if (y < H) {
        for (int xx = tx; xx < S; xx += 64) {
            const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
            const int32_t* k = kk + (size_t)xx * ksize;
            int acc = 1 << 21;
            for (int j = 0; j < n; j++) acc += (int)row[rr * W + xmin + j] * k[j];
            acc >>= 22;
            hp[((size_t)pl * H + y) * S + xx] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
        }
    }
}

// vertical pass + ToTensor (x / 255) + Normalize ((x - mean) / std), float32
__global__ void __launch_bounds__(IT_THREADS)
vresize_norm_kernel(const uint8_t* __restrict__ hp, int planes_per_image, int H, int S, int ksize, const int32_t* __restrict__ bounds,
                    const int32_t* __restrict__ kk, float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
    const int pl = blockIdx.y, c = pl % planes_per_image;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    for (int i = blockIdx.x * IT_THREADS + threadIdx.x; i < S * S; i += gridDim.x * IT_THREADS) {
        const int yy = i / S, xx = i - yy * S;
        const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
        const int32_t* k = kk + (size_t)yy * ksize;
        const uint8_t* p = hp + ((size_t)pl * H + ymin) * S + xx;
        int acc = 1 << 21;
        for (int j = 0; j < n; j++) acc += (int)p[(size_t)j * S] * k[j];
        acc >>= 22;
        acc = acc < 0 ? 0 : (acc > 255 ? 255 : acc);
        out[(size_t)pl * S * S + i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)acc, 255.0f), mean), sd);
    }
}

struct ITLayout { size_t plane, lut, clahe, pmap, soft, hpass, maxv, ctl, dirty, zero_end, xb, xk, yb, yk, total; int ksx, ksy, htx, hty; };
int resize_ksize(int in_size, int out_size) {
    const double scale = (double)in_size / out_size, fscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(fscale) * 2 + 1;
}
ITLayout it_layout(int B, int H, int W, int planes, int tiles, int S) {
    ITLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t img = (size_t)B * planes * H * W;
    L.plane = take(img); L.lut = take((size_t)B * planes * tiles * 256); L.clahe = take(img); L.pmap = take(img); L.soft = take(img);
    L.hpass = take((size_t)B * planes * H * S);
    L.maxv = take((size_t)B * 4); L.ctl = take(sizeof(HystCtl));
    L.htx = (W + HY_IW - 1) / HY_IW; L.hty = (H + HY_IH - 1) / HY_IH;
    L.dirty = take((size_t)2 * B * L.htx * L.hty);
    L.zero_end = off;                                          // [maxv, zero_end) is cleared before every run
    L.ksx = resize_ksize(W, S); L.ksy = resize_ksize(H, S);
    L.xb = take((size_t)S * 8); L.xk = take((size_t)S * L.ksx * 4); L.yb = take((size_t)S * 8); L.yk = take((size_t)S * L.ksy * 4);
    L.total = off;
    return L;
}

}  // namespace
}  // namespace kpd

using namespace kpd;

extern "C" size_t kpd_itransform_scratch_bytes(int batch, int height, int width, int grayscale, int tiles_x, int tiles_y, int out_size) {
    if (batch <= 0 || height <= 0 || width <= 0 || tiles_x <= 0 || tiles_y <= 0 || out_size <= 0) return 0;
    return it_layout(batch, height, width, grayscale ? 1 : 3, tiles_x * tiles_y, out_size).total;
}

extern "C" int kpd_itransform_stage_offsets(int batch, int height, int width, int grayscale, int tiles_x, int tiles_y, int out_size,
                                            size_t* offsets /*[6]: plane, lut, clahe, pmap, soft, hpass*/) {
    KPD_CHECK_ARG(offsets && batch > 0 && height > 0 && width > 0 && tiles_x > 0 && tiles_y > 0 && out_size > 0, "kpd_itransform_stage_offsets: bad arguments");
    const ITLayout L = it_layout(batch, height, width, grayscale ? 1 : 3, tiles_x * tiles_y, out_size);
    offsets[0] = L.plane; offsets[1] = L.lut; offsets[2] = L.clahe; offsets[3] = L.pmap; offsets[4] = L.soft; offsets[5] = L.hpass;
    return 0;
}

extern "C" int kpd_itransform(const uint8_t* d_src, int batch, int height, int width, int channels, int grayscale, double clip_limit,
                              int tiles_x, int tiles_y, int out_size, float* d_out, void* d_scratch, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    KPD_CHECK_ARG(d_src && d_out && d_scratch, "kpd_itransform: null pointer");
    KPD_CHECK_ARG(batch > 0 && batch <= 65535, "kpd_itransform: batch %d out of range", batch);
    KPD_CHECK_ARG(channels == 1 || channels == 3 || channels == 4, "kpd_itransform: channels must be 1, 3 or 4 (got %d)", channels);
    KPD_CHECK_ARG(height >= 16 && width >= 16 && height <= 16384 && width <= 16384, "kpd_itransform: image %dx%d outside 16..16384", height, width);
    KPD_CHECK_ARG(tiles_x >= 1 && tiles_y >= 1 && tiles_x * tiles_y <= 4096, "kpd_itransform: bad CLAHE tile grid %dx%d", tiles_x, tiles_y);
    KPD_CHECK_ARG(out_size >= 1 && out_size <= 4096, "kpd_itransform: output size %d out of range", out_size);
    const int planes = grayscale ? 1 : 3, ntile = tiles_x * tiles_y, S = out_size, H = height, W = width;
    KPD_CHECK_ARG(batch * planes <= 65535, "kpd_itransform: batch x planes exceeds 65535");
    KPD_CHECK_ARG((int64_t)batch * planes * H * W < ((int64_t)1 << 31) && (int64_t)batch * planes * ntile < ((int64_t)1 << 30),
                  "kpd_itransform: %d x %dx%d pixels per call exceed the 2^31 limit of the tile indices (split the batch)", batch, H, W);
    const ITLayout L = it_layout(batch, H, W, planes, ntile, S);
    char* ws = (char*)d_scratch;
    ITParams p{};
    p.src = d_src; p.B = batch; p.H = H; p.W = W; p.C = channels; p.planes_per_image = planes; p.to_gray = grayscale ? 1 : 0;
    p.tiles_x = tiles_x; p.tiles_y = tiles_y;
    // clahe.cpp: the image is padded (reflect-101) on the bottom / right unless BOTH sides divide; note the full extra tile when one side divides
    int eh = H, ew = W;
    if (!(W % tiles_x == 0 && H % tiles_y == 0)) { eh = H + (tiles_y - H % tiles_y); ew = W + (tiles_x - W % tiles_x); }
    p.th = eh / tiles_y; p.tw = ew / tiles_x;
    KPD_CHECK_ARG(eh - H < H && ew - W < W, "kpd_itransform: image %dx%d too small for a %dx%d CLAHE grid", H, W, tiles_x, tiles_y);
    const int area = p.tw * p.th;
    p.clip = 0;
    if (clip_limit > 0.0) { p.clip = (int)(clip_limit * area / 256); if (p.clip < 1) p.clip = 1; }
    p.lut_scale = 255.0f / (float)area;
    uint8_t* plane = (uint8_t*)(ws + L.plane); uint8_t* lut = (uint8_t*)(ws + L.lut); uint8_t* clahe = (uint8_t*)(ws + L.clahe);
    uint8_t* pmap = (uint8_t*)(ws + L.pmap); uint8_t* soft = (uint8_t*)(ws + L.soft); uint8_t* hp = (uint8_t*)(ws + L.hpass);
    unsigned int* maxv = (unsigned int*)(ws + L.maxv); HystCtl* ctl = (HystCtl*)(ws + L.ctl); uint8_t* dirty = (uint8_t*)(ws + L.dirty);
    int32_t* xb = (int32_t*)(ws + L.xb); int32_t* xk = (int32_t*)(ws + L.xk); int32_t* yb = (int32_t*)(ws + L.yb); int32_t* yk = (int32_t*)(ws + L.yk);
    ProfScope ps("itransform", s);

    resize_coeff_kernel<<<(S + 127) / 128, 128, 0, s>>>(W, S, L.ksx, xb, xk);
    KPD_LAUNCH_CHECK();
    resize_coeff_kernel<<<(S + 127) / 128, 128, 0, s>>>(H, S, L.ksy, yb, yk);
    KPD_LAUNCH_CHECK();
    { ProfScope pp("itransform.lut", s);
    clahe_lut_kernel<<<batch * planes * ntile, IT_THREADS, 0, s>>>(p, plane, lut);
    KPD_LAUNCH_CHECK(); }
    ClaheGeom g{tiles_x, tiles_y, 1.0f / (float)p.tw, 1.0f / (float)p.th};
    const dim3 tgrid((W + TW - 1) / TW, (H + TH - 1) / TH, batch * planes);
    const dim3 cgrid((W + CT_W - 1) / CT_W, (H + CT_H - 1) / CT_H, batch * planes);
    const uint8_t* resize_src; const uint8_t* soft_arg = nullptr;
    if (grayscale) {
        { ProfScope pp("itransform.clahe_canny", s);
        clahe_canny_kernel<<<cgrid, IT_THREADS, 0, s>>>(plane, lut, g, H, W, 100, 200, clahe, pmap);
        KPD_LAUNCH_CHECK(); }
        KPD_CUDA(cudaMemsetAsync(maxv, 0, L.zero_end - L.maxv, s));                    // per-image maxima, hysteresis control block, dirty flags
        int dev = 0, sms = 0, per_sm = 0;
        KPD_CUDA(cudaGetDevice(&dev));
        KPD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        KPD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hysteresis_kernel, IT_THREADS, 0));
        const int ntiles = batch * L.htx * L.hty;
        const int hgrid = std::max(1, std::min((ntiles + IT_THREADS / 32 - 1) / (IT_THREADS / 32), sms * std::min(per_sm, 4)));
        int Bv = batch, Hv = H, Wv = W, htx = L.htx, hty = L.hty;
        void* args[] = {&pmap, &Bv, &Hv, &Wv, &htx, &hty, &dirty, &ctl};
        { ProfScope pp("itransform.hysteresis", s);
        KPD_CUDA(cudaLaunchCooperativeKernel((const void*)hysteresis_kernel, dim3(hgrid), dim3(IT_THREADS), args, 0, s));
        KPD_LAUNCH_CHECK(); }
        const int mtx = (W + MO_IW - 1) / MO_IW, mty = (H + MO_IH - 1) / MO_IH, mtiles = batch * mtx * mty;
        { ProfScope pp("itransform.morph", s);
        morph_kernel<<<(mtiles + IT_THREADS / 32 - 1) / (IT_THREADS / 32), IT_THREADS, 0, s>>>(pmap, H, W, mtx, mty, mtiles, soft, maxv);
        KPD_LAUNCH_CHECK(); }
        resize_src = clahe; soft_arg = soft;
    } else {
        { ProfScope pp("itransform.clahe_gauss3", s);
        clahe_gauss3_kernel<<<tgrid, IT_THREADS, 0, s>>>(plane, lut, g, H, W, clahe);
        KPD_LAUNCH_CHECK(); }
        resize_src = clahe;
    }
    const size_t hsm = (size_t)HR_ROWS * W;
    KPD_CHECK_ARG(hsm <= 48 * 1024 - 512, "kpd_itransform: width %d too large for the row staging", W);
    { ProfScope pp("itransform.hresize", s);
    blend_hresize_kernel<<<dim3((H + HR_ROWS - 1) / HR_ROWS, batch * planes), IT_THREADS, hsm, s>>>(resize_src, soft_arg, maxv, planes, H, W, S, L.ksx, xb, xk, hp);
    KPD_LAUNCH_CHECK(); }
    const float m[3] = {grayscale ? 0.5f : 0.485f, 0.456f, 0.406f}, sd[3] = {grayscale ? 0.5f : 0.229f, 0.224f, 0.225f};
    { ProfScope pp("itransform.vresize", s);
    vresize_norm_kernel<<<dim3(std::max(1, std::min(64, (S * S + IT_THREADS - 1) / IT_THREADS)), batch * planes), IT_THREADS, 0, s>>>(
        hp, planes, H, S, L.ksy, yb, yk, m[0], m[1], m[2], sd[0], sd[1], sd[2], d_out);
    KPD_LAUNCH_CHECK(); }
    return 0;
}
