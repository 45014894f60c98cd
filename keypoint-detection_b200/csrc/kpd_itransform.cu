// ITransform on the GPU: the reference's per-image OpenCV / Pillow preprocessing (dll/data/transforms.py:9-114)
// as batched byte kernels, bit-exact with the libraries (oracle/itransform.py states every stage and is pinned on them).
//
//   grayscale:  gray -> CLAHE -> { Gaussian 5x5 s1.5 -> median 5x5 -> Canny(100,200) -> dilate,erode,dilate,erode 3x3
//               -> Gaussian 3x3 -> rescale to the maximum } -> 0.7 CLAHE + 0.3 edges -> antialiased bilinear resize -> normalise
//   rgb:        per-channel CLAHE -> Gaussian 3x3 s0.5 -> resize -> ImageNet normalise
//
// All stages are HBM / L2-bound integer work on 1-byte pixels, so the design is about touching every plane once:
//   clahe_lut_kernel      one CTA per (plane, CLAHE tile): gray conversion + shared-memory histogram + clip / redistribute / scan -> LUT
//   clahe_canny_kernel    one CTA per 32x64 tile (+6 halo): CLAHE blend -> Gaussian -> median -> Sobel -> NMS, stage to stage in
//                         shared memory; writes the CLAHE plane and the Canny map (0 candidate / 1 none / 2 strong)
//   hysteresis_kernel     persistent cooperative grid: tile-local flood fill in shared memory, rounds separated by a grid barrier
//   morph_kernel          edges -> dilate -> erode -> dilate -> erode -> Gaussian 3x3 in one tile pass (+5 halo) + per-image maximum
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
constexpr int TH = 32, TW = 64;                 // output tile of the fused stencil kernels
constexpr int HALO_C = 6;                       // gauss 2 + median 2 + sobel 1 + nms 1
constexpr int HALO_M = 5;                       // 4 morphology steps + gauss 1

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
    const int area = p.tw * p.th;
    uint8_t* dst = plane + (size_t)pl * p.H * p.W;
    for (int i = t; i < area; i += IT_THREADS) {
        const int yy = i / p.tw, xx = i - yy * p.tw;
        const int ye = ty * p.th + yy, xe = tx * p.tw + xx;            // coordinates in the padded image
        const int y = reflect101(ye, p.H), x = reflect101(xe, p.W);    // copyMakeBorder(BORDER_REFLECT_101) on the bottom / right
        const uint32_t v = plane_px(p, b, c, y, x);
        if (ye < p.H && xe < p.W) dst[(size_t)y * p.W + x] = (uint8_t)v;
        atomicAdd(&hist[warp][v], 1u);
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
// Forgetful selection: of 14 values neither the minimum nor the maximum can be the 13th of 25; drop both, take the next value in.
#define IT_CE(a, b) { const uint32_t _lo = min(a, b); b = max(a, b); a = _lo; }
template <int N> __device__ __forceinline__ void ends_minmax(uint32_t (&v)[14]) {      // min -> v[0], max -> v[N-1]
#pragma unroll
    for (int i = 0; i < N / 2; i++) IT_CE(v[i], v[N - 1 - i]);
#pragma unroll
    for (int i = 1; i < (N + 1) / 2; i++) IT_CE(v[0], v[i]);
#pragma unroll
    for (int i = N / 2; i < N - 1; i++) IT_CE(v[i], v[N - 1]);
}
template <int N> struct Forget {
    static __device__ __forceinline__ uint32_t run(uint32_t (&v)[14], const uint32_t (&rest)[11]) {
        ends_minmax<N>(v);
        v[0] = rest[14 - N];                       // the maximum at v[N-1] falls off the end, the minimum is overwritten
        return Forget<N - 1>::run(v, rest);
    }
};
template <> struct Forget<3> {
    static __device__ __forceinline__ uint32_t run(uint32_t (&v)[14], const uint32_t (&)[11]) {
        IT_CE(v[0], v[1]); IT_CE(v[1], v[2]); IT_CE(v[0], v[1]);
        return v[1];
    }
};

// ------------------------------------------------------------------------------------------------ CLAHE -> ... -> Canny map
constexpr int RC_H = TH + 2 * HALO_C, RC_W = TW + 2 * HALO_C;        // 44 x 76 region
__global__ void __launch_bounds__(IT_THREADS)
clahe_canny_kernel(const uint8_t* __restrict__ gray, const uint8_t* __restrict__ lut, ClaheGeom g, int H, int W, int low, int high,
                   uint8_t* __restrict__ clahe, uint8_t* __restrict__ pmap) {
    __shared__ uint8_t s_c[RC_H * RC_W];         // CLAHE, extended by reflect-101 (GaussianBlur's border)
    __shared__ uint8_t s_g[RC_H * RC_W];         // Gaussian, extended by replication (medianBlur's border)
    __shared__ uint8_t s_m[RC_H * RC_W];         // median, extended by replication (Canny's Sobel border)
    __shared__ uint16_t s_mag[RC_H * RC_W];      // |dx| + |dy|, zero outside the image
    const int t = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * TW - HALO_C, y0 = blockIdx.y * TH - HALO_C;
    const uint8_t* gp = gray + (size_t)b * H * W;
    const uint8_t* lp = lut + (size_t)b * g.tiles_x * g.tiles_y * 256;
    for (int i = t; i < RC_H * RC_W; i += IT_THREADS) {
        const int v = i / RC_W, u = i - v * RC_W;
        const int y = reflect101(y0 + v, H), x = reflect101(x0 + u, W);
        const int yc = clampi(y, H), xc = clampi(x, W);                 // far outside (never consumed): keep the address valid
        s_c[i] = clahe_px(g, lp, xc, yc, gp[(size_t)yc * W + xc]);
    }
    __syncthreads();
    // interior of the CLAHE plane leaves here
    for (int i = t; i < TH * TW; i += IT_THREADS) {
        const int v = i / TW, u = i - v * TW, y = y0 + HALO_C + v, x = x0 + HALO_C + u;
        if (y < H && x < W) clahe[((size_t)b * H + y) * W + x] = s_c[(v + HALO_C) * RC_W + u + HALO_C];
    }
    // Gaussian 5x5 sigma 1.5 in fixed point: taps 31 60 74 60 31 (Q8), one rounding (x + 2^15) >> 16
    for (int i = t; i < (RC_H - 4) * (RC_W - 4); i += IT_THREADS) {
        const int v = i / (RC_W - 4) + 2, u = i % (RC_W - 4) + 2;
        const int ym = clampi(y0 + v, H) - y0, xm = clampi(x0 + u, W) - x0;       // value the median's replicate border sees
        uint32_t acc = 0;
#pragma unroll
        for (int dy = -2; dy <= 2; dy++) {
            const uint8_t* r = s_c + (ym + dy) * RC_W + xm;
            const uint32_t hrow = 31u * (r[-2] + r[2]) + 60u * (r[-1] + r[1]) + 74u * r[0];
            acc += hrow * (dy == 0 ? 74u : ((dy == -1 || dy == 1) ? 60u : 31u));
        }
        s_g[v * RC_W + u] = (uint8_t)min((acc + (1u << 15)) >> 16, 255u);
    }
    __syncthreads();
    // median 5x5
    for (int i = t; i < (RC_H - 8) * (RC_W - 8); i += IT_THREADS) {
        const int v = i / (RC_W - 8) + 4, u = i % (RC_W - 8) + 4;
        const int ym = clampi(y0 + v, H) - y0, xm = clampi(x0 + u, W) - x0;       // value Sobel's replicate border sees
        uint32_t w14[14], rest[11];
#pragma unroll
        for (int k = 0; k < 25; k++) {
            const uint32_t px = s_g[(ym + k / 5 - 2) * RC_W + xm + k % 5 - 2];
            if (k < 14) w14[k] = px; else rest[k - 14] = px;
        }
        s_m[v * RC_W + u] = (uint8_t)Forget<14>::run(w14, rest);
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
    for (int i = t; i < TH * TW; i += IT_THREADS) {
        const int v = i / TW + HALO_C, u = i % TW + HALO_C;
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

// ------------------------------------------------------------------------------------------------ hysteresis
// Persistent grid (cooperative launch: every CTA resident).  A round = every tile floods its strong pixels into its
// candidates in shared memory until nothing changes; tiles whose pixels changed publish them and raise the round's flag;
// rounds repeat (grid barrier in between) until a round changes nothing.  The fixed point is unique, so the result
// does not depend on the sweep order.
struct HystCtl { unsigned int barrier; unsigned int flag[3]; };
constexpr int HT_H = 32, HT_W = 64, HS_W = HT_W + 2 + 2;             // padded row (68) keeps 4-byte alignment irrelevant: byte accesses only
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
hysteresis_kernel(uint8_t* __restrict__ pmap, int B, int H, int W, int tiles_x, int tiles_y, uint8_t* __restrict__ tile_live, HystCtl* ctl) {
    __shared__ uint8_t s[(HT_H + 2) * HS_W];
    const int t = threadIdx.x;
    const int ntiles = B * tiles_x * tiles_y;
    for (unsigned int round = 0;; round++) {
        int cta_changed = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (round > 0 && !tile_live[tile]) continue;                       // no candidate left in this tile
            const int b = tile / (tiles_x * tiles_y), r = tile - b * tiles_x * tiles_y;
            const int y0 = (r / tiles_x) * HT_H - 1, x0 = (r % tiles_x) * HT_W - 1;
            uint8_t* gp = pmap + (size_t)b * H * W;
            int has_strong = 0, has_cand = 0;
            for (int i = t; i < (HT_H + 2) * (HT_W + 2); i += IT_THREADS) {
                const int v = i / (HT_W + 2), u = i - v * (HT_W + 2);
                const int y = y0 + v, x = x0 + u;
                uint8_t px = 1;
                if (y >= 0 && y < H && x >= 0 && x < W) px = gp[(size_t)y * W + x];
                s[v * HS_W + u] = px;
                has_strong |= (px == 2);
                has_cand |= (px == 0 && v >= 1 && v <= HT_H && u >= 1 && u <= HT_W);
            }
            has_strong = __syncthreads_or(has_strong);
            has_cand = __syncthreads_or(has_cand);
            if (t == 0) tile_live[tile] = (uint8_t)(has_cand != 0);
            int tile_changed = 0;
            if (has_strong && has_cand) {
                int c;
                do {
                    c = 0;
                    for (int i = t; i < HT_H * HT_W; i += IT_THREADS) {
                        const int v = i / HT_W + 1, u = i % HT_W + 1;
                        uint8_t* q = s + v * HS_W + u;
                        if (*q == 0) {
                            const bool hit = q[-HS_W - 1] == 2 || q[-HS_W] == 2 || q[-HS_W + 1] == 2 || q[-1] == 2 || q[1] == 2 ||
                                             q[HS_W - 1] == 2 || q[HS_W] == 2 || q[HS_W + 1] == 2;
                            if (hit) { *q = 2; c = 1; }
                        }
                    }
                    c = __syncthreads_or(c);
                    tile_changed |= c;
                } while (c);
            }
            if (tile_changed) {
                for (int i = t; i < HT_H * HT_W; i += IT_THREADS) {
                    const int v = i / HT_W + 1, u = i % HT_W + 1;
                    const int y = y0 + v, x = x0 + u;
                    if (y < H && x < W && s[v * HS_W + u] == 2) gp[(size_t)y * W + x] = 2;
                }
                cta_changed = 1;
            }
            __syncthreads();
        }
        if (cta_changed && t == 0) atomicExch(&ctl->flag[round % 3], 1u);
        grid_barrier(&ctl->barrier, (round + 1) * gridDim.x);
        unsigned int f;
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(f) : "l"(&ctl->flag[round % 3]) : "memory");
        if (blockIdx.x == 0 && t == 0) ctl->flag[(round + 2) % 3] = 0;        // last read before this round's barrier, next written after the next one
        if (!f) break;
    }
}

// ------------------------------------------------------------------------------------------------ morphology + soft edges
constexpr int RM_H = TH + 2 * HALO_M, RM_W = TW + 2 * HALO_M;        // 42 x 74
template <bool DILATE> __device__ __forceinline__ uint8_t box3(const uint8_t* q, int stride) {
    uint32_t r = q[0];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) r = DILATE ? max(r, (uint32_t)q[dy * stride + dx]) : min(r, (uint32_t)q[dy * stride + dx]);
    return (uint8_t)r;
}
__global__ void __launch_bounds__(IT_THREADS)
morph_kernel(const uint8_t* __restrict__ pmap, int H, int W, uint8_t* __restrict__ soft, unsigned int* __restrict__ maxv) {
    __shared__ uint8_t sa[RM_H * RM_W], sb[RM_H * RM_W];
    __shared__ unsigned int smax;
    const int t = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * TW - HALO_M, y0 = blockIdx.y * TH - HALO_M;
    const uint8_t* gp = pmap + (size_t)b * H * W;
    if (t == 0) smax = 0;
    // edges = 255 where the map says 2; outside the image 0 (dilate ignores it)
    for (int i = t; i < RM_H * RM_W; i += IT_THREADS) {
        const int v = i / RM_W, u = i - v * RM_W, y = y0 + v, x = x0 + u;
        sa[i] = (y >= 0 && y < H && x >= 0 && x < W && gp[(size_t)y * W + x] == 2) ? 255 : 0;
    }
    __syncthreads();
    // four 3x3 steps, ping-pong; the value outside the image is the neutral element of the NEXT step (erode: 255, dilate: 0)
    auto step = [&](const uint8_t* in, uint8_t* out, int shrink, bool dilate, uint8_t outside) {
        const int hh = RM_H - 2 * shrink, ww = RM_W - 2 * shrink;
        for (int i = t; i < hh * ww; i += IT_THREADS) {
            const int v = i / ww + shrink, u = i % ww + shrink, y = y0 + v, x = x0 + u;
            uint8_t r = outside;
            if (y >= 0 && y < H && x >= 0 && x < W) r = dilate ? box3<true>(in + v * RM_W + u, RM_W) : box3<false>(in + v * RM_W + u, RM_W);
            out[v * RM_W + u] = r;
        }
        __syncthreads();
    };
    step(sa, sb, 1, true, 255);
    step(sb, sa, 2, false, 0);
    step(sa, sb, 3, true, 255);
    // last erode: its consumer is GaussianBlur (reflect-101) -> compute at the mirrored coordinate
    {
        const int hh = RM_H - 8, ww = RM_W - 8;
        for (int i = t; i < hh * ww; i += IT_THREADS) {
            const int v = i / ww + 4, u = i % ww + 4;
            const int ym = clampi(reflect101(y0 + v, H), H) - y0, xm = clampi(reflect101(x0 + u, W), W) - x0;
            sa[v * RM_W + u] = box3<false>(sb + ym * RM_W + xm, RM_W);
        }
        __syncthreads();
    }
    unsigned int mx = 0;
    for (int i = t; i < TH * TW; i += IT_THREADS) {
        const int v = i / TW + HALO_M, u = i % TW + HALO_M, y = y0 + v, x = x0 + u;
        if (y >= H || x >= W) continue;
        const uint8_t* q = sa + v * RM_W + u;
        const uint32_t acc = (uint32_t)q[-RM_W - 1] + 2u * q[-RM_W] + q[-RM_W + 1] + 2u * q[-1] + 4u * q[0] + 2u * q[1] + q[RM_W - 1] + 2u * q[RM_W] + q[RM_W + 1];
        const uint32_t r = (acc + 8u) >> 4;                       // taps 64 128 64 (Q8): (acc * 4096 + 2^15) >> 16
        soft[((size_t)b * H + y) * W + x] = (uint8_t)r;
        mx = max(mx, r);
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((t & 31) == 0 && mx) atomicMax(&smax, mx);
    __syncthreads();
    if (t == 0 && smax) atomicMax(maxv + b, smax);
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
__global__ void __launch_bounds__(IT_THREADS)
blend_hresize_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ soft, const unsigned int* __restrict__ maxv, int planes_per_image,
                     int H, int W, int S, int ksize, const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, uint8_t* __restrict__ hp) {
    extern __shared__ uint8_t sm_dyn[];
    uint8_t* row = sm_dyn;                                        // [HR_ROWS][W]
    __shared__ uint8_t rescale[256];
    const int t = threadIdx.x, pl = blockIdx.y;
    const int y0 = blockIdx.x * HR_ROWS;
    const uint8_t* ap = a + (size_t)pl * H * W;
    if (soft) {
        // (e.astype(float) / e.max() * 255).astype(uint8) (float64, truncation); unchanged when the maximum is 0
        const unsigned int mx = maxv[pl / planes_per_image];
        rescale[t] = mx ? (uint8_t)__double2int_rz(__dmul_rn(__ddiv_rn((double)t, (double)mx), 255.0)) : (uint8_t)t;
        __syncthreads();
        const uint8_t* sp = soft + (size_t)pl * H * W;
        for (int i = t; i < HR_ROWS * W; i += IT_THREADS) {
            const int r = i / W, x = i - r * W, y = y0 + r;
            if (y >= H) continue;
            const float e = (float)rescale[sp[(size_t)y * W + x]];
            // cv2.addWeighted: fma(a, 0.7f, round_f32(e * 0.3f)), round half to even, saturate
            row[i] = sat_rint(__fmaf_rn((float)ap[(size_t)y * W + x], 0.7f, __fmul_rn(e, 0.3f)));
        }
    } else {
        for (int i = t; i < HR_ROWS * W; i += IT_THREADS) {
            const int r = i / W, x = i - r * W, y = y0 + r;
            if (y < H) row[i] = ap[(size_t)y * W + x];
        }
    }
    __syncthreads();
    for (int i = t; i < HR_ROWS * S; i += IT_THREADS) {
        const int r = i / S, xx = i - r * S, y = y0 + r;
        if (y >= H) continue;
        const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
        const int32_t* k = kk + (size_t)xx * ksize;
        int acc = 1 << 21;
        for (int j = 0; j < n; j++) acc += (int)row[r * W + xmin + j] * k[j];
        acc >>= 22;
        hp[((size_t)pl * H + y) * S + xx] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
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

struct ITLayout { size_t plane, lut, clahe, pmap, soft, hpass, maxv, ctl, tile_live, xb, xk, yb, yk, total; int ksx, ksy, htx, hty; };
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
    L.htx = (W + HT_W - 1) / HT_W; L.hty = (H + HT_H - 1) / HT_H;
    L.tile_live = take((size_t)B * L.htx * L.hty);
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
    unsigned int* maxv = (unsigned int*)(ws + L.maxv); HystCtl* ctl = (HystCtl*)(ws + L.ctl); uint8_t* tile_live = (uint8_t*)(ws + L.tile_live);
    int32_t* xb = (int32_t*)(ws + L.xb); int32_t* xk = (int32_t*)(ws + L.xk); int32_t* yb = (int32_t*)(ws + L.yb); int32_t* yk = (int32_t*)(ws + L.yk);
    ProfScope ps("itransform", s);

    resize_coeff_kernel<<<(S + 127) / 128, 128, 0, s>>>(W, S, L.ksx, xb, xk);
    KPD_LAUNCH_CHECK();
    resize_coeff_kernel<<<(S + 127) / 128, 128, 0, s>>>(H, S, L.ksy, yb, yk);
    KPD_LAUNCH_CHECK();
    clahe_lut_kernel<<<batch * planes * ntile, IT_THREADS, 0, s>>>(p, plane, lut);
    KPD_LAUNCH_CHECK();
    ClaheGeom g{tiles_x, tiles_y, 1.0f / (float)p.tw, 1.0f / (float)p.th};
    const dim3 tgrid((W + TW - 1) / TW, (H + TH - 1) / TH, batch * planes);
    const uint8_t* resize_src; const uint8_t* soft_arg = nullptr;
    if (grayscale) {
        clahe_canny_kernel<<<tgrid, IT_THREADS, 0, s>>>(plane, lut, g, H, W, 100, 200, clahe, pmap);
        KPD_LAUNCH_CHECK();
        KPD_CUDA(cudaMemsetAsync(maxv, 0, L.tile_live - L.maxv, s));                   // per-image maxima + hysteresis control block
        int dev = 0, sms = 0, per_sm = 0;
        KPD_CUDA(cudaGetDevice(&dev));
        KPD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        KPD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hysteresis_kernel, IT_THREADS, 0));
        const int ntiles = batch * L.htx * L.hty;
        const int hgrid = std::max(1, std::min(ntiles, sms * std::min(per_sm, 4)));
        int Bv = batch, Hv = H, Wv = W, htx = L.htx, hty = L.hty;
        void* args[] = {&pmap, &Bv, &Hv, &Wv, &htx, &hty, &tile_live, &ctl};
        KPD_CUDA(cudaLaunchCooperativeKernel((const void*)hysteresis_kernel, dim3(hgrid), dim3(IT_THREADS), args, 0, s));
        KPD_LAUNCH_CHECK();
        morph_kernel<<<tgrid, IT_THREADS, 0, s>>>(pmap, H, W, soft, maxv);
        KPD_LAUNCH_CHECK();
        resize_src = clahe; soft_arg = soft;
    } else {
        clahe_gauss3_kernel<<<tgrid, IT_THREADS, 0, s>>>(plane, lut, g, H, W, clahe);
        KPD_LAUNCH_CHECK();
        resize_src = clahe;
    }
    const size_t hsm = (size_t)HR_ROWS * W;
    KPD_CHECK_ARG(hsm <= 48 * 1024 - 512, "kpd_itransform: width %d too large for the row staging", W);
    blend_hresize_kernel<<<dim3((H + HR_ROWS - 1) / HR_ROWS, batch * planes), IT_THREADS, hsm, s>>>(resize_src, soft_arg, maxv, planes, H, W, S, L.ksx, xb, xk, hp);
    KPD_LAUNCH_CHECK();
    const float m[3] = {grayscale ? 0.5f : 0.485f, 0.456f, 0.406f}, sd[3] = {grayscale ? 0.5f : 0.229f, 0.224f, 0.225f};
    vresize_norm_kernel<<<dim3(std::max(1, std::min(64, (S * S + IT_THREADS - 1) / IT_THREADS)), batch * planes), IT_THREADS, 0, s>>>(
        hp, planes, H, S, L.ksy, yb, yk, m[0], m[1], m[2], sd[0], sd[1], sd[2], d_out);
    KPD_LAUNCH_CHECK();
    return 0;
}
