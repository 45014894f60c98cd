// Batch-wide pointwise (1x1) convolution on tcgen05 with fp32-level accuracy (3xTF32 split done on chip).
//   reference ops: every 1x1 conv of the trunk -- tv:models/mobilenetv3.py:87-108 (expand / project), features.12 -- and
//   dll/models/backbone.py:33-37 (FPN laterals 1..3 + nearest top-down add).
//
// out[M][N] = act((A * se_scale)[M][K] W[K][N] + bias) (+ residual), fp32 in / out.  The trunk needs fp32-level accuracy (the
// channel top-k is decided by score gaps of 1e-6..1e-5), which one tf32 MMA does not give; three do:
//     A W  ~=  A_hi W_hi + A_lo W_hi + A_hi W_lo,   x_hi = tf32(x),  x_lo = tf32(x - x_hi)       (error ~2^-21 relative)
// W_hi / W_lo are split on the host.  A is split ON CHIP: a 128-row x 32-float slab lands in shared memory by TMA (128-byte
// swizzle), four "split" warps rewrite it element-wise (A_hi in place, A_lo into a second slab at the same swizzled offsets (multiplying
// by the SE excitation on the way -- the swizzle is undone only to find each 16-byte chunk's k index), fence.proxy.async, and
// the MMA thread issues 3 x 4 tcgen05.mma (kind::tf32) per slab into a TMEM accumulator.  Even at three passes the tensor
// core has 5x the MAC rate of the 128 FFMA lanes, so these GEMMs become streams of their activations.
//
// Warp roles (14 warps): 0 TMA producer, 1 MMA issuer (+ TMEM alloc), 2-5 split, 6-13 epilogue (two warps per TMEM lane group on
// alternating 32-column chunks: tcgen05.ld -> bias / act -> shared-memory transpose -> line-coalesced float4 stores with the
// residual added on the way).  Persistent CTAs walk
// (128-row tile, column chunk <= 128) items; 2-3 slab stages; two TMEM accumulator stages.
#include "kpd_tma.cuh"
#include <vector>
#include <stdlib.h>

namespace kpd {

struct PwUParams {
    const float* bias; float* out; const float* res; const float* se_scale;
    uint32_t M; int K, N, Nc, nchunks, tiles_m, total, stages, stage_bytes;
    int act, hw, res_mode, H, W, up, round_tf32, Np;
    long long* trace;                 // debugging aid (KPD_PWU_TRACE=1): clock64 stamps of CTA 0 [role][event]
    float* out2; int split, n2;       // optional two-tensor output: columns [0, split) -> out [M][split], [split, split + n2) -> out2 [M][n2]
};

#define PU_TRACE(role, ev) do { if (p.trace && blockIdx.x == 0 && (ev) < 128) p.trace[(role) * 128 + (ev)] = clock64(); } while (0)
#ifndef KPD_PU_SPLIT_WARPS
#define KPD_PU_SPLIT_WARPS 8
#endif
constexpr int PU_SPLIT_WARPS = KPD_PU_SPLIT_WARPS, PU_SPLIT_THREADS = 32 * PU_SPLIT_WARPS, PU_EPI0 = 2 + PU_SPLIT_WARPS;   // warps: TMA, MMA, split, epilogue
static_assert(PU_SPLIT_WARPS == 4 || PU_SPLIT_WARPS == 8, "split warps: 4 or 8 (the epilogue warps must start at a multiple of 4 + 2)");
constexpr int PU_EPI_WARPS = 8, PU_THREADS = 64 + PU_SPLIT_THREADS + 32 * PU_EPI_WARPS, PU_RS = 36, PU_MAXN = 640;
constexpr int PU_SC_FLOATS = 4 * 576;                         // SE factors staged per tile: (images a 128-row tile can span) x K floats
constexpr int PU_A = 16384;                                   // one 128 x 32-float slab

template <typename ActFn>
__device__ __forceinline__ float4 pu_act4(float4 v, ActFn f) { v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w); return v; }
struct PuNone  { __device__ __forceinline__ float operator()(float x) const { return x; } };
struct PuRelu  { __device__ __forceinline__ float operator()(float x) const { return fmaxf(x, 0.f); } };
struct PuHswish { __device__ __forceinline__ float operator()(float x) const { return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f); } };

__global__ void __launch_bounds__(PU_THREADS, 1)
pw_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PwUParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // (offset arithmetic on the __shared__ array, not an integer round trip: shared-memory accesses must compile to LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.stages;
    // stage: [A (raw fp32 from TMA, rewritten in place as A_hi) | A_lo | B_hi (Nc rows x 128 B) | B_lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * p.stage_bytes);
    // bars: [0,S) full (TMA landed), [S,2S) split done, [2S,3S) free (MMAs retired), [3S,3S+2) tmem full, [3S+2,3S+4) tmem empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);
    // 16-byte aligned: stages are 1024-byte multiples, 3S+4 barriers of 8 bytes -> round the barrier block up to 16
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + ((8 * (3 * S + 4) + 16 + 15) & ~15));   // [PU_MAXN] bias of every (padded) output column
    float* red = sbias + PU_MAXN;                              // [8 warps][32][PU_RS] transpose tiles
    float* sscale = red + PU_EPI_WARPS * 32 * PU_RS;            // [PU_SC_IMGS][K] SE excitation factors of the current tile's images

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto split_bar = [&](int s) { return bar0 + 8u * (S + s); };
    auto free_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * S + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * S + 2 + a); };
    int tmem_cols = 32; while (tmem_cols < 2 * p.Nc + 16) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < S; s++) { mbar_init(full_bar(s), 1); mbar_init(split_bar(s), PU_SPLIT_WARPS); mbar_init(free_bar(s), 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), PU_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.Np; i += blockDim.x) sbias[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nslabs = (p.K + 31) >> 5;
    const uint32_t b_bytes = (uint32_t)p.Nc * 128u;

    if (warp == 0) {
        if (lane == 0) {                                // ---------------- TMA producer
            int stage = 0; uint32_t phase = 0;
            int tr0 = 0;
            for (int item = blockIdx.x; item < p.total; item += gridDim.x) {
                const int tm = item / p.nchunks, n0 = (item - tm * p.nchunks) * p.Nc;
                for (int sl = 0; sl < nslabs; sl++) {
                    mbar_wait(free_bar(stage), phase ^ 1);
                    PU_TRACE(0, tr0++);
                    mbar_expect_tx(full_bar(stage), (uint32_t)PU_A + 2u * b_bytes);
                    const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
                    tma_load_2d(sa, &tmA, full_bar(stage), sl * 32, tm * 128);
                    tma_load_2d(sa + 2 * PU_A, &tmB, full_bar(stage), sl * 32, n0);
                    tma_load_2d(sa + 2 * PU_A + b_bytes, &tmB, full_bar(stage), sl * 32, p.Np + n0);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Nc >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0; int tr1 = 0;
            for (int item = blockIdx.x; item < p.total; item += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Nc);
                for (int sl = 0; sl < nslabs; sl++) {
                    mbar_wait(split_bar(stage), phase);
                    PU_TRACE(1, tr1++);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
                    const uint64_t a_hi = make_sdesc_rb(sa, 128), a_lo = make_sdesc_rb(sa + PU_A, 128);
                    const uint64_t b_hi = make_sdesc_rb(sa + 2 * PU_A, 128), b_lo = make_sdesc_rb(sa + 2 * PU_A + b_bytes, 128);
                    const int ksteps = min(4, (p.K - sl * 32 + 7) >> 3);        // 8 floats of K per MMA
                    for (int k4 = 0; k4 < ksteps; k4++) {                        // small terms first
                        tc_mma<false>(d_tmem, a_lo + 2 * k4, b_hi + 2 * k4, idesc, (sl | k4) != 0 ? 1u : 0u);
                        tc_mma<false>(d_tmem, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
                        tc_mma<false>(d_tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, 1u);
                    }
                    tc_commit(free_bar(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp < PU_EPI0) {                        // ---------------- split warps: raw fp32 slab -> tf32 hi / lo slabs
        const int t = (warp - 2) * 32 + lane;           // float4 f = t + PU_SPLIT_THREADS * i lives in row f >> 3 at chunk position f & 7
        constexpr int NI = 1024 / PU_SPLIT_THREADS;      // float4 per thread per slab
        int stage = 0; uint32_t phase = 0; int tr2 = 0;
        for (int item = blockIdx.x; item < p.total; item += gridDim.x) {
            const int tm = item / p.nchunks;
            const uint32_t m0 = (uint32_t)tm * 128u;
            // SE excitation factors of the tile's images: staged once per item ([image][K], the images of a tile are consecutive, so
            // it is one contiguous block) -- as per-slab global loads they doubled the kernel's L2 requests and put an L2 round trip
            // into every slab of the serial split stage
            const float* srow[8];
            if (p.se_scale) {
                const uint32_t img0 = m0 / (uint32_t)p.hw, img1 = min(m0 + 127u, p.M - 1) / (uint32_t)p.hw;
                const int nv = (int)(img1 - img0 + 1) * (p.K >> 2);
                asm volatile("bar.sync 2, %0;" ::"n"(PU_SPLIT_THREADS) : "memory");   // every split thread is done with the previous item's factors
                const float4* src = reinterpret_cast<const float4*>(p.se_scale + (size_t)img0 * p.K);
                for (int i = t; i < nv; i += PU_SPLIT_THREADS) reinterpret_cast<float4*>(sscale)[i] = __ldg(src + i);
                asm volatile("bar.sync 2, %0;" ::"n"(PU_SPLIT_THREADS) : "memory");
#pragma unroll
                for (int i = 0; i < NI; i++) srow[i] = sscale + (size_t)(min(m0 + (uint32_t)((t >> 3) + (PU_SPLIT_THREADS / 8) * i), p.M - 1) / (uint32_t)p.hw - img0) * p.K;
            }
            for (int sl = 0; sl < nslabs; sl++) {
                mbar_wait(full_bar(stage), phase);
                if (t == 0) PU_TRACE(2, tr2++);
                uint8_t* sp = smem + (size_t)stage * p.stage_bytes;
#pragma unroll
                for (int i = 0; i < NI; i++) {
                    const int f = t + PU_SPLIT_THREADS * i;                      // float4 index: row f >> 3, stored chunk position f & 7
                    float4 a = *reinterpret_cast<const float4*>(sp + f * 16);
                    if (p.se_scale) {
                        const int row = f >> 3, k = sl * 32 + (((f & 7) ^ (row & 7)) << 2);      // 128-byte swizzle: chunk position -> k
                        const float4 sc = k < p.K ? *reinterpret_cast<const float4*>(srow[i] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        a.x *= sc.x; a.y *= sc.y; a.z *= sc.z; a.w *= sc.w;
                    }
                    // x = hi + lo: hi = x with the 13 low mantissa bits cleared (what the tensor core reads of an fp32 operand), lo = x - hi
                    // (exact; the tensor core truncates it to tf32 itself): |x - hi - tf32(lo)| <= 2^-20 |x| for one LOP3 + one FADD per element
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u);
                    hi.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u);
                    lo.x = a.x - hi.x; lo.y = a.y - hi.y; lo.z = a.z - hi.z; lo.w = a.w - hi.w;
                    *reinterpret_cast<float4*>(sp + f * 16) = hi;                  // in place: same thread, same 16 bytes
                    *reinterpret_cast<float4*>(sp + PU_A + f * 16) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(split_bar(stage));
                if (t == 0) PU_TRACE(2, tr2++);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else {                                            // ---------------- epilogue warps: two per TMEM lane group, alternating 32-column chunks
        const int lg = warp & 3;                        // TMEM lane group this warp may access
        const int cpar = (warp - PU_EPI0) >> 2;         // chunk parity this warp drains
        const int rr4 = lane >> 3, q = lane & 7;
        float* rt = red + (warp - PU_EPI0) * (32 * PU_RS);
        int acc = 0; uint32_t acc_phase = 0; int tr3 = 0;
        for (int item = blockIdx.x; item < p.total; item += gridDim.x) {
            const int tm = item / p.nchunks, n0 = (item - tm * p.nchunks) * p.Nc;
            const uint32_t m0 = (uint32_t)tm * 128u;
            // per-row output / residual offsets of the 8 rows this lane stores (row = lg*32 + it*4 + rr4), once per item
            size_t ooff[8], roff[8];
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const uint32_t m = min(m0 + (uint32_t)(lg * 32 + it * 4 + rr4), p.M - 1);
                ooff[it] = (size_t)m * p.N;
                roff[it] = ooff[it];
                if (p.res_mode == 2) {
                    const uint32_t rowi = m / (uint32_t)p.W, xx = m - rowi * p.W, b = rowi / (uint32_t)p.H, yy = rowi - b * p.H;
                    roff[it] = ((size_t)(b * (p.H / p.up) + yy / p.up) * (p.W / p.up) + xx / p.up) * p.N;
                }
            }
            if (warp == PU_EPI0 && lane == 0) PU_TRACE(3, tr3++);
            mbar_wait_warp(tfull_bar(acc), acc_phase, lane);
            if (warp == PU_EPI0 && lane == 0) PU_TRACE(3, tr3++);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * p.Nc);
            auto drain = [&](auto actfn) {
                for (int c0 = cpar * 32; c0 < p.Nc; c0 += 64) {
                    uint32_t r[32];
                    tc_ld32(taddr + c0, r);
                    tc_wait_ld();
                    if (warp == PU_EPI0 && lane == 0) PU_TRACE(3, tr3++);
                    const float* bp = sbias + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 bv = *reinterpret_cast<const float4*>(bp + 4 * j);
                        float4 v;
                        v.x = __uint_as_float(r[4 * j + 0]) + bv.x; v.y = __uint_as_float(r[4 * j + 1]) + bv.y;
                        v.z = __uint_as_float(r[4 * j + 2]) + bv.z; v.w = __uint_as_float(r[4 * j + 3]) + bv.w;
                        *reinterpret_cast<float4*>(rt + lane * PU_RS + 4 * j) = pu_act4(v, actfn);
                    }
                    __syncwarp();
                    if (warp == PU_EPI0 && lane == 0) PU_TRACE(3, tr3++);
                    // transposed read: lane = (row rr4 of 4, 4 channels q): one warp store covers four 128-byte row segments
                    const int n = n0 + c0 + 4 * q;
                    if (n < p.N && c0 + 4 * q < p.Nc) {                      // columns past the chunk belong to another item
#pragma unroll
                        for (int it = 0; it < 8; it++) {
                            const int rr = it * 4 + rr4;
                            if (m0 + (uint32_t)(lg * 32 + rr) < p.M) {
                                float4 x = *reinterpret_cast<const float4*>(rt + rr * PU_RS + 4 * q);
                                if (p.res_mode) {
                                    const float4 rv = __ldg(reinterpret_cast<const float4*>(p.res + roff[it] + n));
                                    x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
                                }
                                if (p.round_tf32) { x.x = tf32_rna(x.x); x.y = tf32_rna(x.y); x.z = tf32_rna(x.z); x.w = tf32_rna(x.w); }
                                if (p.out2) {                               // person heads: box | cls columns to two row-major tensors
                                    const size_t m = ooff[it] / (size_t)p.N;
                                    if (n < p.split) *reinterpret_cast<float4*>(p.out + m * p.split + n) = x;
                                    else {
                                        float* o2 = p.out2 + m * p.n2 + (n - p.split);
                                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                        for (int e = 0; e < 4; e++) if (n - p.split + e < p.n2) o2[e] = xs[e];
                                    }
                                } else
                                *reinterpret_cast<float4*>(p.out + ooff[it] + n) = x;
                            }
                        }
                    }
                    __syncwarp();
                }
            };
            if (p.act == ACT_RELU) drain(PuRelu()); else if (p.act == ACT_HSWISH) drain(PuHswish()); else drain(PuNone());
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (warp == PU_EPI0 && lane == 0) PU_TRACE(3, tr3++);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// column chunking of the N outputs: nchunks chunks of Nc (multiple of 16, <= 128) columns, minimising the padded width
void pwu_chunks(int N, int* nchunks, int* Nc) {
    int best_n = 0, best_c = 0, best_w = 1 << 30;
    for (int n = (N + 127) / 128; n <= (N + 127) / 128 + 2; n++) {
        const int c = (((N + n - 1) / n) + 15) & ~15;
        if (c > 128) continue;
        if (n * c < best_w) { best_w = n * c; best_n = n; best_c = c; }
    }
    *nchunks = best_n; *Nc = best_c;
}

int pointwise_umma(const PwMma& a, int sm_count, cudaStream_t s) {
    if (!a.Wu || (a.K & 3) || (a.N & 3)) { set_error("pointwise_umma: missing split weights or unsupported shape %d x %d", a.K, a.N); return KPD_ERR_UNSUPPORTED; }
    if (a.M <= 0) return 0;
    if (a.M >= ((int64_t)1 << 31) - 128) { set_error("pointwise_umma: M = %lld pixels exceeds the 32-bit pixel index", (long long)a.M); return KPD_ERR_UNSUPPORTED; }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    PwUParams p{};
    p.bias = a.bias; p.out = a.out; p.res = a.res; p.se_scale = a.se_scale;
    p.M = (uint32_t)a.M; p.K = a.K; p.N = a.N;
    pwu_chunks(a.N, &p.nchunks, &p.Nc);
    p.Np = p.nchunks * p.Nc;
    const int Kp = (a.K + 31) & ~31;
    p.tiles_m = (int)((a.M + 127) / 128);
    const int64_t total = (int64_t)p.tiles_m * p.nchunks;
    if (total >= ((int64_t)1 << 31)) { set_error("pointwise_umma: too many tiles"); return KPD_ERR_UNSUPPORTED; }
    p.total = (int)total;
    p.out2 = a.out2; p.split = a.split; p.n2 = a.n2;
    p.act = a.act; p.hw = a.hw > 0 ? a.hw : 1; p.res_mode = a.res ? a.res_mode : 0; p.H = a.H; p.W = a.W; p.up = a.up > 0 ? a.up : 1; p.round_tf32 = a.round_tf32;
    p.stage_bytes = 2 * PU_A + 2 * p.Nc * 128;
    const int misc = 1024 /*align*/ + 8 * (3 * 8 + 4) + 32 + 4 * PU_MAXN + 4 * (PU_EPI_WARPS * 32 * PU_RS) + 4 * PU_SC_FLOATS + 64;
    if (a.se_scale && (int64_t)((127 + p.hw - 1) / p.hw + 1) * a.K > PU_SC_FLOATS) { set_error("pointwise_umma: SE factors of a tile (%d px/image, K = %d) exceed the staging buffer", p.hw, a.K); return KPD_ERR_UNSUPPORTED; }
    if (p.Np > PU_MAXN) { set_error("pointwise_umma: N = %d too wide", a.N); return KPD_ERR_UNSUPPORTED; }
    int stages = (227 * 1024 - misc) / p.stage_bytes; if (stages > 6) stages = 6;
    if (stages < 2) { set_error("pointwise_umma: not enough shared memory"); return KPD_ERR_UNSUPPORTED; }
    p.stages = stages;
    const size_t smem = (size_t)stages * p.stage_bytes + misc;

    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
        cuuint64_t strides[1] = {(cuuint64_t)a.K * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("pointwise_umma: cuTensorMapEncodeTiled(A) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * p.Np)};
        cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)p.Nc};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.Wu), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("pointwise_umma: cuTensorMapEncodeTiled(W) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    static uint64_t attr = 0;
    KPD_CUDA(ensure_smem_attr(pw_umma_kernel, 227 * 1024, attr));
    const int grid = (int)std::min<int64_t>(total, (int64_t)sm_count);
    static const bool trace_on = getenv("KPD_PWU_TRACE") != nullptr;
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) cudaMalloc(&trace_buf, 4 * 128 * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 4 * 128 * sizeof(long long), s);
        p.trace = trace_buf;
    }
    pw_umma_kernel<<<grid, PU_THREADS, smem, s>>>(tmA, tmB, p);
    KPD_LAUNCH_CHECK();
    if (trace_on) {                                          // debugging only: synchronous dump of CTA 0's timeline (SM clocks)
        std::vector<long long> h(4 * 128);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        long long t0 = 0; for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        const char* roles[4] = {"producer(free ok)", "mma(split ok)", "split(full ok, done)", "epilogue(start, tfull ok, done)"};
        fprintf(stderr, "kpd pwu trace: M=%u K=%d N=%d Nc=%d items=%d stages=%d grid=%d\n", p.M, p.K, p.N, p.Nc, p.total, p.stages, grid);
        for (int r = 0; r < 4; r++) {
            fprintf(stderr, "  %s:", roles[r]);
            for (int e = 0; e < 60; e++) if (h[r * 128 + e]) fprintf(stderr, " %lld", h[r * 128 + e] - t0);
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

}  // namespace kpd
