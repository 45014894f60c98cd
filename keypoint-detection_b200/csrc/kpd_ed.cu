// Fused front half of an inverted-residual block: expand 1x1 (+BN+act) -> depthwise k x k (+BN+act) -> SE squeeze partials,
// with the expanded tensor never leaving the SM.
//   reference ops: tv:models/mobilenetv3.py:87-104 (the expand and depthwise Conv2dNormActivation of InvertedResidual) and
//   the AdaptiveAvgPool2d of its SqueezeExcitation (tv:ops/misc.py:243-251), as used by dll/models/backbone.py:247-264.
//
// The expanded tensor is the largest intermediate of the trunk (3-6x the block input; 231 MB at features.2, B = 256) and in
// the layer-by-layer pipeline it is written by the expand GEMM and read back by the depthwise kernel.  Here a work item is
// (group of G images, band of TH output rows): the band's input pixels (full width, halo rows included, rows outside the image
// clamped away) are ONE contiguous run of the NHWC tensor, so they arrive as plain 2-D TMA boxes [128 pixels x K] in the
// canonical K-major swizzled UMMA layout.  Per item:
//   producer warp : TMA of the band (raw fp32) into a 2-deep ring; per 32-channel chunk the expand weights (tf32 hi / lo rows)
//                   and one parameter block (expand bias, depthwise bias, k*k tap rows) into a 3-deep ring
//   split warps   : raw -> tf32 hi (in place) + lo (x = hi + lo to ~2^-21: the 3xTF32 scheme of kpd_pwu.cu)
//   MMA thread    : per chunk and M-tile 3 x ceil(K/8) tcgen05.mma (kind::tf32, N = 32) into a double-buffered TMEM accumulator
//   compute warps : (8 warps) tcgen05.ld -> +bias, activation -> the chunk's expanded tile in shared memory
//                   [image][row + halo][col + halo][32 ch + 4 pad], zero where the depthwise conv pads -> depthwise conv with a
//                   lane per channel and a k-row register window (the inner loop of kpd_dw.cu) -> coalesced 128-byte stores of
//                   the block's depthwise output, per-(image, output row) squeeze partials in a canonical order
// While the compute warps work on chunk c the tensor core already holds chunk c+1 and the TMA ring chunk c+2..3; the next
// item's pixels land during the current item.  Only the depthwise output (what the SE excitation and the projection need)
// goes back to memory.
//
// Measured (B = 256, 224 x 224, clock64 traces of KPD_ED_TRACE=1 and ncu, profiles/r02_*): the kernel is bound by the
// compute warps' FP32 work, not by the tensor core or the TMA rings -- a 3-register FFMA issues once per 2 cycles per SM
// sub-partition on this part, a 5x5 row of 2 pixels is 50 of them, and the per-chunk skeleton (TMEM drain, barrier, unit
// set-up) is about as long again on the 14 x 14 / 7 x 7 maps.  It runs the ten expand + depthwise pairs of the trunk in
// 0.91 ms against 0.77 ms + the expanded tensor's round trip for the separate kernels: time parity, with the 0.9 GB of
// expanded-tensor traffic gone.  Packed FFMA2 over channel pairs is the next step for the depthwise loop.
#define KPD_WAIT_SLEEP_NS 20    // the compute warps are the critical path here: wake up quickly
#include "kpd_tma.cuh"
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include <algorithm>

namespace kpd {

constexpr int ED_CW = 32;            // expanded channels per chunk (= N of one MMA)
constexpr int ED_ES = 36;            // floats per pixel of the expanded tile: 32 + 4 padding => conflict-free row stores and lane = channel loads
// Eight compute warps (two per SM sub-partition) measured faster than sixteen (0.86 vs 0.96 ms over the ten layers): the phase is
// bound by instruction issue and per-chunk hand-offs, not by latency hiding, and at <= 146 registers per thread the 5x5 window loop
// neither spills nor needs the narrow 2-column form.  (-DKPD_ED_CWARPS=16 rebuilds the old shape for A/B runs.)
#ifndef KPD_ED_CWARPS
#define KPD_ED_CWARPS 8
#endif
constexpr int ED_CWARPS = KPD_ED_CWARPS, ED_CTHREADS = 32 * ED_CWARPS, ED_EC = 32 / (ED_CWARPS / 4);   // ED_EC: chunk columns a compute warp drains
static_assert(ED_CWARPS == 8 || ED_CWARPS == 16, "compute warps: 2 or 4 per TMEM lane group");
constexpr int ED_CTHREADS_ = ED_CTHREADS;   // compute warps: four per SM sub-partition hide each other's latencies
constexpr int ED_THREADS = 192 + ED_CTHREADS;                 // 14 warps: TMA, MMA, 4 split, 8 compute
constexpr int ED_MAXT = 8, ED_MAXG = 8, ED_MAXBST = 4;

struct EdParams {
    float* out; float* pool;         // depthwise output [B][Ho][Wo][C]; squeeze partials [B][Ho][C], one per output row (may be null)
    const float* params;             // [nchunks][2 + k*k][32]: expand bias | depthwise bias | tap rows
    int B, H, W, Ho, Wo, C, K, act;
    int G, TH, bands, total_items, nchunks;
    int RB, slabs, ksteps;           // bytes of K per operand row (64 | 128), K slabs, 8-float MMA steps = ceil(K / 8)
    int P8, T, tail_rows, part_bytes;   // operand rows per slab, M tiles of a full item, rows of the last tile's TMA box, slabs * P8 * RB
    int in_h, Wp, ebuf_floats, nbuf, YS;
    int bstages, bstage_bytes, b_part_bytes, par_bytes;
    int tmem_cols;
    int off_b, off_e, off_red, off_bar;
    long long* trace;                // debugging aid (KPD_ED_TRACE=1): clock64 stamps of CTA 0 [role][event]
};

// Waits of the helper roles (TMA / MMA threads, split warps): they have slack, and a spinning warp takes issue slots from the
// compute warps of its SM sub-partition -- poll with a short sleep.
__device__ __forceinline__ void ed_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0; long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if ((++spins & 4095u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) { printf("kpd: ed mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
        }
    }
}

#define ED_TRACE(role, ev) do { if (p.trace && blockIdx.x == 0 && (ev) < 256) p.trace[(role) * 256 + (ev)] = clock64(); } while (0)

// output columns per depthwise work unit: 4 (measured: 2 and 3 columns for the 5x5 layers and 7 columns for either kernel size are
// slower -- 7 spills the 5x5 window, 2 doubles the per-unit set-up)
#ifndef KPD_ED_XR5
#define KPD_ED_XR5 4
#endif
#ifndef KPD_ED_XR3
#define KPD_ED_XR3 4
#endif
__host__ __device__ constexpr int ed_xr(int K, int S) { return K == 5 ? KPD_ED_XR5 : KPD_ED_XR3; }

template <int KS, int S, bool HSW>       // HSW: hardswish (else ReLU) after the expand and the depthwise conv
__global__ void __launch_bounds__(ED_THREADS, 1)
ed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAt, const __grid_constant__ CUtensorMap tmB, const EdParams p) {
    constexpr int KK = KS * KS, pad = (KS - 1) / 2, XR = ed_xr(KS, S), SEG = (XR - 1) * S + KS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // (offset arithmetic on the __shared__ array, not an integer round trip: the tile loads / stores must compile to LDS / STS)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                     // [3 parts][slabs][P8 rows][RB]: parts 0 / 1 = raw -> hi ring, part 2 = lo
    uint8_t* sB = smem + p.off_b;                           // [bstages][hi rows | lo rows | params]
    float* sE = reinterpret_cast<float*>(smem + p.off_e);   // [nbuf][G][in_h][Wp][ED_ES]
    float* rowp = reinterpret_cast<float*>(smem + p.off_red);  // [G][TH][runs][32] squeeze partials of the current chunk, one per (output row, run)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
    const int BST = p.bstages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * ED_MAXBST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto raw_full = [&](int i) { return bar0 + 8u * i; };
    auto raw_empty = [&](int i) { return bar0 + 8u * (2 + i); };
    const uint32_t lo_free = bar0 + 8u * 4, op_full = bar0 + 8u * 5;
    auto acc_full = [&](int i) { return bar0 + 8u * (6 + i); };
    auto acc_empty = [&](int i) { return bar0 + 8u * (8 + i); };
    auto b_full = [&](int i) { return bar0 + 8u * (10 + i); };
    auto b_empty = [&](int i) { return bar0 + 8u * (10 + ED_MAXBST + i); };

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAt) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int i = 0; i < 2; i++) { mbar_init(raw_full(i), 1); mbar_init(raw_empty(i), 1); mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), ED_CWARPS); }
        mbar_init(lo_free, 1); mbar_init(op_full, 4);
        for (int i = 0; i < BST; i++) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1 + ED_CWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // item -> (first image, images, first output row, output rows, first (possibly negative) input row, first in-image row,
    //          pixels, index of the first pixel)
    struct Item { int b0, gcnt, band, oy0, th, iy0, r_lo, r_hi, P; uint32_t pix0; };
    auto decode = [&](int item) {
        Item it;
        it.band = item % p.bands; const int grp = item / p.bands;
        it.b0 = grp * p.G; it.gcnt = min(p.G, p.B - it.b0);
        it.oy0 = it.band * p.TH; it.th = min(p.TH, p.Ho - it.oy0);
        it.iy0 = it.oy0 * S - pad;
        const int iy1 = (it.oy0 + it.th - 1) * S - pad + KS - 1;
        it.r_lo = max(it.iy0, 0); it.r_hi = min(iy1, p.H - 1);
        it.P = p.bands == 1 ? it.gcnt * p.H * p.W : (it.r_hi - it.r_lo + 1) * p.W;
        it.pix0 = ((uint32_t)it.b0 * (uint32_t)p.H + (uint32_t)it.r_lo) * (uint32_t)p.W;
        return it;
    };
    const int step = gridDim.x;
    const int slab_bytes = p.P8 * p.RB;                     // one K slab of one operand part

    if (warp == 0) {
        if (lane == 0) {                                    // ---------------- TMA producer
            auto load_raw = [&](int item, int nn) {
                const Item it = decode(item);
                const int slot = nn & 1; const uint32_t ph = (uint32_t)(nn >> 1) & 1u;
                ed_wait(raw_empty(slot), ph ^ 1u);
                const int Tit = (it.P + 127) >> 7;
                const int last_rows = (Tit == p.T) ? p.tail_rows : 128;
                mbar_expect_tx(raw_full(slot), (uint32_t)(p.slabs * ((Tit - 1) * 128 + last_rows) * p.RB));
                const uint32_t base = smem_u32(sA + (size_t)slot * p.part_bytes);
                for (int sl = 0; sl < p.slabs; sl++)
                    for (int t = 0; t < Tit; t++) {
                        const uint32_t dst = base + (uint32_t)(sl * slab_bytes + t * 128 * p.RB);
                        const CUtensorMap* tm = (t == Tit - 1 && Tit == p.T) ? &tmAt : &tmA;
                        tma_load_2d(dst, tm, raw_full(slot), sl * (p.RB >> 2), (int)(it.pix0 + (uint32_t)t * 128u));
                    }
            };
            int s = 0; uint32_t ph = 0;
            if ((int)blockIdx.x < p.total_items) load_raw(blockIdx.x, 0);
            int n = 0;
            for (int item = blockIdx.x; item < p.total_items; item += step, n++) {
                if (item + step < p.total_items) load_raw(item + step, n + 1);
                for (int c = 0; c < p.nchunks; c++) {
                    ED_TRACE(3, (n * p.nchunks + c) * 2);
                    ed_wait(b_empty(s), ph ^ 1u);
                    ED_TRACE(3, (n * p.nchunks + c) * 2 + 1);
                    mbar_expect_tx(b_full(s), (uint32_t)(2 * p.b_part_bytes + p.par_bytes));
                    const uint32_t dst = smem_u32(sB + (size_t)s * p.bstage_bytes);
                    for (int part = 0; part < 2; part++)
                        for (int sl = 0; sl < p.slabs; sl++)
                            tma_load_2d(dst + (uint32_t)(part * p.b_part_bytes + sl * ED_CW * p.RB), &tmB, b_full(s), sl * (p.RB >> 2),
                                        (part * p.nchunks + c) * ED_CW);
                    bulk_load_1d(dst + 2u * (uint32_t)p.b_part_bytes, p.params + (size_t)c * (2 + KK) * ED_CW, (uint32_t)p.par_bytes, b_full(s));
                    if (++s == BST) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                    // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ED_CW >> 3) << 17) | ((128u >> 4) << 24);
            // The issuing thread's own instruction stream bounds the MMA rate, so everything is reduced to 64-bit adds on
            // descriptors built once: desc(base + off) = desc(base) + (off >> 4) (addresses < 256 KB: no carry out of the field)
            const int kps = p.RB >> 5;                      // 8-float MMA steps per slab row
            const uint32_t a_slab16 = (uint32_t)slab_bytes >> 4, a_tile16 = (uint32_t)(128 * p.RB) >> 4, b_slab16 = (uint32_t)(ED_CW * p.RB) >> 4;
            const uint64_t aLo0 = make_sdesc_rb(smem_u32(sA + 2 * (size_t)p.part_bytes), p.RB);
            uint32_t q = 0; int n = 0;
            int s = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < p.total_items; item += step, n++) {
                const Item it = decode(item);
                const int Tit = (it.P + 127) >> 7;
                ed_wait(op_full, (uint32_t)n & 1u);
                tc_fence_after();
                const uint64_t aHi0 = make_sdesc_rb(smem_u32(sA + (size_t)(n & 1) * p.part_bytes), p.RB);
                for (int c = 0; c < p.nchunks; c++, q++) {
                    const int a = (int)(q & 1u); const uint32_t aph = (q >> 1) & 1u;
                    ED_TRACE(1, (int)q * 4);
                    ed_wait(acc_empty(a), aph ^ 1u);
                    ED_TRACE(1, (int)q * 4 + 1);
                    ed_wait(b_full(s), ph);
                    ED_TRACE(1, (int)q * 4 + 2);
                    tc_fence_after();
                    const uint64_t bHi0 = make_sdesc_rb(smem_u32(sB + (size_t)s * p.bstage_bytes), p.RB);
                    const uint64_t bLo0 = bHi0 + ((uint32_t)p.b_part_bytes >> 4);
                    uint32_t d = tmem_base + (uint32_t)(a * p.T * ED_CW);
                    uint64_t ah_t = aHi0, al_t = aLo0;
                    for (int t = 0; t < Tit; t++, d += ED_CW, ah_t += a_tile16, al_t += a_tile16) {
                        uint64_t ah = ah_t, al = al_t, bh = bHi0, bl = bLo0;
                        int left = p.ksteps;
                        uint32_t accum = 0u;
                        for (int sl = 0; sl < p.slabs; sl++, ah += a_slab16, al += a_slab16, bh += b_slab16, bl += b_slab16) {
                            const int nk = min(kps, left); left -= nk;
                            for (int k4 = 0; k4 < nk; k4++) {
                                tc_mma<false>(d, al + 2 * k4, bh + 2 * k4, idesc, accum);      // small terms first
                                tc_mma<false>(d, ah + 2 * k4, bl + 2 * k4, idesc, 1u);
                                tc_mma<false>(d, ah + 2 * k4, bh + 2 * k4, idesc, 1u);
                                accum = 1u;
                            }
                        }
                    }
                    tc_commit(b_empty(s));
                    tc_commit(acc_full(a));
                    ED_TRACE(1, (int)q * 4 + 3);
                    if (++s == BST) { s = 0; ph ^= 1u; }
                }
                tc_commit(raw_empty(n & 1));                // the item's operands are free once every MMA above has retired
                tc_commit(lo_free);
            }
        }
    } else if (warp < 6) {                                  // ---------------- split warps: raw fp32 -> tf32 hi (in place) + lo
        const int t = (warp - 2) * 32 + lane;
        int n = 0;
        for (int item = blockIdx.x; item < p.total_items; item += step, n++) {
            const Item it = decode(item);
            const int rows = min(p.P8, (it.P + 7) & ~7);
            const int nvec = rows * (p.RB >> 4);             // 16-byte vectors per slab in use
            if (t == 0) ED_TRACE(2, n * 4);
            mbar_wait_warp(raw_full(n & 1), (uint32_t)(n >> 1) & 1u, lane);
            if (t == 0) ED_TRACE(2, n * 4 + 1);
            mbar_wait_warp(lo_free, ((uint32_t)n & 1u) ^ 1u, lane);
            if (t == 0) ED_TRACE(2, n * 4 + 2);
            uint8_t* hp = sA + (size_t)(n & 1) * p.part_bytes;
            uint8_t* lp = sA + 2 * (size_t)p.part_bytes;
            for (int sl = 0; sl < p.slabs; sl++) {
                const size_t so = (size_t)sl * slab_bytes;
                for (int f = t; f < nvec; f += 128) {
                    const float4 a = *reinterpret_cast<const float4*>(hp + so + (size_t)f * 16);
                    float4 hi, lo;
                    hi.x = tf32_rna(a.x); hi.y = tf32_rna(a.y); hi.z = tf32_rna(a.z); hi.w = tf32_rna(a.w);
                    lo.x = tf32_rna(a.x - hi.x); lo.y = tf32_rna(a.y - hi.y); lo.z = tf32_rna(a.z - hi.z); lo.w = tf32_rna(a.w - hi.w);
                    *reinterpret_cast<float4*>(hp + so + (size_t)f * 16) = hi;
                    *reinterpret_cast<float4*>(lp + so + (size_t)f * 16) = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(op_full);
            if (t == 0) ED_TRACE(2, n * 4 + 3);
        }
    } else {                                                // ---------------- compute warps 6..13
        const int cw = warp - 6, ctid = threadIdx.x - 192;
        const int lg = warp & 3, cpart = cw >> 2;           // TMEM lane group of this warp; which 8 of the chunk's 32 columns it drains
        constexpr bool hsw = HSW;
        const int rowstride = p.Wp * ED_ES;
        const int img_floats = p.in_h * rowstride;
        {   // the tile starts as zeros: halo columns and the rows above / below a whole image are never written again
            const int total = p.nbuf * p.ebuf_floats;
            for (int i = ctid * 4; i < total; i += ED_CTHREADS * 4) *reinterpret_cast<float4*>(sE + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ED_CTHREADS) : "memory");
        uint32_t q = 0;
        int bs = 0; uint32_t bph = 0;
        // Squeeze partials in a CANONICAL order, so that an image's SE excitation does not depend on how this launch was cut into
        // bands / image groups / work units (the trunk stays bit-identical across batch sizes): a thread stores the sum of its
        // XR outputs of ONE output row; here the runs of a row are added left to right and the row total goes to
        // pool[image][output row][channel]; the SE kernel adds the rows top to bottom.
        auto finish_pool = [&](int c_done, const Item& it, int runs) {
            if (!p.pool) return;
            const int per_g = it.th * 32;
            for (int i = ctid; i < it.gcnt * per_g; i += ED_CTHREADS) {
                const int g = i / per_g, rem = i - g * per_g, y = rem >> 5, ch = rem & 31;
                if (c_done * ED_CW + ch < p.C) {
                    const float* rp = rowp + ((g * p.TH + y) * runs) * 32 + ch;
                    float tot = 0.f;
                    for (int r = 0; r < runs; r++) tot += rp[r * 32];
                    p.pool[((size_t)(it.b0 + g) * p.Ho + it.oy0 + y) * p.C + c_done * ED_CW + ch] = tot;
                }
            }
        };
        for (int item = blockIdx.x; item < p.total_items; item += step) {
            const Item it = decode(item);
            const int Tit = (it.P + 127) >> 7;
            if (p.bands > 1) {
                // banded items: the rows of the tile that fall outside the image differ from item to item -- re-zero them
                // (all warps are past the previous item's depthwise reads after this barrier)
                asm volatile("bar.sync 1, %0;" ::"n"(ED_CTHREADS) : "memory");
                const int top = it.r_lo - it.iy0, bot0 = it.r_hi - it.iy0 + 1;
                for (int b = 0; b < p.nbuf; b++) {
                    float* eb = sE + (size_t)b * p.ebuf_floats;
                    for (int i = ctid * 4; i < top * rowstride; i += ED_CTHREADS * 4) *reinterpret_cast<float4*>(eb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int i = bot0 * rowstride + ctid * 4; i < p.in_h * rowstride; i += ED_CTHREADS * 4) *reinterpret_cast<float4*>(eb + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // where this thread's accumulator rows (pixel m = t*128 + lg*32 + lane of the item) live in the expanded tile
            int eoff[ED_MAXT];
#pragma unroll
            for (int t = 0; t < ED_MAXT; t++) {
                const int m = t * 128 + lg * 32 + lane;
                eoff[t] = -1;
                if (t < Tit && m < it.P) {
                    int g = 0, r, x;
                    if (p.bands == 1) { const int hw = p.H * p.W; g = m / hw; const int rem = m - g * hw; r = rem / p.W; x = rem - r * p.W; }
                    else { const int rr = m / p.W; r = it.r_lo + rr; x = m - rr * p.W; }
                    eoff[t] = g * img_floats + (r - it.iy0) * rowstride + (x + pad) * ED_ES;
                }
            }
            const int runs = (p.Wo + XR - 1) / XR, ysegs = (it.th + p.YS - 1) / p.YS, per_img = ysegs * runs, nwork = it.gcnt * per_img;
            for (int c = 0; c < p.nchunks; c++, q++) {
                const int s = bs; const uint32_t ph = bph;
                if (++bs == BST) { bs = 0; bph ^= 1u; }
                const int a = (int)(q & 1u); const uint32_t aph = (q >> 1) & 1u;
                const int tev = (int)q * 6;
                if (cw == 0 && lane == 0) ED_TRACE(0, tev);
                mbar_wait_warp(b_full(s), ph, lane);
                if (cw == 0 && lane == 0) ED_TRACE(0, tev + 1);
                const float* par = reinterpret_cast<const float*>(sB + (size_t)s * p.bstage_bytes + 2 * (size_t)p.b_part_bytes);
                const float bdw = par[ED_CW + lane];
                float w[KK];
#pragma unroll
                for (int t = 0; t < KK; t++) w[t] = par[(2 + t) * ED_CW + lane];
                float4 be[ED_EC / 4];
#pragma unroll
                for (int j = 0; j < ED_EC / 4; j++) be[j] = *reinterpret_cast<const float4*>(par + cpart * ED_EC + 4 * j);
                __syncwarp();
                if (lane == 0) mbar_arrive(b_empty(s));      // the chunk's parameters are in registers
                mbar_wait_warp(acc_full(a), aph, lane);
                if (cw == 0 && lane == 0) ED_TRACE(0, tev + 2);
                tc_fence_after();
                float* eb = sE + (p.nbuf == 2 ? (size_t)(q & 1u) * p.ebuf_floats : 0);
                // ---- expand epilogue: accumulators -> +bias, activation -> expanded tile
#pragma unroll
                for (int t = 0; t < ED_MAXT; t++) {
                    if (t < Tit) {
                        uint32_t r[ED_EC];
                        if constexpr (ED_EC == 8) tc_ld8(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((a * p.T + t) * ED_CW + cpart * ED_EC), r);
                        else tc_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((a * p.T + t) * ED_CW + cpart * ED_EC), r);
                        tc_wait_ld();
                        if (eoff[t] >= 0) {
                            float* dst = eb + eoff[t] + cpart * ED_EC;
#pragma unroll
                            for (int j = 0; j < ED_EC / 4; j++) {
                                float4 v;
                                v.x = __uint_as_float(r[4 * j + 0]) + be[j].x; v.y = __uint_as_float(r[4 * j + 1]) + be[j].y;
                                v.z = __uint_as_float(r[4 * j + 2]) + be[j].z; v.w = __uint_as_float(r[4 * j + 3]) + be[j].w;
                                if (hsw) {
                                    v.x *= __saturatef(fmaf(v.x, 1.f / 6.f, 0.5f)); v.y *= __saturatef(fmaf(v.y, 1.f / 6.f, 0.5f));
                                    v.z *= __saturatef(fmaf(v.z, 1.f / 6.f, 0.5f)); v.w *= __saturatef(fmaf(v.w, 1.f / 6.f, 0.5f));
                                } else {
                                    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                                }
                                *reinterpret_cast<float4*>(dst + 4 * j) = v;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(a));
                if (cw == 0 && lane == 0) ED_TRACE(0, tev + 3);
                asm volatile("bar.sync 1, %0;" ::"n"(ED_CTHREADS) : "memory");        // the chunk's tile is complete
                if (cw == 0 && lane == 0) ED_TRACE(0, tev + 4);
                // ---- depthwise conv on the tile: lane = channel, warp = (image, row segment, run of XR columns) work units
                const bool active = c * ED_CW + lane < p.C;
                for (int idx = cw; idx < nwork; idx += ED_CWARPS) {
                    const int g = idx / per_img, rem = idx - g * per_img, yseg = rem / runs, run = rem - yseg * runs;
                    const int y0 = yseg * p.YS, nrows = min(p.YS, it.th - y0);
                    const int ox0 = run * XR, nvalid = active ? min(XR, p.Wo - ox0) : 0;
                    const float* rp = eb + g * img_floats + (y0 * S) * rowstride + (ox0 * S) * ED_ES + lane;   // tap (0,0) of output (y0, ox0)
                    // (lanes past the last channel of a partial chunk run on zero weights with nvalid = 0: they store nothing)
                    float* op = p.out + (((size_t)(it.b0 + g) * p.Ho + it.oy0 + y0) * p.Wo + ox0) * p.C + c * ED_CW + lane;
                    const int opitch = p.Wo * p.C;
                    float* rsum = rowp + ((g * p.TH + y0) * runs + run) * 32 + lane;      // this unit's (row, run) squeeze partials
                    // rolling window of KS input rows in registers (slot = input row mod KS, segment-relative): one output row costs
                    // S * SEG shared-memory loads against XR * k*k FMAs; the other three warps of the sub-partition cover the latency
                    float win[KS][SEG];
#pragma unroll
                    for (int r = 0; r < KS - S; r++)
#pragma unroll
                        for (int e = 0; e < SEG; e++) win[r][e] = rp[r * rowstride + e * ED_ES];
                    for (int base_r = 0; base_r < nrows; base_r += KS) {
#pragma unroll
                        for (int r = 0; r < KS; r++) {          // output row i = base_r + r; base_r * S is a multiple of KS, so window slots are static
                            const int i = base_r + r;
                            if (i < nrows) {
#pragma unroll
                                for (int t = 0; t < S; t++) {
                                    const float* lp = rp + (i * S + KS - S + t) * rowstride;
#pragma unroll
                                    for (int e = 0; e < SEG; e++) win[(r * S + KS - S + t) % KS][e] = lp[e * ED_ES];
                                }
                                float acc[XR];
#pragma unroll
                                for (int j = 0; j < XR; j++) acc[j] = bdw;
#pragma unroll
                                for (int ky = 0; ky < KS; ky++)
#pragma unroll
                                    for (int kx = 0; kx < KS; kx++)
#pragma unroll
                                        for (int j = 0; j < XR; j++) acc[j] = fmaf(win[(r * S + ky) % KS][j * S + kx], w[ky * KS + kx], acc[j]);
                                // all XR chains are complete HERE: without this the compiler sinks each k*k-FMA chain into its own
                                // `j < nvalid` branch, where it runs as one dependent chain (4 cycles per FMA instead of 1)
#pragma unroll
                                for (int j = 0; j < XR; j++) asm volatile("" : "+f"(acc[j]));
                                float* orow = op + (size_t)i * opitch;
                                float rs = 0.f;
#pragma unroll
                                for (int j = 0; j < XR; j++) {
                                    const float x = acc[j];
                                    // hardswish x * relu6(x + 3) / 6 as x * sat(x / 6 + 0.5): one FFMA.SAT + one FMUL
                                    const float v = hsw ? x * __saturatef(fmaf(x, 1.f / 6.f, 0.5f)) : fmaxf(x, 0.f);
                                    if (j < nvalid) { orow[j * p.C] = v; rs += v; }
                                }
                                rsum[i * runs * 32] = rs;
                            }
                        }
                    }
                }
                if (cw == 0 && lane == 0) ED_TRACE(0, tev + 5);
                // every warp is done with the tile (a single tile buffer may now be overwritten) and the row partials are complete
                asm volatile("bar.sync 1, %0;" ::"n"(ED_CTHREADS) : "memory");
                finish_pool(c, it, runs);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
// Operand row width for a K-channel input: 64-byte rows (16 floats) or 128-byte rows (32 floats), whichever pads K less.
void ed_operand_shape(int K, int* RB, int* slabs) {
    const int s64 = (K + 15) / 16, s128 = (K + 31) / 32;
    if (s64 * 64 < s128 * 128) { *RB = 64; *slabs = s64; } else { *RB = 128; *slabs = s128; }
}

struct EdCfg { int G, TH, nbuf, bstages, YS; bool ok; };

static size_t ed_smem_bytes(const EdParams& p) { return (size_t)p.off_bar + 8 * (10 + 2 * ED_MAXBST) + 64 + 1024; }

// fills the geometry-dependent fields of p for (G, TH, nbuf, bstages); returns false if the item does not fit the SM
static bool ed_layout(EdParams& p, int KS, int S, int G, int TH, int nbuf, int bstages) {
    const int pad = (KS - 1) / 2, KK = KS * KS;
    p.G = G; p.TH = TH; p.bands = (p.Ho + TH - 1) / TH; p.nbuf = nbuf; p.bstages = bstages;
    if (G > 1 && p.bands != 1) return false;
    int Pmax;
    if (p.bands == 1) Pmax = G * p.H * p.W;
    else {
        Pmax = 0;
        for (int b = 0; b < p.bands; b++) {
            const int oy0 = b * TH, th = std::min(TH, p.Ho - oy0);
            const int lo = std::max(oy0 * S - pad, 0), hi = std::min((oy0 + th - 1) * S - pad + KS - 1, p.H - 1);
            Pmax = std::max(Pmax, (hi - lo + 1) * p.W);
        }
    }
    p.T = (Pmax + 127) / 128;
    if (p.T > ED_MAXT || G > ED_MAXG) return false;
    p.tail_rows = ((Pmax - 128 * (p.T - 1)) + 7) & ~7;
    p.P8 = 128 * (p.T - 1) + p.tail_rows;
    p.part_bytes = p.slabs * p.P8 * p.RB;
    p.in_h = (TH - 1) * S + KS; p.Wp = p.W + 2 * pad;
    p.ebuf_floats = (G * p.in_h * p.Wp * ED_ES + 3) & ~3;
    p.b_part_bytes = p.slabs * ED_CW * p.RB;
    p.par_bytes = (2 + KK) * ED_CW * 4;
    p.bstage_bytes = (2 * p.b_part_bytes + p.par_bytes + 1023) & ~1023;
    p.off_b = (3 * p.part_bytes + 1023) & ~1023;
    p.off_e = p.off_b + bstages * p.bstage_bytes;
    const int slack = 8 * ED_ES * 4 + 64;                    // a partial last run reads a few pixels past its row
    p.off_red = (p.off_e + nbuf * p.ebuf_floats * 4 + slack + 15) & ~15;
    const int runs = (p.Wo + ed_xr(KS, S) - 1) / ed_xr(KS, S);
    p.off_bar = (p.off_red + G * std::min(TH, p.Ho) * runs * 32 * 4 + 15) & ~15;
    int cols = 32; while (cols < 2 * p.T * ED_CW) cols <<= 1;
    p.tmem_cols = cols;
    return cols <= 512 && ed_smem_bytes(p) <= 227 * 1024;
}

static EdCfg choose_ed_cfg(EdParams p, int KS, int S, int sm_count) {
    EdCfg best{0, 0, 0, 0, 0, false};
    double best_cost = 1e300;
    const int KK = KS * KS, XR = ed_xr(KS, S);
    auto consider = [&](int G, int TH) {
        for (int nbuf = 2; nbuf >= 1; nbuf--)
            for (int bst = 3; bst >= 2; bst--) {
                if (!ed_layout(p, KS, S, G, TH, nbuf, bst)) continue;
                // rows of the expanded tensor computed per image (halo rows of neighbouring bands are recomputed)
                double rows = 0;
                const int pad = (KS - 1) / 2;
                for (int b = 0; b < p.bands; b++) {
                    const int oy0 = b * TH, th = std::min(TH, p.Ho - oy0);
                    rows += std::min((oy0 + th - 1) * S - pad + KS - 1, p.H - 1) - std::max(oy0 * S - pad, 0) + 1;
                }
                if (rows > 1.6 * p.H) return;
                const double items = (double)((p.B + G - 1) / G) * p.bands;
                const double waves = std::ceil(items / sm_count);
                const double epi = (rows / p.bands) * p.W * G * p.nchunks * 32 * 12.0;
                const double dw = (double)G * std::min(TH, p.Ho) * p.Wo * p.nchunks * 32 * (KK + 12.0);
                const double fixed = p.nchunks * 256.0 * (nbuf == 2 ? 250.0 : 400.0) + 256.0 * 4000.0;
                const double cost = waves * (epi + dw + fixed);
                if (cost < best_cost) {
                    best_cost = cost;
                    // rows per work unit: fill the compute warps evenly
                    const int runs = (p.Wo + XR - 1) / XR, th = std::min(TH, p.Ho);
                    int ys = th;
                    double best_u = 1e300;
                    for (int YS = 1; YS <= th; YS++) {
                        const int units = G * ((th + YS - 1) / YS) * runs;
                        const double u = std::ceil(units / (double)ED_CWARPS) * (YS + 0.4 * (KS - S)) + 0.01 * YS;
                        if (u < best_u) { best_u = u; ys = YS; }
                    }
                    best = EdCfg{G, TH, nbuf, bst, ys, true};
                }
                return;                                       // the roomiest buffering that fits this (G, TH)
            }
    };
    for (int G = 1; G <= ED_MAXG; G++) consider(G, p.Ho);
    for (int TH = p.Ho - 1; TH >= 1; TH--) consider(1, TH);
    return best;
}

// Fused expand + depthwise of block m on `in` [B][H][H][cin] -> `out` [B][Ho][Ho][exp]; squeeze partials as depthwise_tma.
// Returns KPD_ERR_UNSUPPORTED (without an error message being fatal) when the layer does not fit: the caller falls back to
// the separate expand GEMM + depthwise kernels.
int expand_depthwise_fused(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts,
                           int B, int H, int sm_count, cudaStream_t s) {
    if (!m.wed_exp || !m.ed_params) return KPD_ERR_UNSUPPORTED;
    const int S = m.stride, KS = m.k;
    if ((KS != 3 && KS != 5) || (S != 1 && S != 2) || (m.act != ACT_RELU && m.act != ACT_HSWISH) || (m.cin & 3) || (m.exp & 3)) return KPD_ERR_UNSUPPORTED;
    if ((int64_t)B * H * H >= ((int64_t)1 << 31) - 1024 || (int64_t)(H / S) * (H / S) * m.exp >= ((int64_t)1 << 31)) return KPD_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    EdParams p{};
    p.out = out; p.pool = pool; p.params = m.ed_params;
    p.B = B; p.H = H; p.W = H; p.Ho = H / S; p.Wo = H / S; p.C = m.exp; p.K = m.cin; p.act = m.act;
    p.nchunks = (m.exp + ED_CW - 1) / ED_CW;
    ed_operand_shape(m.cin, &p.RB, &p.slabs);
    p.ksteps = (m.cin + 7) / 8;
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int, int>, EdCfg> cache;
    EdCfg cfg;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_tuple(B, H, m.cin, m.exp, KS, S);
        auto it = cache.find(key);
        if (it == cache.end()) {
            it = cache.emplace(key, choose_ed_cfg(p, KS, S, sm_count)).first;
            if (getenv("KPD_ED_DEBUG")) {
                EdParams d = p; const EdCfg& c = it->second;
                if (c.ok) ed_layout(d, KS, S, c.G, c.TH, c.nbuf, c.bstages);
                fprintf(stderr, "kpd ed: B=%d H=%d K=%d C=%d k=%d s=%d -> %s G=%d TH=%d bands=%d T=%d RB=%d slabs=%d nbuf=%d bst=%d YS=%d smem=%zu tmem=%d\n", B, H, m.cin, m.exp, KS, S,
                        c.ok ? "fused" : "unfused", c.G, c.TH, d.bands, d.T, d.RB, d.slabs, c.nbuf, c.bstages, c.YS, c.ok ? ed_smem_bytes(d) : (size_t)0, d.tmem_cols);
            }
        }
        cfg = it->second;
    }
    if (!cfg.ok || !ed_layout(p, KS, S, cfg.G, cfg.TH, cfg.nbuf, cfg.bstages)) return KPD_ERR_UNSUPPORTED;
    p.YS = cfg.YS;
    if (pool_parts) *pool_parts = p.Ho;                      // one squeeze partial per output row (canonical summation order)
    if (pool && (size_t)p.Ho * m.exp > pool_capacity) { set_error("expand_depthwise_fused: SE partial buffer too small"); return KPD_ERR_STATE; }
    const int64_t items = (int64_t)((B + p.G - 1) / p.G) * p.bands;
    if (items >= ((int64_t)1 << 31)) return KPD_ERR_UNSUPPORTED;
    p.total_items = (int)items;

    CUtensorMap tmA, tmAt, tmB;
    const CUtensorMapSwizzle sw = p.RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int which = 0; which < 2; which++) {
        cuuint64_t dims[2] = {(cuuint64_t)m.cin, (cuuint64_t)B * H * H};
        cuuint64_t strides[1] = {(cuuint64_t)m.cin * 4};
        cuuint32_t box[2] = {(cuuint32_t)(p.RB / 4), (cuuint32_t)(which ? p.tail_rows : 128)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(which ? &tmAt : &tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(in), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("expand_depthwise_fused: cuTensorMapEncodeTiled(input) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        const int Kw = p.slabs * (p.RB / 4);
        cuuint64_t dims[2] = {(cuuint64_t)Kw, (cuuint64_t)(2 * p.nchunks * ED_CW)};
        cuuint64_t strides[1] = {(cuuint64_t)Kw * 4};
        cuuint32_t box[2] = {(cuuint32_t)(p.RB / 4), (cuuint32_t)ED_CW};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(m.wed_exp), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("expand_depthwise_fused: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    const size_t smem = ed_smem_bytes(p);
    const int grid = std::min(p.total_items, sm_count);
    static const bool trace_on = getenv("KPD_ED_TRACE") != nullptr;
    static long long* trace_buf = nullptr;
    if (trace_on) {
        if (!trace_buf) cudaMalloc(&trace_buf, 4 * 256 * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 4 * 256 * sizeof(long long), s);
        p.trace = trace_buf;
    }
    static uint64_t attr[4] = {};
    if ((KS == 3) == (m.act == ACT_HSWISH)) return KPD_ERR_UNSUPPORTED;
#define KPD_ED_LAUNCH(KK_, SS_, HS_, SLOT)                                                   \
    do {                                                                                     \
        KPD_CUDA(ensure_smem_attr(ed_kernel<KK_, SS_, HS_>, 227 * 1024, attr[SLOT]));        \
        ed_kernel<KK_, SS_, HS_><<<grid, ED_THREADS, smem, s>>>(tmA, tmAt, tmB, p);          \
    } while (0)
    // MobileNetV3 pairs 3x3 depthwise convs with ReLU and 5x5 ones with hardswish (tv:models/mobilenetv3.py:246-258): only
    // those combinations are instantiated; anything else takes the layer-by-layer kernels
    const bool hs = m.act == ACT_HSWISH;
    if (KS == 3 && S == 1 && !hs) KPD_ED_LAUNCH(3, 1, false, 0);
    else if (KS == 3 && S == 2 && !hs) KPD_ED_LAUNCH(3, 2, false, 1);
    else if (KS == 5 && S == 1 && hs) KPD_ED_LAUNCH(5, 1, true, 2);
    else if (KS == 5 && S == 2 && hs) KPD_ED_LAUNCH(5, 2, true, 3);
    else return KPD_ERR_UNSUPPORTED;
#undef KPD_ED_LAUNCH
    KPD_LAUNCH_CHECK();
    if (trace_on) {                                          // debugging only: synchronous dump of CTA 0's timeline (SM clocks)
        std::vector<long long> h(4 * 256);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        long long t0 = 0; for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        const char* roles[4] = {"compute", "mma", "split", "producer"};
        fprintf(stderr, "kpd ed trace: K=%d C=%d k=%d s=%d H=%d items=%d nchunks=%d T=%d\n", m.cin, m.exp, KS, S, H, p.total_items, p.nchunks, p.T);
        for (int r = 0; r < 4; r++) {
            fprintf(stderr, "  %s:", roles[r]);
            for (int e = 0; e < 100; e++) if (h[r * 256 + e]) fprintf(stderr, " %d:%lld", e, h[r * 256 + e] - t0);
            fprintf(stderr, "\n");
        }
    }
    return 0;
}

}  // namespace kpd
