// Batch-wide depthwise k x k convolution (+ folded BN + activation, optional SE squeeze), NHWC fp32, with TMA-staged tiles.
//   reference op: tv:models/mobilenetv3.py:92-104 (the depthwise Conv2dNormActivation of every InvertedResidual block)
//   and the AdaptiveAvgPool2d of its SqueezeExcitation (tv:ops/misc.py:243-251).
//
// Work item ("tile") = (image, group of CG channels, TH x TW output tile).  Persistent CTAs walk the tiles with a ring of
// shared-memory stages: ONE 4-D TMA box {CG channels, in_w, in_h, 1 image} brings the zero-padded input tile (halo
// included -- out-of-image coordinates are zero-filled by TMA, which implements the padding) and one 2-D box brings the
// group's k*k tap weights; both complete on the stage's mbarrier.  The loads of the next tiles are in flight while the
// 256 threads compute the current one: no index arithmetic, register staging or exposed DRAM latency on the load side.
//
// Compute: a lane owns ONE channel (a warp = 32 consecutive channels of one pixel run, so every shared-memory access is
// a conflict-free 128-byte row and every global store a full 128-byte line).  Its k*k weights live in registers for the
// whole tile, and it marches down the rows of a (y-segment, run of XR output columns) keeping a k-row register window of
// the input: one output row costs (XR-1)*stride+k LDS.32 per new input row against XR*k*k FMAs, i.e. the kernel issues
// ~8 FMAs per shared-memory load (the float4-per-thread predecessors, dw_tiled_kernel in kpd_pw.cu and the first TMA
// version, needed a 16-byte LDS per ~6 FMAs and were LDS/issue-bound at <= 45% issue utilisation).
#include "kpd_tma.cuh"
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>

namespace kpd {

struct DwArgs {
    const float* bias; float* out; float* pool;
    int B, H, Ho, C, act;
    int TH, TW, YS, tiles_x, tiles, CG, groups, in_h, in_w;
    int stages, stage_bytes, w_off, tx_bytes;   // w_off = weight offset in the stage (128-aligned); tx_bytes = bytes the two boxes deliver
    int total;
};

constexpr int DW_THREADS = 256, DW_MAXCG = 128, DW_SLOTS = 16;
// output columns per run: 4, or 2 for the 5x5 stride-2 layers (their 5 x 11 window would not fit the 128-register budget)
__host__ __device__ constexpr int dw_xr(int K, int S) { return (K == 5 && S == 2) ? 2 : 4; }

template <int K, int S, int CG>
__global__ void __launch_bounds__(DW_THREADS, 2) dw_tma_kernel(const __grid_constant__ CUtensorMap tmIn,
                                                               const __grid_constant__ CUtensorMap tmW, const DwArgs a) {
    constexpr int XR = dw_xr(K, S), SEG = (XR - 1) * S + K, pad = (K - 1) / 2;
    extern __shared__ __align__(128) uint8_t dw_smem[];       // indexed directly (no generic-pointer round trip) so the loads compile to LDS
    uint8_t* const base = dw_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)a.stages * a.stage_bytes);
    float* red = reinterpret_cast<float*>(bars + 8);           // [2][DW_SLOTS][DW_MAXCG] SE squeeze partials, double-buffered by tile parity
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int step = gridDim.x;
    const bool hsw = a.act == ACT_HSWISH;
    const int opitch = a.Ho * a.C;                             // floats between output rows (< 2^31 per image, host-checked)

    auto issue = [&](int item, int s) {                        // one thread: both TMA loads of tile `item` into stage s
        const uint32_t grp = (uint32_t)item % (uint32_t)a.groups, rest = (uint32_t)item / (uint32_t)a.groups;   // channel group fastest:
        const uint32_t tile = rest % (uint32_t)a.tiles, b = rest / (uint32_t)a.tiles;   // CTAs running side by side share the tile's DRAM lines
        const int ty0 = (int)(tile / (uint32_t)a.tiles_x) * a.TH, tx0 = (int)(tile % (uint32_t)a.tiles_x) * a.TW;
        const uint32_t bar = smem_u32(bars + s), dst = smem_u32(base + (size_t)s * a.stage_bytes);
        mbar_expect_tx(bar, (uint32_t)a.tx_bytes);
        tma_load_4d(dst, &tmIn, bar, (int)grp * CG, tx0 * S - pad, ty0 * S - pad, (int)b);
        tma_load_2d(dst + a.w_off, &tmW, bar, (int)grp * CG, 0);
    };

    if (tid == 0) {
        if (smem_u32(base) & 127u) { printf("kpd: depthwise shared-memory window is not 128-byte aligned\n"); __trap(); }
        for (int s = 0; s < a.stages; s++) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < a.stages; s++) { const int item = blockIdx.x + s * step; if (item < a.total) issue(item, s); }
    }

    // lane -> channel: groups of <= 16 channels pack two (y-segment, run) slots into a warp
    constexpr int lpi = CG <= 16 ? 16 : 32, spw = 32 / lpi;        // lanes per slot, slots per warp
    constexpr int nsub = (CG + 31) >> 5;                           // 32-channel sub-blocks of the group (1, 2 or 4: divides the slot stride)
    const int ch_lane = lane & (lpi - 1);
    constexpr int nslots = (DW_THREADS / 32) * spw;
    const int slot0 = warp * spw + (lane / lpi);
    const int sub = slot0 % nsub, ch = sub * 32 + ch_lane;     // channel within the group: fixed for the thread (nslots % nsub == 0)

    int n = 0;
    for (int item = blockIdx.x; item < a.total; item += step, n++) {
        const int s = n % a.stages;
        const uint32_t phase = (uint32_t)(n / a.stages) & 1u;
        const uint32_t grp = (uint32_t)item % (uint32_t)a.groups, rest = (uint32_t)item / (uint32_t)a.groups;
        const uint32_t tile = rest % (uint32_t)a.tiles, b = rest / (uint32_t)a.tiles;
        const int c0 = (int)grp * CG, cg = min(CG, a.C - c0);
        const int ty0 = (int)(tile / (uint32_t)a.tiles_x) * a.TH, tx0 = (int)(tile % (uint32_t)a.tiles_x) * a.TW;
        const int th = min(a.TH, a.Ho - ty0), tw = min(a.TW, a.Ho - tx0);
        const int runs = (tw + XR - 1) / XR, ysegs = (th + a.YS - 1) / a.YS;
        const int nwork = ysegs * runs * nsub;
        const bool active = ch < cg;
        const float* tilep = reinterpret_cast<const float*>(base + (size_t)s * a.stage_bytes) + ch;            // [in_h][in_w][CG]
        const float* wsm = reinterpret_cast<const float*>(base + (size_t)s * a.stage_bytes + a.w_off) + ch;    // [K*K][CG]
        const float bb = active ? __ldg(a.bias + c0 + ch) : 0.f;
        float* outb = a.out + ((size_t)b * a.Ho * a.Ho) * a.C + c0 + ch;
        const int rowstride = a.in_w * CG;                     // floats between input rows of the tile
        float ssum = 0.f;

        mbar_wait(smem_u32(bars + s), phase);

        float w[K * K];
#pragma unroll
        for (int t = 0; t < K * K; t++) w[t] = wsm[t * CG];

        for (int idx = slot0; idx < nwork; idx += nslots) {    // idx = (yseg * runs + run) * nsub + sub
            const int q = idx / nsub, yseg = q / runs, run = q - yseg * runs;
            const int y0 = yseg * a.YS, nrows = min(a.YS, th - y0);
            const int ox0 = run * XR, nvalid = min(XR, tw - ox0);
            const float* rp = tilep + (size_t)(y0 * S) * rowstride + (ox0 * S) * CG;     // input pixel of tap (0,0) of output (y0, ox0)
            float* op = outb + ((size_t)(ty0 + y0) * a.Ho + tx0 + ox0) * a.C;
            float win[K][SEG];                                  // rolling window of K input rows (slot = input row mod K, segment-relative)
#pragma unroll
            for (int r = 0; r < K - S; r++)
#pragma unroll
                for (int e = 0; e < SEG; e++) win[r][e] = rp[(size_t)r * rowstride + e * CG];
            for (int base_r = 0; base_r < nrows; base_r += K) {
#pragma unroll
                for (int r = 0; r < K; r++) {                   // output row i = base_r + r; base_r * S is a multiple of K, so slots are static
                    const int i = base_r + r;
                    if (i < nrows) {
#pragma unroll
                        for (int t = 0; t < S; t++) {
                            const float* lp = rp + (size_t)(i * S + K - S + t) * rowstride;
#pragma unroll
                            for (int e = 0; e < SEG; e++) win[(r * S + K - S + t) % K][e] = lp[e * CG];
                        }
                        float acc[XR];
#pragma unroll
                        for (int j = 0; j < XR; j++) acc[j] = bb;
#pragma unroll
                        for (int ky = 0; ky < K; ky++)
#pragma unroll
                            for (int kx = 0; kx < K; kx++)
#pragma unroll
                                for (int j = 0; j < XR; j++) acc[j] = fmaf(win[(r * S + ky) % K][j * S + kx], w[ky * K + kx], acc[j]);
                        // all XR chains are complete HERE: without this the compiler sinks each k*k-FMA chain into its own
                        // `j < nvalid` branch, where it runs as one dependent chain (4 cycles per FMA instead of 1)
#pragma unroll
                        for (int j = 0; j < XR; j++) asm volatile("" : "+f"(acc[j]));
                        if (active) {
                            float* orow = op + i * opitch;
#pragma unroll
                            for (int j = 0; j < XR; j++) {
                                if (j < nvalid) {
                                    const float x = acc[j];                      // branch-free: the trunk only uses ReLU and hardswish here
                                    const float v = hsw ? x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f) : fmaxf(x, 0.f);
                                    orow[j * a.C] = v;
                                    ssum += v;
                                }
                            }
                        }
                    }
                }
            }
        }
        // SE squeeze, deterministic: every thread parks its sum at [slot][channel]; after the barrier one thread per channel
        // adds the slots that own that channel in a fixed order and stores the tile's partial (summed over tiles by se_fc)
        float* redb = red + (n & 1) * (DW_SLOTS * DW_MAXCG);
        if (a.pool && active) redb[slot0 * DW_MAXCG + ch] = ssum;
        __syncthreads();                                      // stage s is free again; this tile's squeeze partials are complete
        if (a.pool && tid < cg) {
            float tot = 0.f;
            for (int sl = (lpi == 32 ? (tid >> 5) : 0); sl < nslots; sl += nsub) tot += redb[sl * DW_MAXCG + tid];
            a.pool[((size_t)b * a.tiles + tile) * a.C + c0 + tid] = tot;
        }
        if (tid == 0) { const int nxt = item + a.stages * step; if (nxt < a.total) issue(nxt, s); }
    }
}

// Tile shape per layer: channel group CG (<= 32, or 64 / 128 for the small late maps so that all 8 warps have work), output
// tile TH x TW, rows per y-segment YS.  Cost model = (tiles per image) x (passes over the tile's (y-segment, run, sub-block)
// work list x rows per segment + a fixed per-tile term for the barrier hand-over), subject to the stage fitting the budget.
struct DwCfg { int CG, TH, TW, YS; };

static DwCfg choose_dw_cfg(int C, int Ho, int K, int S, size_t budget) {
    DwCfg best{0, 0, 0, 0};
    double best_cost = 1e30;
    const int cgs[6] = {16, 24, 32, 64, 128, 0};
    for (int ci = 0; cgs[ci]; ci++) {
        int CG = cgs[ci];
        if (CG > 32 && CG > C) continue;
        if (CG > C && !(CG == 16 && C < 16) && !(CG > C && CG - C < 8)) continue;   // a group wider than C only to cover a small C
        const int groups = (C + CG - 1) / CG;
        const int lpi = CG <= 16 ? 16 : 32, slots = 8 * (32 / lpi), nsub = (CG + 31) / 32;
        for (int TW : {std::min(Ho, 28), std::min(Ho, 56)}) {
            const int XR = dw_xr(K, S), runs = (TW + XR - 1) / XR, tiles_x = (Ho + TW - 1) / TW;
            for (int TH = 1; TH <= std::min(Ho, 56); TH++) {
                const size_t bytes = (size_t)((TH - 1) * S + K) * ((TW - 1) * S + K) * CG * 4;
                if (bytes > budget) break;
                const int tiles_y = (Ho + TH - 1) / TH;
                if ((tiles_y - 1) * TH >= Ho) continue;
                for (int YS = 1; YS <= TH; YS++) {
                    const int ysegs = (TH + YS - 1) / YS;
                    const int passes = (ysegs * runs * nsub + slots - 1) / slots;
                    const double per_tile = passes * (YS + 0.4 * (K - S)) + 3.0;
                    const double cost = (double)groups * tiles_x * tiles_y * per_tile;
                    if (cost < best_cost - 1e-9) { best_cost = cost; best = DwCfg{CG, TH, TW, YS}; }
                }
            }
        }
    }
    return best;
}

// `pool` (SE squeeze, may be null): [B][*pool_parts][C] per-tile partial sums, every entry written exactly once (no atomics,
// no memset); the caller must provide pool_capacity >= *pool_parts * C floats per image.
int depthwise_tma(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts, int B, int H, int sm_count, cudaStream_t s) {
    const int S = m.stride, Ho = H / S, C = m.exp, K = m.k;
    if ((C & 3) || (K != 3 && K != 5) || (S != 1 && S != 2) || (m.act != ACT_RELU && m.act != ACT_HSWISH) || (int64_t)Ho * Ho * C >= ((int64_t)1 << 31)) { set_error("depthwise_tma: unsupported layer C=%d k=%d stride=%d", C, K, S); return KPD_ERR_UNSUPPORTED; }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return KPD_ERR_STATE; }
    constexpr int tile_kb = 48, stages_env = 2;             // tile budget per stage (KB) and ring depth: the measured best
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, DwCfg> cache;
    DwCfg cfg;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_tuple(C, Ho, K, S);
        auto it = cache.find(key);
        if (it == cache.end()) it = cache.emplace(key, choose_dw_cfg(C, Ho, K, S, (size_t)tile_kb * 1024)).first;
        cfg = it->second;
    }
    if (cfg.CG == 0) { set_error("depthwise_tma: no tile shape fits C=%d Ho=%d k=%d", C, Ho, K); return KPD_ERR_UNSUPPORTED; }
    const int CG = cfg.CG, TH = cfg.TH, TW = cfg.TW;
    DwArgs a{};
    a.bias = m.b_dw; a.out = out; a.pool = pool; a.B = B; a.H = H; a.Ho = Ho; a.C = C; a.act = m.act;
    a.TH = TH; a.TW = TW; a.YS = cfg.YS; a.tiles_x = (Ho + TW - 1) / TW; a.tiles = a.tiles_x * ((Ho + TH - 1) / TH); a.CG = CG;
    a.groups = (C + CG - 1) / CG;
    if (pool_parts) *pool_parts = a.tiles;
    if (pool && (size_t)a.tiles * C > pool_capacity) { set_error("depthwise_tma: SE partial buffer too small (%d tiles x %d channels)", a.tiles, C); return KPD_ERR_STATE; }
    a.in_h = (TH - 1) * S + K; a.in_w = (TW - 1) * S + K;
    const int tile_bytes = a.in_h * a.in_w * CG * 4, w_bytes = K * K * CG * 4;
    if (a.in_h > 256 || a.in_w > 256 || CG > DW_MAXCG) { set_error("depthwise_tma: unsupported tile %dx%dx%d", a.in_h, a.in_w, CG); return KPD_ERR_UNSUPPORTED; }
    a.tx_bytes = tile_bytes + w_bytes;
    // a partial last run reads up to XR*S pixels past the tile row and idle lanes up to 31 channels past their group:
    // keep both inside the stage (the values are never used)
    a.w_off = (tile_bytes + dw_xr(K, S) * S * CG * 4 + 128 + 127) & ~127;
    a.stage_bytes = (a.w_off + w_bytes + 128 + 127) & ~127;
    a.stages = std::max(2, std::min(stages_env, 8));
    const int64_t total = (int64_t)a.tiles * B * a.groups;
    if (total >= ((int64_t)1 << 31)) { set_error("depthwise_tma: too many work items"); return KPD_ERR_UNSUPPORTED; }
    a.total = (int)total;
    const size_t smem = (size_t)a.stages * a.stage_bytes + 128 /*align*/ + 64 /*barriers*/ + 2 * DW_SLOTS * DW_MAXCG * 4 /*red*/;
    if (smem > 227 * 1024) { set_error("depthwise_tma: tile does not fit in shared memory"); return KPD_ERR_UNSUPPORTED; }
    const int ctas_per_sm = std::max(1, std::min(2, (int)((227 * 1024) / (smem + 1024))));

    CUtensorMap tmIn, tmW;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)H * C * 4, (cuuint64_t)H * H * C * 4};
        cuuint32_t box[4] = {(cuuint32_t)CG, (cuuint32_t)a.in_w, (cuuint32_t)a.in_h, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CG * 4 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("depthwise_tma: cuTensorMapEncodeTiled(input) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)(K * K)};
        cuuint64_t strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t box[2] = {(cuuint32_t)CG, (cuuint32_t)(K * K)};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(m.w_dw), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("depthwise_tma: cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return KPD_ERR_STATE; }
    }
    const int grid = (int)std::min<int64_t>(total, (int64_t)sm_count * ctas_per_sm);
    static uint64_t attr[20] = {};
#define KPD_DW_LAUNCH(KK, SS, CC, SLOT)                                                          \
    do {                                                                                         \
        KPD_CUDA(ensure_smem_attr(dw_tma_kernel<KK, SS, CC>, 227 * 1024, attr[SLOT]));           \
        dw_tma_kernel<KK, SS, CC><<<grid, DW_THREADS, smem, s>>>(tmIn, tmW, a);                  \
    } while (0)
#define KPD_DW_CG(KK, SS, BASE)                                                                  \
    do {                                                                                         \
        switch (CG) {                                                                            \
            case 16: KPD_DW_LAUNCH(KK, SS, 16, BASE + 0); break;                                 \
            case 24: KPD_DW_LAUNCH(KK, SS, 24, BASE + 1); break;                                 \
            case 32: KPD_DW_LAUNCH(KK, SS, 32, BASE + 2); break;                                 \
            case 64: KPD_DW_LAUNCH(KK, SS, 64, BASE + 3); break;                                 \
            case 128: KPD_DW_LAUNCH(KK, SS, 128, BASE + 4); break;                               \
            default: set_error("depthwise_tma: unsupported channel group %d", CG); return KPD_ERR_UNSUPPORTED; \
        }                                                                                        \
    } while (0)
    if (K == 3 && S == 1) KPD_DW_CG(3, 1, 0);
    else if (K == 3) KPD_DW_CG(3, 2, 5);
    else if (S == 1) KPD_DW_CG(5, 1, 10);
    else KPD_DW_CG(5, 2, 15);
#undef KPD_DW_CG
#undef KPD_DW_LAUNCH
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
