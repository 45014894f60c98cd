// Batch-wide pointwise (1x1) convolution as an fp32 CUDA-core GEMM: out[M][N] = act(A[M][K] * W[K][N] + bias) (+ residual).
//   reference ops: every 1x1 conv of the trunk -- tv:models/mobilenetv3.py:87-108 (expand / project), features.12 -- and
//   dll/models/backbone.py:33-37 (FPN laterals 1..3 + nearest top-down add).
//
// Why CUDA cores: on sm_100a the legacy mma.sync path issues about one m16n8k8 TF32 MMA per 14 clk per SM (measured),
// i.e. ~73 MAC/clk/SM, and the 3xTF32 split this trunk needs for fp32-level accuracy divides that by three; 128 FFMA
// lanes deliver 128 MAC/clk/SM in exact fp32.  (tcgen05 would need the hi/lo operand split materialised in shared memory;
// these GEMMs have K <= 576 and are bound by their activation streams once the math runs at FFMA rate.)
//
// Structure: persistent CTAs walk (128-row tile, column chunk) items; A rows (K-contiguous NHWC pixels) and the matching
// W[k-chunk][column chunk] stream through a 3-stage cp.async ring that runs ahead across item boundaries.  A lane owns R
// rows (lane, lane+32, ...), a warp owns an NT-wide column strip: the A operand is one conflict-free LDS.128 per row per 4 k
// (row pitch 36 floats), the W operand a shared-memory broadcast, so 4*R*NT FMAs are issued per R + NT shared loads.  (R, NT)
// is picked per layer so that narrow outputs (N = 16 ... 96) still occupy all 8 warps.  The SE excitation is applied to the
// A tile in shared memory by the thread that copied it; bias / activation / residual (same-shape or nearest-upsampled) /
// tf32 rounding are fused in the epilogue.
#include "kpd_common.cuh"
#include <stdlib.h>
#include <cmath>

namespace kpd {

constexpr int PF_BM = 128, PF_KC = 32, PF_LDA = PF_KC + 4, PF_STAGES = 3, PF_THREADS = 256, PF_MAXCOVER = 128;

__device__ __forceinline__ void pf_cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;                       // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}

struct ActNone  { __device__ __forceinline__ float operator()(float x) const { return x; } };
struct ActRelu  { __device__ __forceinline__ float operator()(float x) const { return fmaxf(x, 0.f); } };
struct ActHswish { __device__ __forceinline__ float operator()(float x) const { return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f); } };

// R rows per lane, NT columns per warp strip; WS = 2R strips side by side, RG = 4/R row groups of 32R rows.
template <int R, int NT>
__global__ void __launch_bounds__(PF_THREADS, (R * NT <= 24) ? 3 : 2) pw_ffma_kernel(const PwMma p, int tiles_m, int nchunks, int chunk_cols) {
    constexpr int WS = 2 * R, COVER = WS * NT, STAGE_A = PF_BM * PF_LDA, STAGE_W = PF_KC * COVER;
    constexpr int WV = COVER / 4;                          // float4 per staged W row
    extern __shared__ __align__(16) float pf_smem[];       // [PF_STAGES][STAGE_A + STAGE_W]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strip = warp % WS, rg = warp / WS;
    const int K = p.K, N = p.N;
    const uint32_t M = (uint32_t)p.M;
    const int sk4 = tid & 7, srow = tid >> 3;              // A staging map: float4 (row = srow + 32*i, k4 = sk4), i = 0..3
    const int total = tiles_m * nchunks;

    int l_item = blockIdx.x, l_kc = 0;                     // load cursor (runs ahead of the compute cursor)
    auto issue_load = [&](int stage) {
        if (l_item < total) {
            float* dstA = pf_smem + (size_t)stage * (STAGE_A + STAGE_W);
            float* dstW = dstA + STAGE_A;
            const int tm = l_item / nchunks, n0 = (l_item - tm * nchunks) * chunk_cols;
            const uint32_t m0 = (uint32_t)tm * PF_BM;
            const int k = l_kc + sk4 * 4;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t m = m0 + srow + 32 * i;
                const bool ok = (m < M) && (k < K);
                pf_cp_async16(dstA + (srow + 32 * i) * PF_LDA + sk4 * 4, ok ? (const void*)(p.A + (size_t)m * K + k) : (const void*)p.A, ok);
            }
#pragma unroll
            for (int j = 0; j < (PF_KC * WV + PF_THREADS - 1) / PF_THREADS; j++) {
                const int idx = tid + j * PF_THREADS;
                if (idx < PF_KC * WV) {
                    const int wr = idx / WV, wc = (idx - wr * WV) * 4;
                    const bool ok = (l_kc + wr < K) && (wc < chunk_cols) && (n0 + wc < N);
                    pf_cp_async16(dstW + wr * COVER + wc, ok ? (const void*)(p.Wraw + (size_t)(l_kc + wr) * N + n0 + wc) : (const void*)p.Wraw, ok);
                }
            }
            l_kc += PF_KC;
            if (l_kc >= K) { l_kc = 0; l_item += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < PF_STAGES - 1; s++) issue_load(s);

    int stage = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int tm = item / nchunks, n0 = (item - tm * nchunks) * chunk_cols;
        const uint32_t m0 = (uint32_t)tm * PF_BM;
        const int ncols = min(chunk_cols, N - n0);            // columns of this chunk
        const bool busy = strip * NT < ncols;
        const int rbase = rg * (32 * R) + lane;                // tile row of this lane's first row; row i -> rbase + 32*i
        uint32_t simg[4] = {0, 0, 0, 0};                       // SE: image of the 4 rows this thread stages
        if (p.se_scale) {
#pragma unroll
            for (int i = 0; i < 4; i++) simg[i] = min(m0 + srow + 32 * i, M - 1) / (uint32_t)p.hw;
        }
        float acc[R][NT];
        {   // accumulators start at the bias of this warp's strip (loaded here, long before the epilogue needs it)
            const int nb0 = n0 + strip * NT;
#pragma unroll
            for (int c4 = 0; c4 < NT / 4; c4++) {
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && busy && strip * NT + c4 * 4 < ncols) bv = __ldg(reinterpret_cast<const float4*>(p.bias + nb0 + c4 * 4));
#pragma unroll
                for (int i = 0; i < R; i++) { acc[i][c4 * 4] = bv.x; acc[i][c4 * 4 + 1] = bv.y; acc[i][c4 * 4 + 2] = bv.z; acc[i][c4 * 4 + 3] = bv.w; }
            }
        }

        for (int kc = 0; kc < K; kc += PF_KC) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PF_STAGES - 2) : "memory");
            float* As = pf_smem + (size_t)stage * (STAGE_A + STAGE_W);
            const float* Ws = As + STAGE_A;
            if (p.se_scale) {                                  // scale the chunks this thread copied (its own cp.async data has landed)
                const int k = kc + sk4 * 4;
                if (k < K) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float4* ap = reinterpret_cast<float4*>(As + (srow + 32 * i) * PF_LDA + sk4 * 4);
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.se_scale + (size_t)simg[i] * K + k));
                        float4 v = *ap; v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w; *ap = v;
                    }
                }
            }
            __syncthreads();                                   // chunk `stage` is visible; everyone is done with the slot refilled next
            issue_load((stage + PF_STAGES - 1) % PF_STAGES);
            if (busy) {
                const int k4n = min(PF_KC, K - kc) >> 2;
                const float* ap = As + rbase * PF_LDA;
                const float* wp = Ws + strip * NT;
#pragma unroll 1
                for (int k4 = 0; k4 < k4n; k4++) {
                    float4 a[R];
#pragma unroll
                    for (int i = 0; i < R; i++) a[i] = *reinterpret_cast<const float4*>(ap + i * 32 * PF_LDA + k4 * 4);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
                        for (int c4 = 0; c4 < NT / 4; c4++) {
                            const float4 w = *reinterpret_cast<const float4*>(wp + (k4 * 4 + kk) * COVER + c4 * 4);   // broadcast
#pragma unroll
                            for (int i = 0; i < R; i++) {
                                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                                acc[i][c4 * 4 + 0] = fmaf(av, w.x, acc[i][c4 * 4 + 0]); acc[i][c4 * 4 + 1] = fmaf(av, w.y, acc[i][c4 * 4 + 1]);
                                acc[i][c4 * 4 + 2] = fmaf(av, w.z, acc[i][c4 * 4 + 2]); acc[i][c4 * 4 + 3] = fmaf(av, w.w, acc[i][c4 * 4 + 3]);
                            }
                        }
                    }
                }
            }
            stage = (stage + 1) % PF_STAGES;
        }
        if (!busy) continue;
        // epilogue: this lane's R rows x the strip's columns (N % 4 == 0, so float4 groups are all-in or all-out)
        const int nb = n0 + strip * NT;
        auto epilogue = [&](auto actfn) {
#pragma unroll
            for (int i = 0; i < R; i++) {
                const uint32_t r = m0 + rbase + 32 * i;
                if (r >= M) continue;
                const float* resp = nullptr;
                if (p.res_mode == 1) resp = p.res + (size_t)r * N;
                else if (p.res_mode == 2) {
                    const uint32_t row = r / (uint32_t)p.W, x = r - row * p.W, b = row / (uint32_t)p.H, y = row - b * p.H;
                    resp = p.res + ((size_t)(b * (p.H / p.up) + y / p.up) * (p.W / p.up) + x / p.up) * N;
                }
                float* op = p.out + (size_t)r * N;
#pragma unroll
                for (int c4 = 0; c4 < NT / 4; c4++) {
                    const int n = nb + c4 * 4;
                    if (strip * NT + c4 * 4 >= ncols) break;
                    float4 v = make_float4(acc[i][c4 * 4], acc[i][c4 * 4 + 1], acc[i][c4 * 4 + 2], acc[i][c4 * 4 + 3]);
                    v.x = actfn(v.x); v.y = actfn(v.y); v.z = actfn(v.z); v.w = actfn(v.w);
                    if (resp) { const float4 q = __ldg(reinterpret_cast<const float4*>(resp + n)); v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
                    if (p.round_tf32) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
                    *reinterpret_cast<float4*>(op + n) = v;
                }
            }
        };
        if (p.act == ACT_RELU) epilogue(ActRelu());
        else if (p.act == ACT_HSWISH) epilogue(ActHswish());
        else epilogue(ActNone());
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Pixel-per-thread variant for the layers whose GEMM is too shallow for the tiled kernel above (K <= 96 with N <= 24: the
// projections of features.1 / features.3, 49 -> 35 us and 56 -> 50 us at B = 256; and the K = 24 lateral as 4 warps = 4 channel
// chunks per 32 pixels, 95-103 -> 87 us): the tiled kernel spends more
// issue slots on its per-item prologue / epilogue than on 6..22 k-steps, and its lane-per-row float4 stores are 32 half-empty
// sector writes per instruction.  Here a thread owns NPT output channels of one pixel (N / NPT threads per pixel), the weights sit in
// shared memory (one broadcast LDS.128 per 4 FMAs), the pixel's K inputs come as float4 through L1, and the warp's outputs --
// contiguous in memory -- leave through a padded shared-memory transposition as full 512-byte rows.  Accumulation order
// (bias first, k ascending, SE factor applied to the input) is that of the tiled kernel: the results are bit-identical.
template <int NPT>
__global__ void __launch_bounds__(256, NPT == 32 ? 3 : 0) pw_px_kernel(const PwMma p) {
    extern __shared__ __align__(16) float px_smem[];           // W [K][N / NPT][NPT + 4] | staging [8 warps][32][NPT + 4]
    constexpr int LDS_ = NPT + 4;                              // padded chunk / row pitch: the N / NPT chunk addresses of one LDS.128 fall into different banks
    const int K = p.K, N = p.N, tpp = N / NPT;
    float* Ws = px_smem;
    float* stg = px_smem + K * tpp * LDS_ + (threadIdx.x >> 5) * (32 * LDS_);
    for (int i = threadIdx.x; i < K * N / 4; i += 256) {
        const int k = i / (N / 4), rem = i - k * (N / 4), ch = rem / (NPT / 4), q = rem - ch * (NPT / 4);
        *reinterpret_cast<float4*>(Ws + (k * tpp + ch) * LDS_ + q * 4) = __ldg(reinterpret_cast<const float4*>(p.Wraw) + i);
    }
    __syncthreads();
    // a warp = 32 consecutive pixels x ONE channel chunk (all lanes read the same weights: full-warp broadcasts; four distinct
    // chunk addresses per LDS.128 cost four shared-memory wavefronts each and made the K = 24 lateral LSU-bound at 114 us)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = warp % tpp, n0 = ch * NPT;
    const int64_t m0 = ((int64_t)blockIdx.x * (8 / tpp) + warp / tpp) * 32;       // first pixel of this warp
    if (m0 >= p.M) return;
    const int64_t m = m0 + lane < p.M ? m0 + lane : p.M - 1;
    float acc[NPT];
#pragma unroll
    for (int c4 = 0; c4 < NPT / 4; c4++) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + c4);
        acc[c4 * 4] = bv.x; acc[c4 * 4 + 1] = bv.y; acc[c4 * 4 + 2] = bv.z; acc[c4 * 4 + 3] = bv.w;
    }
    const float4* ap = reinterpret_cast<const float4*>(p.A + (size_t)m * K);
    const float4* sp = p.se_scale ? reinterpret_cast<const float4*>(p.se_scale + (size_t)(m / p.hw) * K) : nullptr;
#pragma unroll 2
    for (int k4 = 0; k4 < K / 4; k4++) {
        float4 a = __ldg(ap + k4);
        if (sp) { const float4 sc = __ldg(sp + k4); a.x *= sc.x; a.y *= sc.y; a.z *= sc.z; a.w *= sc.w; }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const float av = kk == 0 ? a.x : kk == 1 ? a.y : kk == 2 ? a.z : a.w;
            const float4* wr = reinterpret_cast<const float4*>(Ws + ((k4 * 4 + kk) * tpp + ch) * LDS_);
#pragma unroll
            for (int c4 = 0; c4 < NPT / 4; c4++) {
                const float4 w = wr[c4];
                acc[c4 * 4] = fmaf(av, w.x, acc[c4 * 4]); acc[c4 * 4 + 1] = fmaf(av, w.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(av, w.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(av, w.w, acc[c4 * 4 + 3]);
            }
        }
    }
    const float* resp = nullptr;
    if (p.res_mode == 1) resp = p.res + (size_t)m * N + n0;
    else if (p.res_mode == 2) {
        const uint32_t r = (uint32_t)m, row = r / (uint32_t)p.W, x = r - row * p.W, b = row / (uint32_t)p.H, y = row - b * p.H;
        resp = p.res + ((size_t)(b * (p.H / p.up) + y / p.up) * (p.W / p.up) + x / p.up) * N + n0;
    }
#pragma unroll
    for (int c4 = 0; c4 < NPT / 4; c4++) {
        float4 v = make_float4(acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]);
        if (p.act == ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        else if (p.act == ACT_HSWISH) { ActHswish h; v.x = h(v.x); v.y = h(v.y); v.z = h(v.z); v.w = h(v.w); }
        if (resp) { const float4 q = __ldg(reinterpret_cast<const float4*>(resp) + c4); v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
        if (p.round_tf32) { v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w); }
        *reinterpret_cast<float4*>(stg + lane * LDS_ + c4 * 4) = v;
    }
    __syncwarp();
    // store in memory order: (NPT / 4) lanes cover one pixel's chunk (64 .. 128 contiguous bytes), full sectors per instruction
    const int npix = (int)(p.M - m0 < 32 ? p.M - m0 : 32);
#pragma unroll
    for (int j = 0; j < NPT / 4; j++) {
        const int e = j * 32 + lane, src_lane = e / (NPT / 4), q = e - src_lane * (NPT / 4);
        if (src_lane < npix)
            *reinterpret_cast<float4*>(p.out + (size_t)(m0 + src_lane) * N + n0 + q * 4) = *reinterpret_cast<const float4*>(stg + src_lane * LDS_ + q * 4);
    }
}

static int px_npt(const PwMma& p) {             // channels per thread of the pixel-per-thread kernel, 0 = not eligible
    if (p.out2 || (p.K & 3) || p.K > 96) return 0;
    if (p.N == 16 || p.N == 24) return p.N;
    if (p.N == 128 && p.K <= 32) return 32;            // the K = 24 lateral: 4 warps (channel chunks) per 32 pixels
    return 0;
}

// (R, NT, chunks) per layer: minimise chunks * (K/4 * cycles per 4-k step + per-item overhead), the step cost being the
// larger of the shared-memory wavefronts 8*(4R+NT) and the per-sub-partition issue slots 2*(4*R*NT+R+NT); every extra
// column chunk re-streams the A tile and pays the item prologue / epilogue again (~1500 issue cycles); items run in waves
// of two CTAs per SM.
struct PfCfg { int R, NT, nchunks, chunk_cols; };
static PfCfg choose_pf_cfg(int K, int N, int tiles_m, int sm_count) {
    PfCfg best{4, 16, 1, 128};
    double best_cost = 1e30;
    const int rs[3] = {4, 2, 1}, nts[3] = {16, 12, 8};
    for (int R : rs)
        for (int NT : nts) {
            const int cover = 2 * R * NT;
            const int nchunks = (N + cover - 1) / cover;
            int cc = ((N + nchunks - 1) / nchunks + 3) & ~3;          // even split, float4 granular
            if (cc > cover) cc = cover;
            const double lds = 8.0 * (4 * R + NT), iss = 2.0 * (4 * R * NT + R + NT);
            const double waves = std::ceil((double)tiles_m * nchunks / (2.0 * sm_count));      // small maps: chunks also buy parallelism
            const double cost = waves * ((K / 4.0) * std::max(lds, iss) + 1500.0);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = PfCfg{R, NT, nchunks, cc}; }
        }
    return best;
}

int pointwise_ffma(const PwMma& p, int sm_count, cudaStream_t s) {
    if ((p.K & 7) || (p.N & 3)) { set_error("pointwise_ffma: K must be a multiple of 8 and N of 4 (got %d, %d)", p.K, p.N); return KPD_ERR_UNSUPPORTED; }
    if (!p.Wraw) { set_error("pointwise_ffma: missing weights"); return KPD_ERR_ARG; }
    if (p.M <= 0) return 0;
    if (p.M >= ((int64_t)1 << 31) - PF_BM) { set_error("pointwise_ffma: M = %lld pixels exceeds the 32-bit pixel index", (long long)p.M); return KPD_ERR_UNSUPPORTED; }
    if (const int npt = px_npt(p)) {
        const int tpp = p.N / npt;                                // warps (channel chunks) per group of 32 pixels: 1 or 4
        const size_t smem = sizeof(float) * ((size_t)p.K * tpp * (npt + 4) + 8 * 32 * (npt + 4));
        const int64_t groups = (p.M + 31) / 32;                   // 8 / tpp pixel groups per CTA
        const int grid = (int)((groups + (8 / tpp) - 1) / (8 / tpp));
        constexpr int PX_SMEM_MAX = 96 * 1024;          // the attribute is set once per device: K <= 96, N <= 128 always fit
        static uint64_t px_attr[3] = {};
        if (npt == 16) { KPD_CUDA(ensure_smem_attr(pw_px_kernel<16>, PX_SMEM_MAX, px_attr[0])); pw_px_kernel<16><<<grid, 256, smem, s>>>(p); }
        else if (npt == 24) { KPD_CUDA(ensure_smem_attr(pw_px_kernel<24>, PX_SMEM_MAX, px_attr[1])); pw_px_kernel<24><<<grid, 256, smem, s>>>(p); }
        else { KPD_CUDA(ensure_smem_attr(pw_px_kernel<32>, PX_SMEM_MAX, px_attr[2])); pw_px_kernel<32><<<grid, 256, smem, s>>>(p); }
        KPD_LAUNCH_CHECK();
        return 0;
    }
    const int tiles_m = (int)((p.M + PF_BM - 1) / PF_BM);
    const PfCfg c = choose_pf_cfg(p.K, p.N, tiles_m, sm_count);
    const int cover = 2 * c.R * c.NT;
    const size_t smem = sizeof(float) * PF_STAGES * (PF_BM * PF_LDA + PF_KC * cover);
    const int grid = (int)std::min<int64_t>((int64_t)tiles_m * c.nchunks, (int64_t)sm_count * 2);
    static uint64_t attr[9] = {};
#define KPD_PF_LAUNCH(RR, NN, SLOT)                                                              \
    do {                                                                                         \
        KPD_CUDA(ensure_smem_attr(pw_ffma_kernel<RR, NN>, (int)smem, attr[SLOT]));               \
        pw_ffma_kernel<RR, NN><<<grid, PF_THREADS, smem, s>>>(p, tiles_m, c.nchunks, c.chunk_cols); \
    } while (0)
#define KPD_PF_NT(RR, BASE)                                                                      \
    do {                                                                                         \
        if (c.NT == 16) KPD_PF_LAUNCH(RR, 16, BASE);                                             \
        else if (c.NT == 12) KPD_PF_LAUNCH(RR, 12, BASE + 1);                                    \
        else KPD_PF_LAUNCH(RR, 8, BASE + 2);                                                     \
    } while (0)
    if (c.R == 4) KPD_PF_NT(4, 0);
    else if (c.R == 2) KPD_PF_NT(2, 3);
    else KPD_PF_NT(1, 6);
#undef KPD_PF_NT
#undef KPD_PF_LAUNCH
    KPD_LAUNCH_CHECK();
    return 0;
}

}  // namespace kpd
