// Shared declarations of the kpd_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/kpd_b200.h"

namespace kpd {

// ---------------------------------------------------------------- error handling
void set_error(const char* fmt, ...);
extern thread_local int64_t g_launches;
void prof_begin(const char* tag, cudaStream_t s);
void prof_end(cudaStream_t s);

#define KPD_CHECK_ARG(cond, ...)                                  \
    do { if (!(cond)) { kpd::set_error(__VA_ARGS__); return KPD_ERR_ARG; } } while (0)
#define KPD_CUDA(expr)                                            \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) {        \
        kpd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
        return (int)_e; } } while (0)
#define KPD_LAUNCH_CHECK()                                        \
    do { kpd::g_launches++; cudaError_t _e = cudaPeekAtLastError(); if (_e != cudaSuccess) { \
        kpd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
        return (int)_e; } } while (0)
#define KPD_TRY(expr) do { int _r = (expr); if (_r != 0) return _r; } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember which devices a kernel was configured on.
template <typename F>
inline cudaError_t ensure_smem_attr(F func, int bytes, uint64_t& done_mask) {
    int dev = 0; cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && (done_mask >> dev) & 1ull) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) done_mask |= 1ull << dev;
    return e;
}

struct ProfScope {
    cudaStream_t s;
    ProfScope(const char* tag, cudaStream_t st) : s(st) { prof_begin(tag, s); }
    ~ProfScope() { prof_end(s); }
};

// ---------------------------------------------------------------- activations
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_HSWISH = 2, ACT_HSIGMOID = 3, ACT_SIGMOID = 4 };

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case ACT_RELU:     return fmaxf(x, 0.f);
        case ACT_HSWISH:   return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case ACT_HSIGMOID: return fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
        case ACT_SIGMOID:  return 1.f / (1.f + expf(-x));
        default:           return x;
    }
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa) so that the tensor core's operand truncation is a no-op
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---------------------------------------------------------------- packed model
struct MBBlock {
    int cin, k, exp, cout, se, act, stride, res, sq;
    // device pointers (fp32). pointwise weights are [Cin][Cout] (K-major rows, cout contiguous)
    const float *w_exp, *b_exp;     // expand 1x1 (+folded BN)        [cin][exp], [exp]   (null if exp==cin)
    const float *wu_exp, *wu_proj;  // the same 1x1 weights as tf32 hi / lo split [2][Np][Kp] for the tcgen05 GEMM (kpd_pwu.cu)
    const float *w_dw, *b_dw;       // depthwise kxk (+folded BN)     [k*k][exp], [exp]
    const float *w_fc1, *b_fc1;     // SE fc1                          [exp][sq], [sq]
    const float *w_fc2, *b_fc2;     // SE fc2                          [sq][exp], [exp]
    const float *w_proj, *b_proj;   // project 1x1 (+folded BN)       [exp][cout], [cout]
    // fused expand + depthwise kernel (kpd_ed.cu): expand weights as tf32 hi / lo rows [2][ceil(exp/32)*32][slabs * row floats]
    // (K-major, zero padded) and per 32-channel chunk one block [2 + k*k][32] = expand bias | depthwise bias | tap rows
    const float *wed_exp, *ed_params;
};

struct DenseConv {          // 3x3 (or 1x1) dense conv, BN folded
    int cin, cout, taps;
    const float* w_simt;            // fp32 [taps][cin][cout]
    const float* w_tf32;            // fp32 [cout][taps*cin]   (K-major, UMMA B operand)
    const __nv_bfloat16* w_bf16;    // bf16 [cout][taps*cin]
    const float* bias;              // fp32 [cout]
};

}  // namespace kpd

struct kpd_plan {
    int device;
    int in_channels, num_keypoints;
    int sm_count;
    void* d_blob;           // one allocation holding every packed weight
    size_t blob_bytes;
    // body
    const float *w_stem, *b_stem;                 // [3*3*cin][16], [16]
    kpd::MBBlock blocks[11];
    const float *w_last, *b_last;                 // f12 [96][576], [576]
    const float* wu_last; const float* wu_lat[4];  // tf32 hi / lo splits for the tcgen05 pointwise GEMM
    // fpn
    const float* w_lat[4];                        // [cin_i][128]
    kpd::DenseConv lat0_conv;                     // lateral 0 (1x1, 16 -> 128) packed for the tcgen05 kernel (64-byte K rows)
    // FPN-0 in composed form (kpd_fpn0.cu): conv3x3(L0 f0 + up4(lat1)) = conv3x3_{W1}(f0) + sum over coarse neighbours W2 lat1
    const float* fpn0c_w1;                        // [128 cout][9 taps * 16 cin] tf32: fpn_convs[0] (BN folded) composed with lateral_convs[0]
    const float* fpn0c_w2;                        // [36][128 cout][128 cin] tf32: per sub-position / coarse-neighbour sums of the 3x3 taps
    kpd::DenseConv fpn_conv[4];
    // channel attention (model level): fc0 [8][128]+[8], fc2 [128][8]+[128]
    const float *ca_w0, *ca_b0, *ca_w2, *ca_b2;
    // heatmap head
    const float *hca_w0, *hca_b0, *hca_w2, *hca_b2;   // [4][64],[4],[64][4],[64]
    const float *sa_w, *sa_b;                          // [2][7][7], [1]
    kpd::DenseConv head_conv[3];                       // 64->256, 256->256, 256->64
    const float *w_final, *b_final;                    // [64][17], [17]
    // person heads
    const float *w_box[4], *b_box[4], *w_cls[4], *b_cls[4];   // [128][36],[36],[128][9],[9]
    const float *wu_heads[4], *b_heads[4];                    // box | cls heads side by side, padded to 48 columns: tf32 hi / lo split (kpd_pwu.cu) + bias [48]
};

namespace kpd {
constexpr int POOL_PARTS = 64;   // per-image partial blocks of the FPN-0 avg/max pool
constexpr int ROI_PARTS = 28;    // per-ROI partial blocks of the head channel-attention pool (2 rows each)
struct ScratchMap { size_t expand, dwout, ping, pong, se_pool, se_scale, f0_tf32, lat1, lat2, lat3, pool_part, roi_part, att_maps, total;
                    size_t se_pool_floats; };     // capacity of se_pool per image (per-tile squeeze partials)
ScratchMap scratch_map(int B, int S, int n_rois);

// ---- launchers implemented across the .cu files (all enqueue on `s`, return 0 / cudaError) ----
int run_body(const kpd_plan* p, const float* images, int B, int S, int strict_fp32, int want_f0_tf32, char* ws, const kpd_workspace_layout& L, cudaStream_t s);
int run_fpn_laterals(const kpd_plan* p, int B, int S, int round_tf32, int strict_fp32, int composed, char* ws, const kpd_workspace_layout& L, cudaStream_t s);
int fpn0_fused(const kpd_plan* plan, const float* f0t, const float* lat1, float* out, float* pool_part, int B, int S, cudaStream_t s);
int fpn0_fused_parts(int S);
int run_pools(const float* fpn0, int B, int HW, float* avg, float* mx, float* scratch, cudaStream_t s);
int run_pool_final(const float* partials, int B, int parts, int HW, float* avg, float* mx, cudaStream_t s);
int run_scores_topk(const kpd_plan* p, int B, const float* avg, const float* mx, float* scores, int32_t* topk, cudaStream_t s);
int conv_dense_simt(const DenseConv& c, const float* in, float* out, int N, int H, int W, int act,
                    const int32_t* n_valid, cudaStream_t s);
// final_layer[3] (1x1 conv 64 -> 17 + bias) + sigmoid fused into the epilogue of the 64-channel head conv (kpd_conv_umma.cu)
struct FinalFuse { const float* w; const float* b; float* out; const int32_t* slot; };   // [64][17], [17], heat maps [slot][17][H][W], slot per instance
int conv_dense_umma(const kpd_plan* p, const DenseConv& c, int prec_bf16, const void* in, void* out, int out_bf16,
                    int N, int H, int W, int act, int round_out_tf32, const int32_t* n_valid, cudaStream_t s,
                    float* pool_part = nullptr, const float* res = nullptr, int res_up = 1, const FinalFuse* fin = nullptr);
int pointwise(const float* in, const float* w, const float* bias, float* out, int64_t pixels, int cin, int cout,
              int act, const float* residual, const float* se_scale, int hw_per_image, cudaStream_t s);
// Batch-wide pointwise (1x1) convolution as a GEMM over all pixels: out[M][N] = act(A[M][K] * W[K][N] + bias) (+ residual).
struct PwMma {
    const float* A; const float* bias; float* out;
    int64_t M; int K, N, act;
    const float* se_scale; int hw;           // optional [image][K] scale applied to A (SE excitation), hw = pixels/image
    const float* res; int res_mode;          // 0 none, 1 same-shape add, 2 nearest-upsampled add (FPN top-down)
    int H, W, up;                            // res_mode 2: output map H x W, coarse map (H/up) x (W/up), N channels
    int round_tf32;                          // round the result to tf32 (it feeds a kind::tf32 tcgen05 MMA)
    const float* Wraw;                       // [K][N] row-major fp32 (CUDA-core FFMA GEMM, kpd_pwf.cu)
    const float* Wu;                         // [2][Np][Kp] tf32 hi / lo split, K-major rows (tcgen05 3xTF32 GEMM, kpd_pwu.cu)
    int exact;                               // 1: strict-fp32 mode -- CUDA-core FFMA GEMM only
    float* out2; int split, n2;              // tcgen05 kernel only: columns [0, split) go to out [M][split], columns [split, split + n2) to out2 [M][n2]
};
int pointwise_umma(const PwMma& p, int sm_count, cudaStream_t s);
void ed_operand_shape(int K, int* RB, int* slabs);
int expand_depthwise_fused(const MBBlock& m, const float* in, float* out, float* pool, size_t pool_capacity, int* pool_parts,
                           int B, int H, int sm_count, cudaStream_t s);
void pwu_chunks(int N, int* nchunks, int* Nc);
int pointwise_gemm(const PwMma& p, int sm_count, cudaStream_t s);      // dispatch: tcgen05 3xTF32 / CUDA-core fp32 (kpd_pw.cu)
int pointwise_ffma(const PwMma& p, int sm_count, cudaStream_t s);

}  // namespace kpd
