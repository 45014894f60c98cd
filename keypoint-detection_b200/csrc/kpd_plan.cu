// Plan lifecycle: BN folding, weight repacking and upload; workspace layout; error/profiling plumbing.
// Replaces nn.Module construction + load_state_dict of the reference for the hot path
// (dll/models/keypoint_model.py:49-71, scripts/predict.py:47-57).
#include "kpd_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <map>
#include <algorithm>

namespace kpd {

static thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// ---- optional CUDA-event profiling of tagged regions (bench roofline; off by default) ----
struct ProfEntry { double ms = 0; int64_t n = 0; };
// level 1: stage tags only ("body", "fpn0_conv", ...: one event pair per stage, the launches inside a stage stay back to back);
// level 2: also the nested "stage.part" tags -- the extra events between the short trunk launches cost GPU idle time that lands
// inside the enclosing stage's pair, so stage totals are read at level 1 and the breakdown at level 2.
static int g_prof_on = 0;
static std::map<std::string, ProfEntry> g_prof;
struct Pending { std::string tag; cudaEvent_t a, b; bool live; };
static std::vector<Pending> g_pending;
static std::vector<Pending> g_open;

void prof_begin(const char* tag, cudaStream_t s) {
    if (!g_prof_on) return;
    Pending p; p.tag = tag; p.a = p.b = nullptr;
    p.live = g_prof_on >= 2 || strchr(tag, '.') == nullptr;
    if (p.live) {
        cudaEventCreate(&p.a); cudaEventCreate(&p.b);
        cudaEventRecord(p.a, s);
    }
    g_open.push_back(p);
}
void prof_end(cudaStream_t s) {
    if (!g_prof_on || g_open.empty()) return;
    Pending p = g_open.back(); g_open.pop_back();
    if (!p.live) return;
    cudaEventRecord(p.b, s);
    g_pending.push_back(p);
}
static void prof_collect() {
    for (auto& p : g_pending) {
        if (cudaEventSynchronize(p.b) == cudaSuccess) {
            float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
            g_prof[p.tag].ms += ms; g_prof[p.tag].n += 1;
        }
        cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    g_pending.clear();
}

// ---- host-side packing helpers ----
struct Blob {
    std::vector<char> h;
    size_t alloc(size_t bytes) {
        size_t off = (h.size() + 255) & ~size_t(255);
        h.resize(off + bytes, 0);
        return off;
    }
    float* f(size_t off) { return reinterpret_cast<float*>(h.data() + off); }
};

struct Cursor {
    const void* const* p; int n; int i = 0;
    const float* next() { return i < n ? static_cast<const float*>(p[i++]) : nullptr; }
    void skip() { i++; }
};

struct BN { const float *w, *b, *m, *v; };
static BN read_bn(Cursor& c) { BN r; r.w = c.next(); r.b = c.next(); r.m = c.next(); r.v = c.next(); c.skip(); return r; }

// scale/shift of an eval-mode BatchNorm: y = x*sc + sh
static void bn_affine(const BN& bn, int ch, double eps, std::vector<double>& sc, std::vector<double>& sh) {
    sc.resize(ch); sh.resize(ch);
    for (int i = 0; i < ch; i++) {
        sc[i] = (double)bn.w[i] / sqrt((double)bn.v[i] + eps);
        sh[i] = (double)bn.b[i] - (double)bn.m[i] * sc[i];
    }
}

// conv weight [cout][cin][kh][kw] (+BN, + optional conv bias) -> w_out [(kh*kw)*cin][cout] (tap-major, cout contiguous)
static void pack_conv_kc(Blob& B, const float* w, const float* cbias, const BN* bn, double eps, int cout, int cin, int k,
                         size_t& off_w, size_t& off_b) {
    std::vector<double> sc(cout, 1.0), sh(cout, 0.0);
    if (bn) bn_affine(*bn, cout, eps, sc, sh);
    off_w = B.alloc(sizeof(float) * (size_t)k * k * cin * cout);
    off_b = B.alloc(sizeof(float) * cout);
    float* wo = B.f(off_w); float* bo = B.f(off_b);
    for (int o = 0; o < cout; o++) {
        for (int c = 0; c < cin; c++)
            for (int t = 0; t < k * k; t++)
                wo[((size_t)t * cin + c) * cout + o] = (float)((double)w[((size_t)o * cin + c) * k * k + t] * sc[o]);
        bo[o] = (float)((cbias ? (double)cbias[o] * sc[o] : 0.0) + sh[o]);
    }
}

static __nv_bfloat16 to_bf16(float f) { return __float2bfloat16_rn(f); }
// fp32 -> tf32 round-to-nearest (ties away), same result as cvt.rna.tf32.f32 for finite values
static float to_tf32(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&f, &u, 4);
    return f;
}

// dense conv for the tensor-core path: [cout][taps*cin] K-major (k index = tap*cin + c), fp32 and bf16 copies
static void pack_dense(Blob& B, const float* w, const float* cbias, const BN* bn, double eps, int cout, int cin, int k,
                       size_t off[4]) {
    std::vector<double> sc(cout, 1.0), sh(cout, 0.0);
    if (bn) bn_affine(*bn, cout, eps, sc, sh);
    const int taps = k * k; const size_t K = (size_t)taps * cin;
    off[0] = B.alloc(sizeof(float) * taps * cin * cout);     // simt [taps][cin][cout]
    off[1] = B.alloc(sizeof(float) * cout * K);              // tf32 [cout][K]
    off[2] = B.alloc(sizeof(__nv_bfloat16) * cout * K);      // bf16 [cout][K]
    off[3] = B.alloc(sizeof(float) * cout);
    for (int o = 0; o < cout; o++) {
        for (int c = 0; c < cin; c++)
            for (int t = 0; t < taps; t++) {
                float v = (float)((double)w[((size_t)o * cin + c) * taps + t] * sc[o]);
                B.f(off[0])[((size_t)t * cin + c) * cout + o] = v;
                B.f(off[1])[(size_t)o * K + (size_t)t * cin + c] = to_tf32(v);
                reinterpret_cast<__nv_bfloat16*>(B.h.data() + off[2])[(size_t)o * K + (size_t)t * cin + c] = to_bf16(v);
            }
        B.f(off[3])[o] = (float)((cbias ? (double)cbias[o] * sc[o] : 0.0) + sh[o]);
    }
}

// tf32 hi / lo split of a [K][N] matrix stored in the blob at `off_w`, as K-major rows [2][Np][Kp] (kpd_pwu.cu)
static size_t pack_umma_split(Blob& B, size_t off_w, int K, int N) {
    int nchunks, Nc; kpd::pwu_chunks(N, &nchunks, &Nc);
    const int Np = nchunks * Nc, Kp = (K + 31) & ~31;
    const size_t off = B.alloc(sizeof(float) * 2 * (size_t)Np * Kp);
    float* d = B.f(off);
    memset(d, 0, sizeof(float) * 2 * (size_t)Np * Kp);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            const float w = B.f(off_w)[(size_t)k * N + n];
            const float hi = to_tf32(w);
            d[(size_t)n * Kp + k] = hi;
            d[((size_t)Np + n) * Kp + k] = to_tf32(w - hi);
        }
    return off;
}

static size_t pack_raw(Blob& B, const float* src, size_t n) {
    size_t off = B.alloc(sizeof(float) * n);
    memcpy(B.f(off), src, sizeof(float) * n);
    return off;
}
// [rows][cols] -> transposed [cols][rows]
static size_t pack_transposed(Blob& B, const float* src, int rows, int cols) {
    size_t off = B.alloc(sizeof(float) * (size_t)rows * cols);
    float* d = B.f(off);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) d[(size_t)c * rows + r] = src[(size_t)r * cols + c];
    return off;
}

static const int kBlocks[11][8] = {   // cin, k, exp, cout, se, act(1=RE,2=HS), stride  (tv:models/mobilenetv3.py:246-258)
    {16, 3, 16, 16, 1, ACT_RELU, 2, 0},   {16, 3, 72, 24, 0, ACT_RELU, 2, 0},   {24, 3, 88, 24, 0, ACT_RELU, 1, 0},
    {24, 5, 96, 40, 1, ACT_HSWISH, 2, 0}, {40, 5, 240, 40, 1, ACT_HSWISH, 1, 0}, {40, 5, 240, 40, 1, ACT_HSWISH, 1, 0},
    {40, 5, 120, 48, 1, ACT_HSWISH, 1, 0}, {48, 5, 144, 48, 1, ACT_HSWISH, 1, 0}, {48, 5, 288, 96, 1, ACT_HSWISH, 2, 0},
    {96, 5, 576, 96, 1, ACT_HSWISH, 1, 0}, {96, 5, 576, 96, 1, ACT_HSWISH, 1, 0}};

static int make_divisible(int v, int d) {
    int nv = ((v + d / 2) / d) * d; if (nv < d) nv = d;
    if (nv < 0.9 * v) nv += d;
    return nv;
}

}  // namespace kpd

using namespace kpd;

extern "C" const char* kpd_last_error(void) { return kpd::g_err; }
extern "C" int kpd_abi_version(void) { return KPD_ABI_VERSION; }
extern "C" int64_t kpd_launch_count(int reset) { int64_t v = g_launches; if (reset) g_launches = 0; return v; }
extern "C" int kpd_profile_enable(int on) { prof_collect(); g_open.clear(); g_prof_on = on < 0 ? 0 : on; if (on) g_prof.clear(); return 0; }
extern "C" int kpd_profile_read(const char* tag, double* total_ms, int64_t* launches) {
    prof_collect();
    auto it = g_prof.find(tag ? tag : "");
    if (it == g_prof.end()) { if (total_ms) *total_ms = 0; if (launches) *launches = 0; return KPD_ERR_ARG; }
    if (total_ms) *total_ms = it->second.ms;
    if (launches) *launches = it->second.n;
    return 0;
}

extern "C" int kpd_plan_create(kpd_plan** out, const kpd_model_desc* desc, const void* const* h_state, int n_state,
                               int device) {
    KPD_CHECK_ARG(out && desc && h_state, "kpd_plan_create: null argument");
    KPD_CHECK_ARG(desc->abi_version == KPD_ABI_VERSION, "kpd_plan_create: ABI version %d != %d", desc->abi_version, KPD_ABI_VERSION);
    KPD_CHECK_ARG(n_state == KPD_NUM_STATE_TENSORS, "kpd_plan_create: expected %d state tensors, got %d", KPD_NUM_STATE_TENSORS, n_state);
    KPD_CHECK_ARG(desc->in_channels == 1 || desc->in_channels == 3, "in_channels must be 1 or 3");
    KPD_CHECK_ARG(desc->num_keypoints == 17, "num_keypoints must be 17");
    for (int i = 0; i < n_state; i++) KPD_CHECK_ARG(h_state[i] != nullptr, "state tensor %d is null", i);
    cudaDeviceProp prop;
    KPD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor); return KPD_ERR_UNSUPPORTED; }
    // the upload below runs on `device`; the caller's current device is restored on every exit path (a plan may be
    // built for a GPU other than the thread's current one)
    int prev_device = device;
    KPD_CUDA(cudaGetDevice(&prev_device));
    struct DeviceRestore { int prev, cur; ~DeviceRestore() { if (prev != cur) cudaSetDevice(prev); } } device_restore{prev_device, device};
    KPD_CUDA(cudaSetDevice(device));

    const int cin0 = desc->in_channels;
    Blob B; Cursor c{h_state, n_state};
    kpd_plan* p = new kpd_plan();
    memset(p, 0, sizeof(*p));
    p->device = device; p->in_channels = cin0; p->num_keypoints = desc->num_keypoints; p->sm_count = prop.multiProcessorCount;
    std::vector<std::pair<const void**, size_t>> fix;     // (address of a pointer field, blob offset)
    auto reg = [&](const void* field, size_t off) { fix.push_back({(const void**)field, off}); };
    size_t ow, ob;
    const double EPS_BODY = 1e-3, EPS_HEAD = 1e-5;

    {   // stem: features.0
        const float* w = c.next(); BN bn = read_bn(c);
        pack_conv_kc(B, w, nullptr, &bn, EPS_BODY, 16, cin0, 3, ow, ob);
        reg(&p->w_stem, ow); reg(&p->b_stem, ob);
    }
    for (int i = 0; i < 11; i++) {
        MBBlock& m = p->blocks[i];
        m.cin = kBlocks[i][0]; m.k = kBlocks[i][1]; m.exp = kBlocks[i][2]; m.cout = kBlocks[i][3];
        m.se = kBlocks[i][4]; m.act = kBlocks[i][5]; m.stride = kBlocks[i][6];
        m.res = (m.stride == 1 && m.cin == m.cout);
        m.sq = m.se ? make_divisible(m.exp / 4, 8) : 0;
        size_t o_wexp = 0, o_bexp = 0;
        if (m.exp != m.cin) {
            const float* w = c.next(); BN bn = read_bn(c);
            pack_conv_kc(B, w, nullptr, &bn, EPS_BODY, m.exp, m.cin, 1, ow, ob);
            o_wexp = ow; o_bexp = ob;
            reg(&m.w_exp, ow); reg(&m.b_exp, ob);
            reg(&m.wu_exp, pack_umma_split(B, ow, m.cin, m.exp));
        }
        {   // depthwise [exp][1][k][k] -> [k*k][exp]
            const float* w = c.next(); BN bn = read_bn(c);
            std::vector<double> sc, sh; bn_affine(bn, m.exp, EPS_BODY, sc, sh);
            ow = B.alloc(sizeof(float) * m.k * m.k * m.exp); ob = B.alloc(sizeof(float) * m.exp);
            for (int ch = 0; ch < m.exp; ch++) {
                for (int t = 0; t < m.k * m.k; t++) B.f(ow)[(size_t)t * m.exp + ch] = (float)((double)w[(size_t)ch * m.k * m.k + t] * sc[ch]);
                B.f(ob)[ch] = (float)sh[ch];
            }
            reg(&m.w_dw, ow); reg(&m.b_dw, ob);
            if (m.exp != m.cin) {   // operands of the fused expand + depthwise kernel (kpd_ed.cu)
                int RB, slabs; kpd::ed_operand_shape(m.cin, &RB, &slabs);
                const int Kw = slabs * (RB / 4), nch = (m.exp + 31) / 32, Np = nch * 32, KK = m.k * m.k;
                const size_t owe = B.alloc(sizeof(float) * 2 * (size_t)Np * Kw), ope = B.alloc(sizeof(float) * (size_t)nch * (2 + KK) * 32);
                memset(B.f(owe), 0, sizeof(float) * 2 * (size_t)Np * Kw);
                memset(B.f(ope), 0, sizeof(float) * (size_t)nch * (2 + KK) * 32);
                for (int n = 0; n < m.exp; n++)
                    for (int k = 0; k < m.cin; k++) {
                        const float wv = B.f(o_wexp)[(size_t)k * m.exp + n];
                        const float hi = to_tf32(wv);
                        B.f(owe)[(size_t)n * Kw + k] = hi;
                        B.f(owe)[((size_t)Np + n) * Kw + k] = to_tf32(wv - hi);
                    }
                for (int n = 0; n < m.exp; n++) {
                    float* blk = B.f(ope) + (size_t)(n / 32) * (2 + KK) * 32;
                    blk[n % 32] = B.f(o_bexp)[n];
                    blk[32 + n % 32] = B.f(ob)[n];
                    for (int t = 0; t < KK; t++) blk[(2 + t) * 32 + n % 32] = B.f(ow)[(size_t)t * m.exp + n];
                }
                reg(&m.wed_exp, owe); reg(&m.ed_params, ope);
            }
        }
        if (m.se) {   // fc1 [sq][exp] -> [exp][sq]; fc2 [exp][sq] -> [sq][exp]
            const float* w1 = c.next(); const float* b1 = c.next(); const float* w2 = c.next(); const float* b2 = c.next();
            reg(&m.w_fc1, pack_transposed(B, w1, m.sq, m.exp)); reg(&m.b_fc1, pack_raw(B, b1, m.sq));
            reg(&m.w_fc2, pack_transposed(B, w2, m.exp, m.sq)); reg(&m.b_fc2, pack_raw(B, b2, m.exp));
        }
        {
            const float* w = c.next(); BN bn = read_bn(c);
            pack_conv_kc(B, w, nullptr, &bn, EPS_BODY, m.cout, m.exp, 1, ow, ob);
            reg(&m.w_proj, ow); reg(&m.b_proj, ob);
            reg(&m.wu_proj, pack_umma_split(B, ow, m.exp, m.cout));
        }
    }
    {   // features.12
        const float* w = c.next(); BN bn = read_bn(c);
        pack_conv_kc(B, w, nullptr, &bn, EPS_BODY, 576, 96, 1, ow, ob);
        reg(&p->w_last, ow); reg(&p->b_last, ob);
        reg(&p->wu_last, pack_umma_split(B, ow, 96, 576));
    }
    const int fpn_in[4] = {16, 24, 48, 576};
    const float* w_lat0_raw = nullptr;
    for (int i = 0; i < 4; i++) {
        const float* w = c.next();
        if (i == 0) w_lat0_raw = w;
        pack_conv_kc(B, w, nullptr, nullptr, 0, 128, fpn_in[i], 1, ow, ob);
        reg(&p->w_lat[i], ow);
        reg(&p->wu_lat[i], pack_umma_split(B, ow, fpn_in[i], 128));
    }
    auto dense = [&](DenseConv& d, const float* w, const float* cb, const BN* bn, double eps, int cout, int cin, int k) {
        size_t off[4]; pack_dense(B, w, cb, bn, eps, cout, cin, k, off);
        d.cin = cin; d.cout = cout; d.taps = k * k;
        reg(&d.w_simt, off[0]); reg(&d.w_tf32, off[1]); reg(&d.w_bf16, off[2]); reg(&d.bias, off[3]);
    };
    dense(p->lat0_conv, w_lat0_raw, nullptr, nullptr, 0, 128, 16, 1);
    for (int i = 0; i < 4; i++) {
        const float* w = c.next(); BN bn = read_bn(c);
        dense(p->fpn_conv[i], w, nullptr, &bn, EPS_HEAD, 128, 128, 3);
        if (i == 0) {
            // composed FPN-0 weights (kpd_fpn0.cu).  lateral0 = L0 f0 + up4(lat1) is never materialised:
            //   W1[o][tap][i] = sum_c Wc[o][c][tap] * L0[c][i]      (3x3 conv straight from the 16-channel stem output)
            //   W2[m][o][c]   = sum of the taps of Wc that read coarse neighbour (ny,nx) from sub-position (py,px) of a 4x4 block
            // m enumerates (py, px, ny, nx) in that nesting order over the non-empty neighbours (36 in total).
            std::vector<double> sc(128, 1.0), sh(128, 0.0);
            bn_affine(bn, 128, EPS_HEAD, sc, sh);
            const size_t o1 = B.alloc(sizeof(float) * 128 * 144), o2 = B.alloc(sizeof(float) * 36 * 128 * 128);
            for (int o = 0; o < 128; o++)
                for (int t = 0; t < 9; t++)
                    for (int ci = 0; ci < 16; ci++) {
                        double a = 0.0;
                        for (int cc = 0; cc < 128; cc++) a += (double)w[((size_t)o * 128 + cc) * 9 + t] * (double)w_lat0_raw[cc * 16 + ci];
                        B.f(o1)[(size_t)o * 144 + t * 16 + ci] = to_tf32((float)(a * sc[o]));
                    }
            auto nbr = [](int pp, int d) { const int q = pp + d; return q < 0 ? -1 : (q > 3 ? 1 : 0); };
            int m = 0;
            for (int py = 0; py < 4; py++)
                for (int px = 0; px < 4; px++)
                    for (int ny = -1; ny <= 1; ny++)
                        for (int nx = -1; nx <= 1; nx++) {
                            bool any = false;
                            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) if (nbr(py, dy) == ny && nbr(px, dx) == nx) any = true;
                            if (!any) continue;
                            for (int o = 0; o < 128; o++)
                                for (int cc = 0; cc < 128; cc++) {
                                    double a = 0.0;
                                    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++)
                                        if (nbr(py, dy) == ny && nbr(px, dx) == nx) a += (double)w[((size_t)o * 128 + cc) * 9 + (dy + 1) * 3 + (dx + 1)];
                                    B.f(o2)[((size_t)m * 128 + o) * 128 + cc] = to_tf32((float)(a * sc[o]));
                                }
                            m++;
                        }
            if (m != 36) { set_error("internal: composed FPN-0 weight count %d != 36", m); delete p; return KPD_ERR_STATE; }
            reg(&p->fpn0c_w1, o1); reg(&p->fpn0c_w2, o2);
        }
    }
    c.skip();   // person_detector.anchors (never read by the reference forward, person_head.py:141-166)
    size_t off_box[4], off_bbox[4];
    for (int i = 0; i < 4; i++) {
        const float* w = c.next(); const float* b = c.next();
        pack_conv_kc(B, w, b, nullptr, 0, 36, 128, 1, ow, ob); reg(&p->w_box[i], ow); reg(&p->b_box[i], ob);
        off_box[i] = ow; off_bbox[i] = ob;
    }
    for (int i = 0; i < 4; i++) {
        const float* w = c.next(); const float* b = c.next();
        pack_conv_kc(B, w, b, nullptr, 0, 9, 128, 1, ow, ob); reg(&p->w_cls[i], ow); reg(&p->b_cls[i], ob);
        // both heads of the level as ONE [128][48] matrix (36 box | 9 cls | 3 zero columns) for the tcgen05 GEMM
        const size_t oc = B.alloc(sizeof(float) * 128 * 48), obias = B.alloc(sizeof(float) * 48);
        for (int k = 0; k < 128; k++)
            for (int n = 0; n < 48; n++)
                B.f(oc)[k * 48 + n] = n < 36 ? B.f(off_box[i])[k * 36 + n] : (n < 45 ? B.f(ow)[k * 9 + n - 36] : 0.f);
        for (int n = 0; n < 48; n++) B.f(obias)[n] = n < 36 ? B.f(off_bbox[i])[n] : (n < 45 ? B.f(ob)[n - 36] : 0.f);
        reg(&p->wu_heads[i], pack_umma_split(B, oc, 128, 48)); reg(&p->b_heads[i], obias);
    }
    {   // heatmap head attention: Linear weights kept in their [out][in] form
        const float* w0 = c.next(); const float* b0 = c.next(); const float* w2 = c.next(); const float* b2 = c.next();
        reg(&p->hca_w0, pack_raw(B, w0, 4 * 64)); reg(&p->hca_b0, pack_raw(B, b0, 4));
        reg(&p->hca_w2, pack_raw(B, w2, 64 * 4)); reg(&p->hca_b2, pack_raw(B, b2, 64));
        const float* sw = c.next(); const float* sb = c.next();
        reg(&p->sa_w, pack_raw(B, sw, 2 * 49)); reg(&p->sa_b, pack_raw(B, sb, 1));
    }
    {
        const float* w = c.next(); const float* b = c.next(); BN bn = read_bn(c);
        dense(p->head_conv[0], w, b, &bn, EPS_HEAD, 256, 64, 3);
        w = c.next(); b = c.next(); bn = read_bn(c);
        dense(p->head_conv[1], w, b, &bn, EPS_HEAD, 256, 256, 3);
        w = c.next(); b = c.next(); bn = read_bn(c);
        dense(p->head_conv[2], w, b, &bn, EPS_HEAD, 64, 256, 3);
        w = c.next(); b = c.next();
        pack_conv_kc(B, w, b, nullptr, 0, 17, 64, 1, ow, ob); reg(&p->w_final, ow); reg(&p->b_final, ob);
    }
    {
        const float* w0 = c.next(); const float* b0 = c.next(); const float* w2 = c.next(); const float* b2 = c.next();
        reg(&p->ca_w0, pack_raw(B, w0, 8 * 128)); reg(&p->ca_b0, pack_raw(B, b0, 8));
        reg(&p->ca_w2, pack_raw(B, w2, 128 * 8)); reg(&p->ca_b2, pack_raw(B, b2, 128));
    }
    if (c.i != n_state) { set_error("internal: consumed %d of %d state tensors", c.i, n_state); delete p; return KPD_ERR_STATE; }

    p->blob_bytes = B.h.size();
    cudaError_t e = cudaMalloc(&p->d_blob, p->blob_bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", p->blob_bytes, cudaGetErrorString(e)); delete p; return (int)e; }
    e = cudaMemcpy(p->d_blob, B.h.data(), p->blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("weight upload failed: %s", cudaGetErrorString(e)); cudaFree(p->d_blob); delete p; return (int)e; }
    for (auto& f : fix) *f.first = static_cast<char*>(p->d_blob) + f.second;
    *out = p;
    return KPD_OK;
}

extern "C" void kpd_plan_destroy(kpd_plan* plan) {
    if (!plan) return;
    if (plan->d_blob) cudaFree(plan->d_blob);
    delete plan;
}

extern "C" int kpd_plan_workspace_layout(const kpd_plan* plan, int B, int S, int max_rois, int precision,
                                         kpd_workspace_layout* L) {
    KPD_CHECK_ARG(plan && L, "kpd_plan_workspace_layout: null argument");
    KPD_CHECK_ARG(B > 0 && S > 0 && S % 32 == 0, "batch must be > 0 and size a positive multiple of 32 (got %d, %d)", B, S);
    KPD_CHECK_ARG(max_rois >= 0, "max_rois must be >= 0");
    KPD_CHECK_ARG(precision >= 0 && precision <= 2, "unknown precision %d", precision);
    memset(L, 0, sizeof(*L));
    L->batch = B; L->size = S; L->max_rois = max_rois; L->precision = precision;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
    const size_t b = (size_t)B, h0 = S / 2, h1 = S / 8, h2 = S / 16, h3 = S / 32;
    const size_t n = (size_t)(max_rois > 0 ? max_rois : 1) + 1;     // +1: tile pairs may touch one ROI past the end
    L->elt_bytes = (precision == KPD_PREC_BF16) ? 2 : 4;
    L->feat[0] = take(4 * b * h0 * h0 * 16);
    L->feat[1] = take(4 * b * h1 * h1 * 24);
    L->feat[2] = take(4 * b * h2 * h2 * 48);
    L->feat[3] = take(4 * b * h3 * h3 * 576);
    L->lateral0 = take(4 * b * h0 * h0 * 128);
    L->fpn0 = take(4 * b * h0 * h0 * 128);
    L->pool_avg = take(4 * b * 128); L->pool_max = take(4 * b * 128); L->scores = take(4 * b * 128);
    L->topk = take(4 * b * 64);
    L->roi_count = take(16);
    L->roi_box = take(16 * n); L->roi_batch = take(4 * n); L->roi_slot = take(4 * n);
    L->img_count = take(4 * b);
    L->roi_feat = take(4 * n * 56 * 56 * 64);
    L->head_cw = take(4 * n * 64);
    L->head_sw = take(4 * n * 56 * 56);
    L->head_in = take(L->elt_bytes * n * 56 * 56 * 64);
    L->head_a = take(L->elt_bytes * n * 56 * 56 * 256);
    L->head_b = take(L->elt_bytes * n * 56 * 56 * 256);
    L->head_c = take(L->elt_bytes * n * 56 * 56 * 64);
    L->scratch = take(kpd::scratch_map(B, S, (int)n).total);
    L->total_bytes = off;
    return KPD_OK;
}

namespace kpd {
// Kernel-private scratch carved out of kpd_workspace_layout::scratch.
ScratchMap scratch_map(int B, int S, int n_rois) {
    ScratchMap m; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~size_t(1023); return o; };
    size_t max_exp = 0, max_dw = 0, max_out = 0, max_sq = 576;
    int h = S / 2;
    for (int i = 0; i < 11; i++) {
        const int exp = kBlocks[i][2], cout = kBlocks[i][3], stride = kBlocks[i][6];
        const int ho = h / stride;
        max_exp = std::max(max_exp, (size_t)h * h * exp);
        max_dw = std::max(max_dw, (size_t)ho * ho * exp);
        max_out = std::max(max_out, (size_t)ho * ho * cout);
        max_sq = std::max(max_sq, (size_t)ho * ((ho + 27) / 28) * exp);      // depthwise tiles (>= 1 row x <= 28.. columns) x channels
        h = ho;
    }
    const size_t b = (size_t)B;
    m.expand = take(4 * b * max_exp);
    m.dwout = take(4 * b * max_dw);
    m.ping = take(4 * b * max_out);
    m.pong = take(4 * b * max_out);
    m.se_pool = take(4 * b * max_sq); m.se_pool_floats = max_sq;
    m.se_scale = take(4 * b * 576);
    const size_t h1 = S / 8, h2 = S / 16, h3 = S / 32;
    m.f0_tf32 = take(4 * (b + 1) * (size_t)(S / 2) * (S / 2) * 16);     // tf32 copy of the stem output (+1 image: pair-wide TMA boxes)
    m.lat1 = take(4 * (b + 1) * h1 * h1 * 128);
    m.lat2 = take(4 * b * h2 * h2 * 128);
    m.lat3 = take(4 * b * h3 * h3 * 128);
    m.pool_part = take(4 * b * std::max<size_t>(std::max<size_t>(POOL_PARTS, (size_t)S * S / 512), (size_t)fpn0_fused_parts(S)) * 128 * 2);   // conv tiles per image
    m.roi_part = take(4 * (size_t)n_rois * ROI_PARTS * 64 * 2);
    m.att_maps = take(4 * (size_t)n_rois * 2 * 56 * 56);          // head attention: per-pixel mean / max over channels
    m.total = off;
    return m;
}
}  // namespace kpd

// ------------------------------------------------------------------------------------------------ tensor-map cache
#include "kpd_tma.cuh"
#include <unordered_map>
#include <mutex>
namespace kpd {
EncodeTiledFn kpd_driver_encode() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(ptr);
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}
namespace {
struct TmKey {
    uint64_t w[20];
    bool operator==(const TmKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmHash {
    size_t operator()(const TmKey& k) const { uint64_t h = 1469598103934665603ull; for (uint64_t v : k.w) { h ^= v; h *= 1099511628211ull; } return (size_t)h; }
};
std::mutex g_tm_mu;
std::unordered_map<TmKey, CUtensorMap, TmHash> g_tm_cache;
}  // namespace
CUresult kpd_encode_tiled(CUtensorMap* tm, CUtensorMapDataType dtype, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                          const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* es, CUtensorMapInterleave il,
                          CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    EncodeTiledFn drv = kpd_driver_encode();
    if (!drv || rank > 5) return CUDA_ERROR_NOT_SUPPORTED;
    int dev = 0; cudaGetDevice(&dev);
    TmKey k; memset(&k, 0, sizeof(k));
    k.w[0] = (uint64_t)(uintptr_t)ptr;
    k.w[1] = (uint64_t)dtype | ((uint64_t)rank << 8) | ((uint64_t)il << 16) | ((uint64_t)sw << 24) | ((uint64_t)l2 << 32) | ((uint64_t)oob << 40) | ((uint64_t)dev << 48);
    for (cuuint32_t i = 0; i < rank; i++) { k.w[2 + i] = dims[i]; k.w[12 + i] = ((uint64_t)box[i] << 32) | es[i]; }
    for (cuuint32_t i = 0; i + 1 < rank; i++) k.w[7 + i] = strides[i];
    {
        std::lock_guard<std::mutex> lock(g_tm_mu);
        auto it = g_tm_cache.find(k);
        if (it != g_tm_cache.end()) { *tm = it->second; return CUDA_SUCCESS; }
    }
    const CUresult r = drv(tm, dtype, rank, ptr, dims, strides, box, es, il, sw, l2, oob);
    if (r == CUDA_SUCCESS) {
        std::lock_guard<std::mutex> lock(g_tm_mu);
        if (g_tm_cache.size() > 8192) g_tm_cache.clear();
        g_tm_cache.emplace(k, *tm);
    }
    return r;
}
}  // namespace kpd
