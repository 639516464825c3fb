// Standalone on-device check of the tcgen05 GEMM kernel against the SIMT verification kernel and a
// CPU dot product.  Build: see tests/cuda/Makefile.  Run on the B200 box:  ./gemm_test
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../video_gcp_b200/csrc/gemm_host.cuh"

extern "C" void gcp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}

using namespace gcp;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

static uint32_t rng_state = 12345;
static float frand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f;
}
static bf16* dev_bf16(size_t n, float scale, std::vector<float>* keep = nullptr) {
    std::vector<bf16> h(n);
    if (keep) keep->resize(n);
    for (size_t i = 0; i < n; ++i) {
        h[i] = __float2bfloat16(frand() * scale);
        if (keep) (*keep)[i] = __bfloat162float(h[i]);
    }
    bf16* d;
    CK(cudaMalloc(&d, n * 2));
    CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
    return d;
}
static float* dev_f32(size_t n, float scale, float offset = 0.f, std::vector<float>* keep = nullptr) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = frand() * scale + offset;
    if (keep) *keep = h;
    float* d;
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    return d;
}
template <class T>
static T* dev_zero(size_t n) {
    T* d;
    CK(cudaMalloc(&d, n * sizeof(T)));
    CK(cudaMemset(d, 0, n * sizeof(T)));
    return d;
}
static double max_diff_bf16(const bf16* a, const bf16* b, size_t n, double* maxabs) {
    std::vector<bf16> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 2, cudaMemcpyDeviceToHost));
    double m = 0, ma = 0;
    for (size_t i = 0; i < n; ++i) {
        double x = __bfloat162float(ha[i]), y = __bfloat162float(hb[i]);
        if (isnan(x) || isnan(y)) return 1e30;
        m = fmax(m, fabs(x - y));
        ma = fmax(ma, fabs(y));
    }
    *maxabs = ma;
    return m;
}
static double max_diff_f32(const float* a, const float* b, size_t n, double* maxabs) {
    std::vector<float> ha(n), hb(n);
    CK(cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost));
    double m = 0, ma = 0;
    for (size_t i = 0; i < n; ++i) {
        if (isnan(ha[i]) || isnan(hb[i])) return 1e30;
        m = fmax(m, fabs((double)ha[i] - hb[i]));
        ma = fmax(ma, fabs((double)hb[i]));
    }
    *maxabs = ma;
    return m;
}

static int n_fail = 0;
static void report(const char* name, double diff, double maxabs, double tol) {
    const bool ok = diff <= tol * fmax(1.0, maxabs) && maxabs > 0;
    printf("%-44s max|tc-ref| = %.3e  (max|ref| = %.3e)  %s\n", name, diff, maxabs, ok ? "OK" : "FAIL");
    if (!ok) ++n_fail;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs, cc %d.%d\n", prop.name, sms, prop.major, prop.minor);

    // ---------------- T1: plain linear, 2 K-segments, BN 128 and 256, CPU spot check ---------------
    for (int BN : {128, 256}) {
        const int rows = 512, N = 512, K1 = 128, K2 = 64, K = K1 + K2;
        std::vector<float> hA1, hA2, hW, hbias;
        bf16* A1 = dev_bf16((size_t)rows * K1, 1.f, &hA1);
        bf16* A2 = dev_bf16((size_t)rows * 256, 1.f, &hA2);  // ld 256, use cols 64..127
        bf16* Wd = dev_bf16((size_t)N * K, 0.1f, &hW);
        float* bias = dev_f32(N, 0.5f, 0.f, &hbias);
        float* o_tc = dev_zero<float>((size_t)rows * N);
        float* o_ref = dev_zero<float>((size_t)rows * N);
        bf16* b_tc = dev_zero<bf16>((size_t)rows * N);
        bf16* b_ref = dev_zero<bf16>((size_t)rows * N);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {A1, K1, 0, K1, ROW_LEVEL, 0, 0, {0}};
        a.seg[1] = {A2, 256, 64, K2, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], A1, rows, K1, K1, 128)) return 2;
        if (make_tmap_bf16(&a.a_map[1], A2, rows, 256, 256, 128)) return 2;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, BN)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {128, 0, 8};
        a.epi.bias = bias; a.epi.act = ACT_LRELU; a.epi.n_valid = N - 5;
        a.epi.out_f32_ld = N; a.epi.out_bf16_ld = N;
        a.epi.out_f32 = o_tc; a.epi.out_bf16 = b_tc;
        if (launch_gemm(a, BN, EPI_LINEAR, false, 0, sms)) return 2;
        a.epi.out_f32 = o_ref; a.epi.out_bf16 = b_ref;
        if (launch_gemm(a, BN, EPI_LINEAR, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_f32(o_tc, o_ref, (size_t)rows * N, &ma);
        char nm[64];
        snprintf(nm, 64, "T1 linear+lrelu BN=%d f32", BN);
        report(nm, d, ma, 1e-4);
        d = max_diff_bf16(b_tc, b_ref, (size_t)rows * N, &ma);
        snprintf(nm, 64, "T1 linear+lrelu BN=%d bf16", BN);
        report(nm, d, ma, 1e-2);
        // CPU spot check of the reference kernel itself
        std::vector<float> href((size_t)rows * N);
        CK(cudaMemcpy(href.data(), o_ref, href.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int r : {0, 1, 127, 128, 300, 511})
            for (int c : {0, 1, 31, 32, 127, 128, 255, 256, 500}) {
                double s = hbias[c];
                for (int k = 0; k < K1; ++k) s += (double)hA1[(size_t)r * K1 + k] * hW[(size_t)c * K + k];
                for (int k = 0; k < K2; ++k) s += (double)hA2[(size_t)r * 256 + 64 + k] * hW[(size_t)c * K + K1 + k];
                s = s > 0 ? s : 0.2 * s;
                worst = fmax(worst, fabs(s - href[(size_t)r * N + c]));
            }
        snprintf(nm, 64, "T1 ref-kernel vs CPU BN=%d", BN);
        report(nm, worst, 1.0, 1e-4);
        // dropped columns must stay zero
        double tail = 0;
        std::vector<float> htc((size_t)rows * N);
        CK(cudaMemcpy(htc.data(), o_tc, htc.size() * 4, cudaMemcpyDeviceToHost));
        for (int r = 0; r < rows; ++r)
            for (int c = N - 5; c < N; ++c) tail = fmax(tail, fabs(htc[(size_t)r * N + c]));
        printf("%-44s %s\n", "T1 n_valid tail untouched", tail == 0 ? "OK" : "FAIL");
        if (tail != 0) ++n_fail;
    }

    // ---------------- T2: slot addressing (LEFT/RIGHT parents -> SELF), GN epilogue -----------------
    {
        const int Bp = 256, level = 2, depth = 8, feat = 128;
        const int rows = Bp << level, N = 128, K = 256;
        const size_t slot_rows = (size_t)257 * Bp;
        bf16* lat = dev_bf16(slot_rows * feat, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.1f);
        float* gam = dev_f32(N, 0.2f, 1.f);
        float* bet = dev_f32(N, 0.1f);
        bf16* o_tc = dev_zero<bf16>(slot_rows * feat);
        bf16* o_ref = dev_zero<bf16>(slot_rows * feat);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {lat, feat, 0, 128, ROW_LEFT, 0, 0, {0}};
        a.seg[1] = {lat, feat, 0, 128, ROW_RIGHT, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], lat, slot_rows, feat, feat, 128)) return 2;
        a.a_map[1] = a.a_map[0];
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 128)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {Bp, level, depth};
        a.epi.gn_gamma = gam; a.epi.gn_beta = bet; a.epi.gn_group = 16; a.epi.act = ACT_LRELU; a.epi.n_valid = N;
        a.epi.out_bf16_ld = feat; a.epi.out_bf16_mode = ROW_SELF;
        a.epi.out_bf16 = o_tc;
        if (launch_gemm(a, 128, EPI_GN, false, 0, sms)) return 2;
        a.epi.out_bf16 = o_ref;
        if (launch_gemm(a, 128, EPI_GN, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_bf16(o_tc, o_ref, slot_rows * feat, &ma);
        report("T2 slot LEFT/RIGHT->SELF + GroupNorm(16)", d, ma, 1.6e-2);
    }

    // ---------------- T3: reparametrisation epilogue (BN 256) ---------------------------------------
    {
        const int Bp = 128, B = 100, level = 3, depth = 8;
        const int rows = Bp << level, N = 512, K = 128;
        bf16* A = dev_bf16((size_t)rows * K, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.05f);
        float* bias = dev_f32(N, 0.1f);
        float* z = dev_f32((size_t)B * 255 * 256, 1.f);
        bf16* o_tc = dev_zero<bf16>((size_t)rows * 256);
        bf16* o_ref = dev_zero<bf16>((size_t)rows * 256);
        float* mu_tc = dev_zero<float>((size_t)B * 255 * 256);
        float* ls_tc = dev_zero<float>((size_t)B * 255 * 256);
        float* mu_ref = dev_zero<float>((size_t)B * 255 * 256);
        float* ls_ref = dev_zero<float>((size_t)B * 255 * 256);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 1;
        a.seg[0] = {A, K, 0, K, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], A, rows, K, K, 128)) return 2;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {Bp, level, depth};
        a.epi.bias = bias; a.epi.z = z; a.epi.n_cand = B; a.epi.nz = 256; a.epi.out_bf16_ld = 256;
        a.epi.out_bf16 = o_tc; a.epi.mu_out = mu_tc; a.epi.ls_out = ls_tc;
        if (launch_gemm(a, 256, EPI_REPARAM, false, 0, sms)) return 2;
        a.epi.out_bf16 = o_ref; a.epi.mu_out = mu_ref; a.epi.ls_out = ls_ref;
        if (launch_gemm(a, 256, EPI_REPARAM, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_bf16(o_tc, o_ref, (size_t)rows * 256, &ma);
        report("T3 reparam zeta", d, ma, 1.6e-2);
        d = max_diff_f32(mu_tc, mu_ref, (size_t)B * 255 * 256, &ma);
        report("T3 reparam mu (df layout)", d, ma, 1e-4);
    }

    // ---------------- T4: LSTM epilogue, K = 1024 (x | h), N = 2048, slot hidden write --------------
    {
        const int Bp = 128, level = 1, depth = 8, H = 512;
        const int rows = Bp << level, N = 2048, K = 1024;
        const size_t slot_rows = (size_t)257 * Bp;
        bf16* X = dev_bf16((size_t)rows * H, 1.f);
        bf16* SH = dev_bf16((size_t)rows * 1536, 1.f);
        bf16* SC = dev_bf16((size_t)rows * 1536, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.03f);
        float* bias = dev_f32(N, 0.2f);
        bf16* xn_tc = dev_zero<bf16>((size_t)rows * H);
        bf16* xn_ref = dev_zero<bf16>((size_t)rows * H);
        bf16* hid_tc = dev_zero<bf16>(slot_rows * 3072);
        bf16* hid_ref = dev_zero<bf16>(slot_rows * 3072);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {X, H, 0, H, ROW_LEVEL, 0, 0, {0}};
        a.seg[1] = {SH, 1536, 512, H, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], X, rows, H, H, 128)) return 2;
        if (make_tmap_bf16(&a.a_map[1], SH, rows, 1536, 1536, 128)) return 2;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {Bp, level, depth};
        a.epi.bias = bias; a.epi.c_prev = SC; a.epi.c_prev_ld = 1536; a.epi.c_prev_col0 = 512;
        a.epi.out_bf16_ld = H; a.epi.hid_ld = 3072; a.epi.hid_col0 = 1024; a.epi.hidden = H; a.epi.write_hid = 1;
        a.epi.out_bf16 = xn_tc; a.epi.hid = hid_tc;
        if (launch_gemm(a, 256, EPI_LSTM, false, 0, sms)) return 2;
        a.epi.out_bf16 = xn_ref; a.epi.hid = hid_ref;
        if (launch_gemm(a, 256, EPI_LSTM, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_bf16(xn_tc, xn_ref, (size_t)rows * H, &ma);
        report("T4 lstm h' (x_next)", d, ma, 1.6e-2);
        d = max_diff_bf16(hid_tc, hid_ref, slot_rows * 3072, &ma);
        report("T4 lstm hidden slots (h,c)", d, ma, 1.6e-2);
    }

    // ---------------- T5: grouped A windows + split outputs (the 6 parent-state projections) --------
    {
        const int Bp = 128, level = 1, depth = 8;
        const int rows = Bp << level, N = 3072, K = 1024;
        const size_t slot_rows = (size_t)257 * Bp;
        bf16* hid = dev_bf16(slot_rows * 3072, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.03f);
        float* bias = dev_f32(N, 0.2f);
        bf16* sh_tc = dev_zero<bf16>((size_t)rows * 1536);
        bf16* sh_ref = dev_zero<bf16>((size_t)rows * 1536);
        float* sc_tc = dev_zero<float>((size_t)rows * 1536);
        float* sc_ref = dev_zero<float>((size_t)rows * 1536);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {hid, 3072, 0, 512, ROW_LEFT, 0, 512, {0, 1024, 2048, 512, 1536, 2560}};
        a.seg[1] = {hid, 3072, 0, 512, ROW_RIGHT, 0, 512, {0, 1024, 2048, 512, 1536, 2560}};
        if (make_tmap_bf16(&a.a_map[0], hid, slot_rows, 3072, 3072, 128)) return 2;
        a.a_map[1] = a.a_map[0];
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 128)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {Bp, level, depth};
        a.epi.bias = bias; a.epi.n_valid = N; a.epi.split_col = 1536; a.epi.out_bf16_ld = 1536; a.epi.out_f32_ld = 1536;
        a.epi.out_bf16 = sh_tc; a.epi.out_f32 = sc_tc;
        if (launch_gemm(a, 128, EPI_LINEAR, false, 0, sms)) return 2;
        a.epi.out_bf16 = sh_ref; a.epi.out_f32 = sc_ref;
        if (launch_gemm(a, 128, EPI_LINEAR, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_bf16(sh_tc, sh_ref, (size_t)rows * 1536, &ma);
        report("T5 projections s_h (bf16)", d, ma, 1.6e-2);
        d = max_diff_f32(sc_tc, sc_ref, (size_t)rows * 1536, &ma);
        report("T5 projections s_c (f32)", d, ma, 1e-4);
    }

    // ---------------- T1b: CTA-pair MMA (cta_group::2, M = 256), linear epilogue, 2 K-segments, several tiles per pair -----
    {
        const int rows = 2048, N = 1024, K1 = 128, K2 = 192, K = K1 + K2;
        bf16* A1 = dev_bf16((size_t)rows * K1, 1.f);
        bf16* A2 = dev_bf16((size_t)rows * 256, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.1f);
        float* bias = dev_f32(N, 0.5f);
        float* o_tc = dev_zero<float>((size_t)rows * N);
        float* o_ref = dev_zero<float>((size_t)rows * N);
        bf16* b_tc = dev_zero<bf16>((size_t)rows * N);
        bf16* b_ref = dev_zero<bf16>((size_t)rows * N);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {A1, K1, 0, K1, ROW_LEVEL, 0, 0, {0}};
        a.seg[1] = {A2, 256, 64, K2, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], A1, rows, K1, K1, 128)) return 2;
        if (make_tmap_bf16(&a.a_map[1], A2, rows, 256, 256, 128)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {128, 0, 8};
        a.epi.bias = bias; a.epi.act = ACT_LRELU; a.epi.n_valid = N;
        a.epi.out_f32_ld = N; a.epi.out_bf16_ld = N;
        a.epi.out_f32 = o_tc; a.epi.out_bf16 = b_tc;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 128)) return 2;
        for (int rep = 0; rep < 3; ++rep)            // repeated launches: barrier phases / TMEM reuse across launches
            if (launch_gemm(a, 256, EPI_LINEAR, false, 0, 8 /* few CTAs: 4 pairs, 8 work items each */, 2)) return 2;
        a.epi.out_f32 = o_ref; a.epi.out_bf16 = b_ref;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256)) return 2;
        if (launch_gemm(a, 256, EPI_LINEAR, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_f32(o_tc, o_ref, (size_t)rows * N, &ma);
        report("T1b CTA-pair linear f32 (cluster 2)", d, ma, 1e-4);
        d = max_diff_bf16(b_tc, b_ref, (size_t)rows * N, &ma);
        report("T1b CTA-pair linear bf16 (cluster 2)", d, ma, 1e-2);
    }

    // ---------------- T4b: cluster multicast of the weight tile (clusters of 2, 4, 8 CTAs along M) ---------
    for (int cl : {2, 4, 8}) {
        const int Bp = 1024, level = 0, depth = 8, H = 512;
        const int rows = Bp << level, N = 2048, K = 1024;
        bf16* X = dev_bf16((size_t)rows * H, 1.f);
        bf16* SH = dev_bf16((size_t)rows * H, 1.f);
        bf16* SC = dev_bf16((size_t)rows * H, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.03f);
        float* bias = dev_f32(N, 0.2f);
        bf16* xn_tc = dev_zero<bf16>((size_t)rows * H);
        bf16* xn_ref = dev_zero<bf16>((size_t)rows * H);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {X, H, 0, H, ROW_LEVEL, 0, 0, {0}};
        a.seg[1] = {SH, H, 0, H, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], X, rows, H, H, 128)) return 2;
        if (make_tmap_bf16(&a.a_map[1], SH, rows, H, H, 128)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {Bp, level, depth};
        a.epi.bias = bias; a.epi.c_prev = SC; a.epi.c_prev_ld = H; a.epi.out_bf16_ld = H; a.epi.hidden = H;
        a.epi.out_bf16 = xn_tc;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256 / cl)) return 2;
        if (launch_gemm(a, 256, EPI_LSTM, false, 0, sms, cl)) return 2;
        a.epi.out_bf16 = xn_ref;
        if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256)) return 2;
        if (launch_gemm(a, 256, EPI_LSTM, true, 0, sms)) return 2;
        CK(cudaDeviceSynchronize());
        double ma, d = max_diff_bf16(xn_tc, xn_ref, (size_t)rows * H, &ma);
        char nm[64];
        snprintf(nm, 64, "T4b lstm h' with cluster %d multicast", cl);
        report(nm, d, ma, 1.6e-2);
    }

    // ---------------- T6: throughput of the LSTM-gate GEMM at level-7 size, per cluster size ------------
    {
        const int rows = 131072, N = 2048, K = 1024, H = 512;
        bf16* X = dev_bf16((size_t)rows * H, 1.f);
        bf16* SH = dev_bf16((size_t)rows * H, 1.f);
        bf16* SC = dev_bf16((size_t)rows * H, 1.f);
        bf16* Wd = dev_bf16((size_t)N * K, 0.03f);
        float* bias = dev_f32(N, 0.2f);
        bf16* xn = dev_zero<bf16>((size_t)rows * H);
        GemmArgs a;
        memset(&a, 0, sizeof(a));
        a.n_seg = 2;
        a.seg[0] = {X, H, 0, H, ROW_LEVEL, 0, 0, {0}};
        a.seg[1] = {SH, H, 0, H, ROW_LEVEL, 0, 0, {0}};
        if (make_tmap_bf16(&a.a_map[0], X, rows, H, H, 128)) return 2;
        if (make_tmap_bf16(&a.a_map[1], SH, rows, H, H, 128)) return 2;
        a.w = Wd; a.w_ld = K; a.rows = rows; a.N = N; a.K = K;
        a.g = {1024, 7, 8};
        a.epi.bias = bias; a.epi.c_prev = SC; a.epi.c_prev_ld = H; a.epi.out_bf16_ld = H; a.epi.hidden = H;
        a.epi.out_bf16 = xn;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int cl : {1, 2, 4, 8}) {
            if (make_tmap_bf16(&a.w_map, Wd, N, K, K, 256 / cl)) return 2;
            for (int i = 0; i < 3; ++i) launch_gemm(a, 256, EPI_LSTM, false, 0, sms, cl);
            CK(cudaEventRecord(e0));
            const int iters = 10;
            for (int i = 0; i < iters; ++i) launch_gemm(a, 256, EPI_LSTM, false, 0, sms, cl);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double fl = 2.0 * rows * N * K;
            printf("T6 lstm-gate GEMM %dx%dx%d cluster %d: %.3f ms/launch, %.1f TFLOP/s\n", rows, N, K, cl, ms / iters,
                   fl / (ms / iters * 1e-3) / 1e12);
        }
        // plain linear epilogue (decoder layer 3 shape: K 2048, N 4096, BN 128)
        {
            const int r2 = 65536, N2 = 4096, K2 = 2048;
            bf16* A2 = dev_bf16((size_t)r2 * K2, 1.f);
            bf16* W2 = dev_bf16((size_t)N2 * K2, 0.02f);
            float* b2 = dev_f32(N2, 0.1f);
            bf16* o2 = dev_zero<bf16>((size_t)r2 * N2);
            GemmArgs g;
            memset(&g, 0, sizeof(g));
            g.n_seg = 1;
            g.seg[0] = {A2, K2, 0, K2, ROW_LEVEL, 0, 0, {0}};
            if (make_tmap_bf16(&g.a_map[0], A2, r2, K2, K2, 128)) return 2;
            g.w = W2; g.w_ld = K2; g.rows = r2; g.N = N2; g.K = K2;
            g.g = {1024, 0, 8};
            g.epi.bias = b2; g.epi.act = ACT_RELU; g.epi.n_valid = N2; g.epi.out_bf16 = o2; g.epi.out_bf16_ld = N2;
            for (int bn : {128, 256})
                for (int cl : {1, 2, 4, 8}) {
                    if (make_tmap_bf16(&g.w_map, W2, N2, K2, K2, bn / cl)) return 2;
                    for (int i = 0; i < 2; ++i) launch_gemm(g, bn, EPI_LINEAR, false, 0, sms, cl);
                    CK(cudaEventRecord(e0));
                    for (int i = 0; i < 5; ++i) launch_gemm(g, bn, EPI_LINEAR, false, 0, sms, cl);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("T6 linear GEMM %dx%dx%d BN %d cluster %d: %.3f ms/launch, %.1f TFLOP/s\n", r2, N2, K2, bn, cl, ms / 5,
                           2.0 * r2 * N2 * K2 / (ms / 5 * 1e-3) / 1e12);
                }
        }
        // T7: where do the K <= 1024 linear GEMMs lose time?  Same launch with (a) full epilogue, (b) stores dropped
        // (n_valid = 0: accumulators drained but nothing written), per decoder / projection shape.
        {
            struct Shape { int r, N, K; const char* what; };
            const Shape shapes[] = {{65536, 1024, 128, "dec1"}, {65536, 2048, 1024, "dec2"}, {65536, 4096, 1024, "dec3 banded"},
                                    {131072, 3072, 1024, "proj"}, {131072, 2048, 1024, "lstm-shape linear"}};
            for (const Shape& sh : shapes) {
                bf16* A2 = dev_bf16((size_t)sh.r * sh.K, 1.f);
                bf16* W2 = dev_bf16((size_t)sh.N * sh.K, 0.02f);
                float* b2 = dev_f32(sh.N, 0.1f);
                bf16* o2 = dev_zero<bf16>((size_t)sh.r * sh.N);
                GemmArgs g;
                memset(&g, 0, sizeof(g));
                g.n_seg = 1;
                g.seg[0] = {A2, sh.K, 0, sh.K, ROW_LEVEL, 0, 0, {0}};
                if (make_tmap_bf16(&g.a_map[0], A2, sh.r, sh.K, sh.K, 128)) return 2;
                g.w = W2; g.w_ld = sh.K; g.rows = sh.r; g.N = sh.N; g.K = sh.K;
                g.g = {1024, 0, 8};
                g.epi.bias = b2; g.epi.act = ACT_RELU; g.epi.out_bf16 = o2; g.epi.out_bf16_ld = sh.N;
                if (make_tmap_bf16(&g.w_map, W2, sh.N, sh.K, sh.K, 128)) return 2;
                for (int nv : {sh.N, 0}) {
                    g.epi.n_valid = nv;
                    for (int i = 0; i < 2; ++i) launch_gemm(g, 256, EPI_LINEAR, false, 0, sms, 2);
                    CK(cudaEventRecord(e0));
                    for (int i = 0; i < 5; ++i) launch_gemm(g, 256, EPI_LINEAR, false, 0, sms, 2);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("T7 %-18s %6dx%4dx%4d %s: %.3f ms, %.1f TFLOP/s, out %.2f TB/s\n", sh.what, sh.r, sh.N, sh.K,
                           nv ? "full epilogue" : "no stores    ", ms / 5, 2.0 * sh.r * sh.N * sh.K / (ms / 5 * 1e-3) / 1e12,
                           nv ? 2.0 * sh.r * sh.N / (ms / 5 * 1e-3) / 1e12 : 0.0);
                }
                cudaFree(A2); cudaFree(W2); cudaFree(b2); cudaFree(o2);
            }
        }
    }
    printf(n_fail ? "GEMM_TEST FAILED (%d)\n" : "GEMM_TEST PASSED\n", n_fail);
    return n_fail ? 1 : 0;
}
