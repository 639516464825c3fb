// Stand-alone check + timing of dec_tail3_kernel against a direct CPU convolution (same rounding points).
//   ./dec_tail3_test            correctness on a few images, then timing at 1024 candidates x 16 slots
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../video_gcp_b200/csrc/dec_tail3.cuh"
extern "C" void gcp_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
using namespace gcp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

static float bfr(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint32_t rng_state = 12345u;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f; }

// Z arrays (see dec_tail3.cuh): [ky][h][row = 16 b + co][8], block b holds filter column kx = b - 3
static void pack_z(const std::vector<float>& w /*[16 co][16 ci][4][4]*/, float scale, int n_co, std::vector<bf16>& z) {
    z.assign(D3_W_BYTES / 2, __float2bfloat16(0.f));
    for (int ky = 0; ky < 4; ++ky)
        for (int h = 0; h < 2; ++h)
            for (int b = 3; b < 7; ++b)
                for (int co = 0; co < n_co; ++co)
                    for (int e = 0; e < 8; ++e)
                        z[(size_t)ky * (D3_Z_KY / 2) + h * (D3_Z_CHUNK / 2) + (16 * b + co) * 8 + e] =
                            __float2bfloat16(scale * w[((co * 16 + 8 * h + e) * 4 + ky) * 4 + (b - 3)]);
}

int main(int argc, char** argv) {
    const int Bp = 128, B = 3, ns = 5;
    std::vector<float> w4(16 * 16 * 16), w5(16 * 16 * 16), b5(16, 0.f);
    for (auto& v : w4) v = bfr(frand() * 0.15f);
    for (auto& v : w5) v = bfr(frand() * 0.3f);
    for (int i = 0; i < 15; ++i) b5[i] = frand() * 0.2f;
    std::vector<bf16> z4, z5;
    pack_z(w4, 1.0f, 16, z4);
    pack_z(w5, 0.5f, 15, z5);
    std::vector<float> b5h(16, 0.f);
    for (int i = 0; i < 15; ++i) b5h[i] = 0.5f * b5[i];
    // inputs
    std::vector<bf16> x3((size_t)ns * Bp * 4096), s4((size_t)B * 256 * 64);
    for (auto& v : x3) v = __float2bfloat16(frand());
    for (auto& v : s4) v = __float2bfloat16(frand() * 0.5f);
    bf16 *dx3, *ds4, *dz4, *dz5; float *db5, *dimg; unsigned long long* dprof;
    CK(cudaMalloc(&dx3, x3.size() * 2)); CK(cudaMemcpy(dx3, x3.data(), x3.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ds4, s4.size() * 2)); CK(cudaMemcpy(ds4, s4.data(), s4.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dz4, D3_W_BYTES)); CK(cudaMemcpy(dz4, z4.data(), D3_W_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dz5, D3_W_BYTES)); CK(cudaMemcpy(dz5, z5.data(), D3_W_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&db5, 64)); CK(cudaMemcpy(db5, b5h.data(), 64, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dimg, (size_t)B * 255 * 3072 * 4)); CK(cudaMemset(dimg, 0, (size_t)B * 255 * 3072 * 4));
    CK(cudaMalloc(&dprof, 128)); CK(cudaMemset(dprof, 0, 128));
    CK(cudaFuncSetAttribute(dec_tail3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM_BYTES));
    DecTail3Args a; memset(&a, 0, sizeof(a));
    a.x3 = dx3; a.s4 = ds4; a.s4_stride = 256 * 64; a.w4 = dz4; a.w5 = dz5; a.b5h = db5; a.images = dimg;
    a.Bp = Bp; a.n_cand = B; a.slot0 = 7; a.n_slots = ns; a.n_nodes = 255; a.prof = nullptr;
    for (int grid : {4, 1, 15}) {
        CK(cudaMemset(dimg, 0, (size_t)B * 255 * 3072 * 4));
        dec_tail3_kernel<<<grid, D3_THREADS, D3_SMEM_BYTES>>>(a);
        CK(cudaDeviceSynchronize());
        std::vector<float> img((size_t)B * 255 * 3072);
        CK(cudaMemcpy(img.data(), dimg, img.size() * 4, cudaMemcpyDeviceToHost));
        // CPU reference
        double maxd = 0, maxr = 0;
        std::vector<float> up(16 * 34 * 36), feat(16 * 36 * 36);
        for (int cand = 0; cand < B; ++cand)
            for (int sl = 0; sl < ns; ++sl) {
                const bf16* xr = &x3[((size_t)sl * Bp + cand) * 4096];
                // padded up-sampled input [ci][Y 35][X 35] (bf16-rounded)
                std::vector<float> P(16 * 35 * 35, 0.f), F(16 * 35 * 35, 0.f);
                for (int ci = 0; ci < 16; ++ci)
                    for (int oy = 0; oy < 32; ++oy)
                        for (int ox = 0; ox < 32; ++ox) {
                            int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
                            up2_src(oy, 16, y0, y1, wy0, wy1); up2_src(ox, 16, x0, x1, wx0, wx1);
                            auto at = [&](int y, int x) { return __bfloat162float(xr[((ci >> 3) * 256 + y * 16 + x) * 8 + (ci & 7)]); };
                            // same evaluation order as the kernel: horizontal first, then vertical
                            const float h0 = wx0 * at(y0, x0) + wx1 * at(y0, x1), h1 = wx0 * at(y1, x0) + wx1 * at(y1, x1);
                            P[(ci * 35 + oy + 1) * 35 + ox + 1] = bfr(wy0 * h0 + wy1 * h1);
                        }
                for (int co = 0; co < 16; ++co)
                    for (int oy = 0; oy < 32; ++oy)
                        for (int ox = 0; ox < 32; ++ox) {
                            float s = 0.f;
                            for (int ci = 0; ci < 16; ++ci)
                                for (int ky = 0; ky < 4; ++ky)
                                    for (int kx = 0; kx < 4; ++kx)
                                        s += P[(ci * 35 + oy + ky) * 35 + ox + kx] * w4[((co * 16 + ci) * 4 + ky) * 4 + kx];
                            s += __bfloat162float(s4[((size_t)cand * 256 + oy * 8 + (ox >> 2)) * 64 + (3 - (ox & 3)) * 16 + co]);
                            F[(co * 35 + oy + 1) * 35 + ox + 1] = bfr(tanhf(s));
                        }
                const float* got = &img[((size_t)cand * 255 + a.slot0 + sl - 1) * 3072];
                for (int oy = 0; oy < 32; ++oy)
                    for (int ox = 0; ox < 32; ++ox) {
                        float rgb[3] = {0, 0, 0};
                        for (int co = 0; co < 15; ++co) {
                            float s = b5[co];
                            for (int ci = 0; ci < 16; ++ci)
                                for (int ky = 0; ky < 4; ++ky)
                                    for (int kx = 0; kx < 4; ++kx)
                                        s += F[(ci * 35 + oy + ky) * 35 + ox + kx] * w5[((co * 16 + ci) * 4 + ky) * 4 + kx];
                            rgb[co % 3] += 1.f / (1.f + expf(-s));
                        }
                        for (int k = 0; k < 3; ++k) {
                            const float ref = rgb[k] * 0.4f - 1.f;
                            const double dd = fabs(ref - got[k * 1024 + oy * 32 + ox]);
                            if (dd > maxd) maxd = dd;
                            if (fabs(ref) > maxr) maxr = fabs(ref);
                        }
                    }
            }
        // untouched nodes must still be zero
        double other = 0;
        for (int cand = 0; cand < B; ++cand)
            for (int node = 0; node < 255; ++node)
                if (node < a.slot0 - 1 || node >= a.slot0 - 1 + ns)
                    for (int i = 0; i < 3072; i += 97) other += fabs(img[((size_t)cand * 255 + node) * 3072 + i]);
        printf("grid %3d: max |image - ref| = %.3e (max |ref| %.3f), stray writes %.1f  %s\n", grid, maxd, maxr, other,
               (maxd < 4e-3 && other == 0) ? "OK" : "FAIL");
    }
    // ---- timing
    {
        const int Bt = 1024, nst = 16;
        bf16* bx3; float* bimg;
        CK(cudaMalloc(&bx3, (size_t)nst * Bt * 4096 * 2)); CK(cudaMemset(bx3, 0, (size_t)nst * Bt * 4096 * 2));
        CK(cudaMalloc(&bimg, (size_t)Bt * 255 * 3072 * 4));
        DecTail3Args t = a;
        t.x3 = bx3; t.images = bimg; t.Bp = Bt; t.n_cand = Bt; t.slot0 = 1; t.n_slots = nst; t.s4_stride = 0; t.prof = dprof;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        dec_tail3_kernel<<<148, D3_THREADS, D3_SMEM_BYTES>>>(t);
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(dprof, 0, 128));
        CK(cudaEventRecord(e0));
        dec_tail3_kernel<<<148, D3_THREADS, D3_SMEM_BYTES>>>(t);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long h[16]; CK(cudaMemcpy(h, dprof, 128, cudaMemcpyDeviceToHost));
        const double n = (double)h[10];
        printf("v3: %.3f ms for %d images (%.2f us/image/SM) -> full step (261120 images) %.2f ms; per-image cycles:\n", ms, nst * Bt,
               ms * 1e3 / (nst * Bt / 148.0), ms * 261120.0 / (nst * Bt));
        printf("  builder: wait %.0f work %.0f\n", h[0] / n, h[1] / n);
        printf("  mma4: wait_in4_full %.0f wait_d_empty %.0f total %.0f | mma5: wait_feat %.0f wait_d_empty %.0f total %.0f\n", h[2] / n, h[3] / n, h[4] / n, h[13] / n, h[14] / n, h[15] / n);
        printf("  epi1: tile0 wait %.0f work %.0f | tile1 wait %.0f work %.0f\n", h[6] / n, h[7] / n, h[11] / n, h[12] / n);
        printf("  epi2: wait %.0f work %.0f   (images %.0f)\n", h[8] / n, h[9] / n, n);
    }
    return 0;
}
