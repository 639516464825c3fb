// Timing of the decoder-tail kernel with per-role cycle counters (wait vs work per warp role).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../video_gcp_b200/csrc/dec_tail2.cuh"
extern "C" void gcp_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
using namespace gcp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
int main() {
    const int Bp = 1024, B = 1024, ns = 16;
    bf16 *x3, *skip, *w4, *w5; float *b4, *b5, *img; unsigned long long* prof;
    CK(cudaMalloc(&x3, (size_t)ns * Bp * 4096 * 2)); CK(cudaMemset(x3, 0, (size_t)ns * Bp * 4096 * 2));
    CK(cudaMalloc(&skip, 2 * DT_PSTRIDE * 8 * 2)); CK(cudaMemset(skip, 0, 2 * DT_PSTRIDE * 8 * 2));
    CK(cudaMalloc(&w4, DT_W4_BYTES)); CK(cudaMemset(w4, 0, DT_W4_BYTES));
    CK(cudaMalloc(&w5, DT_W5_BYTES)); CK(cudaMemset(w5, 0, DT_W5_BYTES));
    CK(cudaMalloc(&b4, 64)); CK(cudaMemset(b4, 0, 64)); CK(cudaMalloc(&b5, 128)); CK(cudaMemset(b5, 0, 128));
    CK(cudaMalloc(&img, (size_t)B * 255 * 3072 * 4)); CK(cudaMalloc(&prof, 64));
    {   // ---- v2
        bf16 *s4, *w4m, *w5m;
        CK(cudaMalloc(&s4, 1152 * 16 * 2)); CK(cudaMemset(s4, 0, 1152 * 16 * 2));
        CK(cudaMalloc(&w4m, D2_W4_BYTES)); CK(cudaMemset(w4m, 0, D2_W4_BYTES));
        CK(cudaMalloc(&w5m, D2_W5_BYTES)); CK(cudaMemset(w5m, 0, D2_W5_BYTES));
        unsigned long long* prof2; CK(cudaMalloc(&prof2, 128));
        CK(cudaFuncSetAttribute(dec_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D2_SMEM_BYTES));
        DecTail2Args a; memset(&a, 0, sizeof(a));
        a.x3 = x3; a.s4 = s4; a.s4_stride = 0; a.w4 = w4m; a.w5 = w5m; a.b5 = b5; a.images = img;
        a.Bp = Bp; a.n_cand = B; a.slot0 = 1; a.n_slots = ns; a.n_nodes = 255; a.slots_per_unit = ns; a.prof = prof2;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        dec_tail2_kernel<<<148, D2_THREADS, D2_SMEM_BYTES>>>(a);
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(prof2, 0, 128));
        CK(cudaEventRecord(e0));
        dec_tail2_kernel<<<148, D2_THREADS, D2_SMEM_BYTES>>>(a);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long h[16]; CK(cudaMemcpy(h, prof2, 128, cudaMemcpyDeviceToHost));
        const double n = (double)h[13];
        printf("v2: %.3f ms for %d images (%.2f us/image/SM), per-image cycles:\n", ms, ns * B, ms * 1e3 / (ns * B / 148.0));
        printf("  builder: wait_x3 %.0f wait_in4_empty %.0f work %.0f\n", h[0] / n, h[1] / n, h[2] / n);
        printf("  mma: wait_in4_full %.0f wait_d4_empty %.0f wait_feat %.0f wait_d5_empty %.0f total %.0f\n", h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n);
        printf("  epi1: wait_d4_full %.0f work %.0f (of which barrier %.0f)\n", h[8] / n, h[9] / n, h[10] / n);
        printf("  epi2: wait_d5_full %.0f work %.0f   (images %.0f)\n", h[11] / n, h[12] / n, n);
    }
    return 0;
}
