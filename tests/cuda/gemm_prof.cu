// One GEMM shape, a few launches: the target of `ncu --set full` captures of gemm_tc_kernel.
//   ./gemm_prof rows N K [n_valid]     (linear epilogue, ReLU, bf16 output, BN 256, cluster 2)
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../video_gcp_b200/csrc/gemm_host.cuh"
extern "C" void gcp_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
using namespace gcp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
int main(int argc, char** argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 65536, N = argc > 2 ? atoi(argv[2]) : 2048, K = argc > 3 ? atoi(argv[3]) : 1024;
    const int nv = argc > 4 ? atoi(argv[4]) : N;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    bf16 *A, *W, *O; float* b;
    CK(cudaMalloc(&A, (size_t)rows * K * 2)); CK(cudaMemset(A, 0x11, (size_t)rows * K * 2));
    CK(cudaMalloc(&W, (size_t)N * K * 2)); CK(cudaMemset(W, 0x11, (size_t)N * K * 2));
    CK(cudaMalloc(&O, (size_t)rows * N * 2)); CK(cudaMalloc(&b, N * 4)); CK(cudaMemset(b, 0, N * 4));
    GemmArgs g; memset(&g, 0, sizeof(g));
    g.n_seg = 1; g.seg[0].ptr = A; g.seg[0].ld = K; g.seg[0].k_len = K; g.seg[0].row_mode = ROW_LEVEL;
    if (make_tmap_bf16(&g.a_map[0], A, rows, K, K, 128)) return 2;
    if (make_tmap_bf16(&g.w_map, W, N, K, K, 128)) return 2;
    g.w = W; g.w_ld = K; g.rows = rows; g.N = N; g.K = K; g.g = {1024, 0, 8};
    g.epi.bias = b; g.epi.act = ACT_RELU; g.epi.out_bf16 = O; g.epi.out_bf16_ld = N; g.epi.n_valid = nv;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch_gemm(g, 256, EPI_LINEAR, false, 0, sms, 2);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) launch_gemm(g, 256, EPI_LINEAR, false, 0, sms, 2);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("gemm %dx%dx%d n_valid %d: %.3f ms, %.1f TFLOP/s\n", rows, N, K, nv, ms / 3, 2.0 * rows * N * K / (ms / 3 * 1e-3) / 1e12);
    return 0;
}
