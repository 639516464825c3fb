// PCIe read rate of upload_rows_kernel (zero-copy gather from pinned host memory) vs cudaMemcpyAsync.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "../../video_gcp_b200/csrc/kernels_misc.cuh"
extern "C" void gcp_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
using namespace gcp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
int main() {
    const int B = 1024, n_nodes = 255, row4 = 64;
    const size_t bytes = (size_t)B * n_nodes * row4 * 16;
    float4 *h, *d;
    CK(cudaMallocHost(&h, bytes)); memset(h, 1, bytes);
    CK(cudaMalloc(&d, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("cudaMemcpyAsync %.0f MB: %.2f ms, %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
    }
    for (int grid : {16, 32, 64, 148, 296})
        for (int threads : {128, 256}) {
            upload_rows_kernel<<<grid, threads>>>(h, d, B, n_nodes, row4, 2, 0, 128);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            upload_rows_kernel<<<grid, threads>>>(h, d, B, n_nodes, row4, 2, 0, 128);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double mb = (double)B * 128 * row4 * 16 / 1e6;
            printf("upload_rows_kernel level 7 (%.0f MB) grid %3d x %3d: %.2f ms, %.1f GB/s\n", mb, grid, threads, ms, mb / ms);
        }
    return 0;
}
