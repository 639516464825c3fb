// Micro-benchmarks that size the decoder-tail design: tcgen05.mma issue rate vs N and operand layout,
// tcgen05.ld drain rate, warp-shuffle rate.  One CTA per SM; cycles from clock64 on the issuing thread.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../video_gcp_b200/csrc/common.cuh"
extern "C" void gcp_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
using namespace gcp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

// mode 0: no-swizzle K-major (A: 128 rows, core matrices 128 B, K halves `lbo` apart), mode 1: SW128
__global__ void __launch_bounds__(160, 1) umma_rate_kernel(int N, int mode, int reps, int a_shift_bytes, long long* out, int M = 128, int ndst = 2, int sbo = 128) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(M, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t ao = a0 + (r & 3) * a_shift_bytes;
            uint64_t da, db;
            if (mode == 0) { da = umma_desc_nosw(ao, 20352, sbo); db = umma_desc_nosw(b0, N * 16, 128); }
            else { da = umma_desc_sw128(ao); db = umma_desc_sw128(b0); }
            umma_bf16(tmem + (r & (ndst - 1)) * 64, da, db, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


__global__ void __launch_bounds__(160, 1) umma_rate2_kernel(int N, int reps, long long* out, int same_d) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0) tmem_alloc(&holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
        const uint64_t da0 = umma_desc_nosw(a0, 20352, 144), db0 = umma_desc_nosw(b0, N * 16, 128);
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint64_t da = da0 + (uint64_t)(u * 9);          // start address += 144 B
                const uint64_t db = db0 + (uint64_t)(u * 16);
                const uint32_t d = tmem + (same_d ? 0 : (u & 3) * 64);
                if (elect_one()) umma_bf16(d, da, db, idesc, 1);
            }
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// nwarps warps each drain `reps` x (32 lanes x 32 columns)
__global__ void __launch_bounds__(512, 1) ldtm_rate_kernel(int reps, int x16, long long* out, float* sink) {
    __shared__ uint32_t holder;
    if (threadIdx.x < 32) tmem_alloc(&holder, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = holder;
    const int warp = threadIdx.x >> 5;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((r * 32) & 255);
        if (x16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(ta) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
        } else {
            float v[32];
            tmem_ld32(ta, v);
            acc += v[0] + v[31];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

__global__ void __launch_bounds__(512, 1) shfl_rate_kernel(int reps, long long* out, float* sink) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.5f + i;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[(i + 1) & 15], 1);
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(512, 1) mufu_rate_kernel(int reps, long long* out, float* sink) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i * 0.01f;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v[i])); v[i] = y + 0.01f; }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) sink[0] = s;
}

int main() {
    long long* out; float* sink;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int reps = 4096;
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {16, 32, 64, 128, 256})
            for (int shift : {0, 16, 560}) {
                if (mode == 1 && shift == 16) continue;
                umma_rate_kernel<<<148, 160, 100 * 1024>>>(N, mode, reps, mode == 1 ? (shift ? 1024 : 0) : shift, out);
                CK(cudaDeviceSynchronize());
                long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
                printf("umma M=128 N=%3d K=16 %s a_shift=%4d: %.1f cycles/MMA\n", N, mode ? "sw128" : "nosw ", shift, (double)h / reps);
            }
    CK(cudaFuncSetAttribute(umma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int N : {16, 32, 64, 96, 128, 256})
        for (int same : {0, 1}) {
            if (N == 256 && !same) continue;
            umma_rate2_kernel<<<148, 160, 100 * 1024>>>(N, reps, out, same);
            CK(cudaDeviceSynchronize());
            long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
            printf("tight issue loop: umma M=128 N=%3d K=16, %s accumulator: %.1f cycles/MMA\n", N, same ? "one" : "4 rotating", (double)h / reps);
        }
    for (int nd : {1, 2, 4})
        for (int sbo : {128, 144}) {
            umma_rate_kernel<<<148, 160, 100 * 1024>>>(64, 0, reps, 16, out, 128, nd, sbo);
            CK(cudaDeviceSynchronize());
            long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
            printf("umma M=128 N= 64 K=16 nosw, %d rotating accumulators, SBO %d: %.1f cycles/MMA\n", nd, sbo, (double)h / reps);
        }
    for (int N : {16, 64, 128}) {
        umma_rate_kernel<<<148, 160, 100 * 1024>>>(N, 0, reps, 16, out, 64);
        CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        printf("umma M= 64 N=%3d K=16 nosw: %.1f cycles/MMA\n", N, (double)h / reps);
    }
    for (int nw : {4, 8, 16})
        for (int x16 = 0; x16 < 2; ++x16) {
            ldtm_rate_kernel<<<148, nw * 32, 0>>>(2048, x16, out, sink);
            CK(cudaDeviceSynchronize());
            long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
            printf("tcgen05.ld 32x32b.x%d + wait, %2d warps: %.1f cycles per load per warp -> %.1f B/cycle/SM\n", x16 ? 16 : 32, nw,
                   (double)h / 2048, (double)nw * 32 * (x16 ? 16 : 32) * 4 * 2048 / h);
        }
    for (int nw : {4, 8, 16}) {
        shfl_rate_kernel<<<148, nw * 32, 0>>>(1024, out, sink);
        CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        printf("shfl.down (dependent-free x16), %2d warps: %.2f cycles per shfl per warp -> %.2f warp-shfl/cycle/SM\n", nw, (double)h / (1024 * 16), (double)nw * 1024 * 16 / h);
        mufu_rate_kernel<<<148, nw * 32, 0>>>(1024, out, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        printf("tanh.approx x16, %2d warps: %.2f cycles per op per warp -> %.2f lane-ops/cycle/SM\n", nw, (double)h / (1024 * 16), (double)nw * 32 * 1024 * 16 / h);
    }
    return 0;
}
