// Decoder tail: the two full-resolution convolutions of the GCP decoder, fused per node image.
// (shared definitions + the SIMT verification kernel; the tcgen05 product kernel is in dec_tail3.cuh)
//
//   x3 [16ch,16x16] --bilinear x2--> cat with up(skip s0 [16ch,16x16]) --ZeroPad(1,2,1,2)--> conv k4 (32->16)
//   + bias, tanh = feat [16,32,32] --ZeroPad(1,2,1,2)--> conv k4 (16->30) + bias --> sigmoid on the 15 mixture
//   means, mean over the 5 mixtures, *2-1 = image [3,32,32]
//   (blox/torch/encoder_decoder.py:56-97,150-218; blox/torch/layers.py:128-150; blox/torch/dist.py:200-201)
//
// B200 mapping: both convolutions run on tcgen05 as implicit GEMMs with M = output pixels, K = (tap,
// input channel), N = output channels.  The padded input lives in shared memory as 8-channel planes
// [C/8][pixel][8] (bf16), which is exactly the un-swizzled K-major UMMA core-matrix layout when 8
// consecutive pixels form a core matrix.  Because output pixel p (indexed on the PADDED width 35) reads
// input pixel p + ky*35 + kx, every filter tap is the same smem operand with a shifted start address:
// no im2col, no data movement between taps.  Columns x >= 32 of each output row are dead rows of the
// GEMM (8.6 % waste).  Accumulators: 9 M-tiles x 16 (+ 9 x 32) TMEM columns.  The intermediate feature
// map never leaves the SM: the first epilogue writes tanh(feat) as bf16 straight into the second conv's
// operand planes.
#pragma once
#include "common.cuh"

namespace gcp {

constexpr int DT_WP = 35;                  // padded width/height
constexpr int DT_NPIX = 32 * DT_WP;        // 1120 output indices (rows 0..31, padded columns)
constexpr int DT_TILES = 9;                // ceil(1120 / 128)
constexpr int DT_PSTRIDE = 1272;           // pixels per 8-channel plane (>= 9*128 + 3*35 + 3, mult of 8)
constexpr int DT_PLANE_BYTES = DT_PSTRIDE * 16;
constexpr int DT_THREADS = 256;
constexpr int DT_C4_IN = 32, DT_C4_OUT = 16, DT_C5_IN = 16, DT_C5_OUT = 32;  // 30 padded to 32
constexpr int DT_W4_BYTES = 16 * 2 * 512;  // [tap][kstep][kchunk 2][n 16][8]
constexpr int DT_W5_BYTES = 16 * 1024;     // [tap][kchunk 2][n 32][8]
constexpr int DT_SMEM_BYTES = 6 * DT_PLANE_BYTES + DT_W4_BYTES + DT_W5_BYTES + 256 + 128;

struct DecTailArgs {
    const bf16* x3;        // [n_slots * Bp][4096] rows = (slot_local, cand); layout [plane 2][y16][x16][8]
    const bf16* skip_up;   // [Bp][2][DT_PSTRIDE][8]  up-sampled + padded skip s0, per candidate
    int skip_stride;       // elements between candidates in skip_up (0: all candidates share one skip)
    const bf16* w4;        // packed, DT_W4_BYTES
    const bf16* w5;        // packed, DT_W5_BYTES
    const float* b4;       // [16]
    const float* b5;       // [32]
    float* images;         // [B][n_nodes][3][32][32]
    int Bp, n_cand;        // padded / valid candidates
    int slot0, n_slots;    // this launch decodes slots slot0 .. slot0+n_slots-1 (slot = node + 1)
    int n_nodes;           // 255
    int slots_per_unit;    // work unit = (candidate, run of slots)
    // head = 1: pixel-copy head (PixelCopyDecoder, blox/torch/encoder_decoder.py:235-259): w5/b5 rows 0-2 = gen_head,
    // 3-5 = mask_head; image = softmax(mask)_0 I_0 + softmax(mask)_1 I_g + softmax(mask)_2 tanh(gen)
    int head;
    const float* src0;     // [n_cand or 1][3][32][32] start image
    const float* srcg;     // goal image
    int src_stride;        // elements between candidates (0: shared)
};

// bilinear x2 (align_corners=False) source rows/weights for output index o of an n-long input
__host__ __device__ __forceinline__ void up2_src(int o, int n, int& i0, int& i1, float& w0, float& w1) {
    const int i = o >> 1;
    if (o & 1) { i0 = i; i1 = min(i + 1, n - 1); w0 = 0.75f; w1 = 0.25f; }
    else       { i0 = max(i - 1, 0); i1 = i;      w0 = 0.25f; w1 = 0.75f; }
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// up-sample the 16x16x16 node feature map into planes 0,1 of in4 (interior of the padded image)
__device__ __forceinline__ void dt_build_up(uint8_t* in4, const bf16* x3row, int tid, int nthreads) {
    const uint4* src = reinterpret_cast<const uint4*>(x3row);  // [plane][y][x] of 16-byte channel groups
    for (int it = tid; it < 2 * 1024; it += nthreads) {
        const int plane = it >> 10, oy = (it >> 5) & 31, ox = it & 31;
        int y0, y1, x0, x1;
        float wy0, wy1, wx0, wx1;
        up2_src(oy, 16, y0, y1, wy0, wy1);
        up2_src(ox, 16, x0, x1, wx0, wx1);
        float a[8], b[8], c[8], d[8];
        unpack8(__ldg(src + plane * 256 + y0 * 16 + x0), a);
        unpack8(__ldg(src + plane * 256 + y0 * 16 + x1), b);
        unpack8(__ldg(src + plane * 256 + y1 * 16 + x0), c);
        unpack8(__ldg(src + plane * 256 + y1 * 16 + x1), d);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = wy0 * (wx0 * a[k] + wx1 * b[k]) + wy1 * (wx0 * c[k] + wx1 * d[k]);
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(in4 + plane * DT_PLANE_BYTES + ((oy + 1) * DT_WP + ox + 1) * 16) = u;
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT verification kernel: same inputs / outputs / rounding points, direct convolution.
// w4p [16][32][4][4], w5p [32][16][4][4] are the plain-layout bf16 weights.
// ---------------------------------------------------------------------------------------------
#ifdef GCPB200_VERIFY
__global__ void __launch_bounds__(256) dec_tail_ref_kernel(const __grid_constant__ DecTailArgs a, const bf16* w4p,
                                                           const bf16* w5p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* in4 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* in5 = in4 + 4 * DT_PLANE_BYTES;
    const int tid = threadIdx.x;
    const int sl = blockIdx.x / a.n_cand, cand = blockIdx.x % a.n_cand;
    for (int i = tid; i < 6 * DT_PLANE_BYTES / 16; i += 256) reinterpret_cast<uint4*>(in4)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.skip_up + (size_t)cand * a.skip_stride);
        uint4* dst = reinterpret_cast<uint4*>(in4 + 2 * DT_PLANE_BYTES);
        for (int i = tid; i < 2 * DT_PSTRIDE; i += 256) dst[i] = __ldg(src + i);
    }
    dt_build_up(in4, a.x3 + ((size_t)sl * a.Bp + cand) * 4096, tid, 256);
    __syncthreads();
    auto in_at = [](const uint8_t* base, int ch, int pix) {
        return __bfloat162float(reinterpret_cast<const bf16*>(base + (ch >> 3) * DT_PLANE_BYTES + pix * 16)[ch & 7]);
    };
    for (int p = tid; p < DT_NPIX; p += 256) {
        const int x = p % DT_WP;
        if (x >= 32) continue;
        for (int co = 0; co < 16; ++co) {
            float s = 0.f;
            for (int ci = 0; ci < 32; ++ci)
                for (int tap = 0; tap < 16; ++tap)
                    s = fmaf(in_at(in4, ci, p + (tap >> 2) * DT_WP + (tap & 3)),
                             __bfloat162float(w4p[(co * 32 + ci) * 16 + tap]), s);
            s = tanhf_(s + a.b4[co]);
            reinterpret_cast<bf16*>(in5 + (co >> 3) * DT_PLANE_BYTES + (p + DT_WP + 1) * 16)[co & 7] = __float2bfloat16_rn(s);
        }
    }
    __syncthreads();
    const int node = a.slot0 + sl - 1;
    float* img = a.images + ((size_t)cand * a.n_nodes + node) * 3072;
    for (int p = tid; p < DT_NPIX; p += 256) {
        const int y = p / DT_WP, x = p - y * DT_WP;
        if (x >= 32) continue;
        float rgb[3] = {0.f, 0.f, 0.f};
        float v6[6];
        for (int co = 0; co < (a.head ? 6 : 15); ++co) {
            float s = 0.f;
            for (int ci = 0; ci < 16; ++ci)
                for (int tap = 0; tap < 16; ++tap)
                    s = fmaf(in_at(in5, ci, p + (tap >> 2) * DT_WP + (tap & 3)),
                             __bfloat162float(w5p[(co * 16 + ci) * 16 + tap]), s);
            if (a.head) v6[co] = s + a.b5[co];
            else rgb[co % 3] += sigmoidf_(s + a.b5[co]);
        }
        if (a.head) {
            const float mx = fmaxf(v6[3], fmaxf(v6[4], v6[5]));
            const float e0 = __expf(v6[3] - mx), e1 = __expf(v6[4] - mx), e2 = __expf(v6[5] - mx);
            const float inv = 1.0f / (e0 + e1 + e2);
            const float* s0 = a.src0 + (size_t)cand * a.src_stride;
            const float* sg = a.srcg + (size_t)cand * a.src_stride;
            for (int k = 0; k < 3; ++k)
                img[k * 1024 + y * 32 + x] = (e0 * s0[k * 1024 + y * 32 + x] + e1 * sg[k * 1024 + y * 32 + x] + e2 * tanhf_(v6[k])) * inv;
            continue;
        }
        for (int k = 0; k < 3; ++k) img[k * 1024 + y * 32 + x] = rgb[k] * 0.4f - 1.0f;
    }
}
#endif

// per-candidate skip preparation: s0 [Bp][16][16][16] fp32 (NCHW) -> up-sampled, padded, bf16 planes
__global__ void skip_prep_kernel(const float* __restrict__ s0, bf16* __restrict__ skip_up, int n_cand) {
    const int cand = blockIdx.x;
    if (cand >= n_cand) return;
    bf16* dst = skip_up + (size_t)cand * 2 * DT_PSTRIDE * 8;
    const float* src = s0 + (size_t)cand * 16 * 256;
    for (int i = threadIdx.x; i < 2 * DT_PSTRIDE * 8; i += blockDim.x) {
        const int c8 = i & 7, pix = (i >> 3) % DT_PSTRIDE, plane = (i >> 3) / DT_PSTRIDE;
        const int py = pix / DT_WP, px = pix - py * DT_WP;
        float v = 0.f;
        if (pix < DT_WP * DT_WP && py >= 1 && py <= 32 && px >= 1 && px <= 32) {
            int y0, y1, x0, x1;
            float wy0, wy1, wx0, wx1;
            up2_src(py - 1, 16, y0, y1, wy0, wy1);
            up2_src(px - 1, 16, x0, x1, wx0, wx1);
            const float* ch = src + (plane * 8 + c8) * 256;
            v = wy0 * (wx0 * ch[y0 * 16 + x0] + wx1 * ch[y0 * 16 + x1]) +
                wy1 * (wx0 * ch[y1 * 16 + x0] + wx1 * ch[y1 * 16 + x1]);
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace gcp
