// Kernels of the training-phase forward + loss (gcpb200_forward_loss; SURVEY 8(f)-2, BASELINE config 1):
// batch-statistic BatchNorm encoders and decoder, the conv-1d inference encoder, frame <-> node matching,
// the fused decoder-tail + discretised-logistic-mixture NLL, KL and the scalar loss reductions.
// Reference: gcp/prediction/models/base_gcp.py:140-304, tree/tree_module.py:67-157, tree/inference.py:16-41,
// tree/frame_binding.py:42-100, blox/torch/encoder_decoder.py:15-232, blox/torch/dist.py:87-197,249-252,
// blox/torch/losses.py:20-140, blox/torch/subnetworks.py:120-147.
//
// Batch 16 x 200 frames is a small problem for a B200 (0.3 TFLOP): everything except the TreeLSTM / row-MLP /
// decoder GEMMs (tcgen05, gemm.cuh) is fp32 SIMT, sized so that no stage is more than a few hundred microseconds,
// with fp64 accumulation for every statistic and loss sum (the reference reduces in fp32 on the CPU; fp64 here keeps
// the atomics order-independent to ~1e-15 so results are reproducible run to run).
#pragma once
#include "common.cuh"
#include "dec_tail.cuh"
#include "kernels_misc.cuh"

namespace gcp {

constexpr int TR_T = 200;   // max_seq_len of the 25-room dataset

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the block; result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum_d(double v, double* sh /*[32]*/) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

// ---------------------------------------------------------------------------------------------
// Conv encoder with batch-statistic BatchNorm (nn.BatchNorm2d in training mode, biased variance).
// The reference runs the encoder three times per step (all frames of the batch, the start images, the goal images:
// base_gcp.py:184-209), each call with its own statistics: three "groups" of one launch here.
// Pass A: conv0 + LReLU (skip s0 of the start images), conv1 raw + channel sums.   y1 [n][32*8*8]
// Pass B: BN(stats1) + LReLU, conv2 raw + channel sums.                             y2 [n][64*4*4]
// Pass C: BN(stats2) + LReLU (skip s2 of the start images), 4x4 head -> latent.
// fp32 SIMT direct convolutions (the latents feed every later stage, so they keep the reference's fp32 arithmetic),
// register-tiled: a thread owns 8-16 output channels of one output pixel, input maps sit zero-bordered in shared
// memory (no bounds tests in the inner loops), weights are staged k-major so that one broadcast LDS.128 feeds four
// FMAs; CTAs are persistent over images so the weight staging is paid once.  B = 16 (3232 images): 3.4 ms -> see
// profiles/r1k_train_forward.txt.
// ---------------------------------------------------------------------------------------------
struct EncTrainArgs {
    const float* img[3];    // group 0: frames [n0][3,32,32]; 1: I_0 [n1]; 2: I_g [n2]
    int n[3];
    EncoderWeights W;       // sc*/sh* unused
    const float *w1t, *w2t, *w3t;     // k-major copies: [16*16][32], [32*16][64], [64*16][128]  (k = ci*16 + ky*4 + kx)
    const float *g1, *b1, *g2, *b2;   // BatchNorm affine of pyramid-0 (32) / pyramid-1 (64)
    float *y1, *y2;
    double *st1, *st2;      // [3][32][2], [3][64][2]  (sum, sum of squares), zeroed by the caller
    float* enc_seq;         // group 0 latents [n0][128]
    float* lat_f32;         // groups 1 / 2 -> rows row0_a + i / row0_b + i (fp32 + bf16 copies)
    bf16* lat_bf16;
    int row0_a, row0_b;
    float *skip0, *skip2;   // group 1 only
    bf16* skip2_bf16;
};
__device__ __forceinline__ void enc_group(const EncTrainArgs& a, int i, int& g, int& li) {
    if (i < a.n[0]) { g = 0; li = i; }
    else if (i < a.n[0] + a.n[1]) { g = 1; li = i - a.n[0]; }
    else { g = 2; li = i - a.n[0] - a.n[1]; }
}
// scale / shift of a training-mode BatchNorm channel from (sum, sumsq) over `count` values
__device__ __forceinline__ void bn_coef(const double* st, double count, float gamma, float beta, float& sc, float& sh) {
    const double mean = st[0] / count;
    double var = st[1] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double s = (double)gamma / sqrt(var + 1e-5);
    sc = (float)s;
    sh = (float)((double)beta - mean * s);
}
__device__ __forceinline__ void fma4x(float (&acc)[16], int j0, float v, const float4 w) {
    acc[j0 + 0] = fmaf(v, w.x, acc[j0 + 0]);
    acc[j0 + 1] = fmaf(v, w.y, acc[j0 + 1]);
    acc[j0 + 2] = fmaf(v, w.z, acc[j0 + 2]);
    acc[j0 + 3] = fmaf(v, w.w, acc[j0 + 3]);
}

constexpr int ENC_T = 128;                                   // threads of the three passes
constexpr int ENCA_IMG = 3 * 34 * 34, ENCA_A1 = 16 * 18 * 18, ENCA_W0 = 48 * 16, ENCA_W1 = 256 * 32;
constexpr int ENCA_SMEM = (ENCA_IMG + ENCA_A1 + ENCA_W0 + ENCA_W1) * 4;           // 70 448 B
__global__ void __launch_bounds__(ENC_T) enc_train_a_kernel(const EncTrainArgs a) {
    extern __shared__ float enc_sm[];
    float* img = enc_sm;               // [3][34][34]  zero border
    float* a1 = img + ENCA_IMG;        // [16][18][18] zero border
    float* w0s = a1 + ENCA_A1;         // [k 48][co 16]
    float* w1s = w0s + ENCA_W0;        // [k 256][co 32]
    __shared__ double ssum[3][32], ssq[3][32];
    const int tid = threadIdx.x;
    for (int i = tid; i < ENCA_IMG + ENCA_A1; i += ENC_T) enc_sm[i] = 0.f;
    for (int i = tid; i < ENCA_W0; i += ENC_T) w0s[(i % 48) * 16 + i / 48] = a.W.w0[i];
    for (int i = tid; i < ENCA_W1; i += ENC_T) w1s[i] = a.w1t[i];
    if (tid < 96) (&ssum[0][0])[tid] = (&ssq[0][0])[tid] = 0.0;
    const int n_total = a.n[0] + a.n[1] + a.n[2];
    for (int im = blockIdx.x; im < n_total; im += gridDim.x) {
        int g, li;
        enc_group(a, im, g, li);
        __syncthreads();               // staging done / previous image consumed
        const float* src = a.img[g] + (size_t)li * 3072;
        for (int k = tid; k < 3072; k += ENC_T) img[(k >> 10) * 1156 + (((k >> 5) & 31) + 1) * 34 + (k & 31) + 1] = src[k];
        __syncthreads();
        // conv0 3 -> 16, 32x32 -> 16x16, + bias, LReLU: thread = output pixels (oy, ox) and (oy + 8, ox), 16 channels
        {
            const int oy = tid >> 4, ox = tid & 15;
            float acc0[16], acc1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc0[j] = acc1[j] = __ldg(a.W.b0 + j);
#pragma unroll 1
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int ky = t >> 2, kx = t & 3;
                    const float v0 = img[ci * 1156 + (2 * oy + ky) * 34 + 2 * ox + kx];
                    const float v1 = img[ci * 1156 + (2 * oy + 16 + ky) * 34 + 2 * ox + kx];
                    const float4* w = reinterpret_cast<const float4*>(w0s + (ci * 16 + t) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ww = w[q];
                        fma4x(acc0, 4 * q, v0, ww);
                        fma4x(acc1, 4 * q, v1, ww);
                    }
                }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float r0 = lrelu_(acc0[j]), r1 = lrelu_(acc1[j]);
                a1[j * 324 + (oy + 1) * 18 + ox + 1] = r0;
                a1[j * 324 + (oy + 9) * 18 + ox + 1] = r1;
                if (g == 1) {
                    a.skip0[(size_t)li * 4096 + j * 256 + oy * 16 + ox] = r0;
                    a.skip0[(size_t)li * 4096 + j * 256 + (oy + 8) * 16 + ox] = r1;
                }
            }
        }
        __syncthreads();
        // conv1 16 -> 32, 16x16 -> 8x8, raw: thread = output pixel p (64) x channel group cg (2 x 16)
        {
            const int p = tid & 63, cg = tid >> 6, oy = p >> 3, ox = p & 7;
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll 1
            for (int ci = 0; ci < 16; ++ci)
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float v = a1[ci * 324 + (2 * oy + (t >> 2)) * 18 + 2 * ox + (t & 3)];
                    const float4* w = reinterpret_cast<const float4*>(w1s + (ci * 16 + t) * 32 + cg * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) fma4x(acc, 4 * q, v, w[q]);
                }
            float* y = a.y1 + (size_t)im * 2048 + cg * 16 * 64 + p;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                y[j * 64] = acc[j];
                const float s = warp_sum(acc[j]), q = warp_sum(acc[j] * acc[j]);
                if ((tid & 31) == 0) {
                    atomicAdd(&ssum[g][cg * 16 + j], (double)s);
                    atomicAdd(&ssq[g][cg * 16 + j], (double)q);
                }
            }
        }
    }
    __syncthreads();
    if (tid < 96 && ((&ssum[0][0])[tid] != 0.0 || (&ssq[0][0])[tid] != 0.0)) {
        atomicAdd(a.st1 + tid * 2, (&ssum[0][0])[tid]);
        atomicAdd(a.st1 + tid * 2 + 1, (&ssq[0][0])[tid]);
    }
}

constexpr int ENCB_IMGS = 4;                                  // images per CTA pass (register-blocked together)
constexpr int ENCB_A2 = 32 * 10 * 10, ENCB_W = 128 * 64;     // one image's padded input; one 8-input-channel weight chunk
constexpr int ENCB_SMEM = (ENCB_IMGS * ENCB_A2 + ENCB_W) * 4;                     // 83 968 B
__global__ void __launch_bounds__(ENC_T) enc_train_b_kernel(const EncTrainArgs a) {
    extern __shared__ float enc_sm[];
    float* a2 = enc_sm;                        // [img 4][32][10][10] zero border
    float* w2s = a2 + ENCB_IMGS * ENCB_A2;     // [k 128][co 64] of the current input-channel chunk
    __shared__ float sc[3][32], sh[3][32];
    __shared__ double ssum[3][64], ssq[3][64];
    const int tid = threadIdx.x;
    if (tid < 96) {
        const int g = tid >> 5, ch = tid & 31;
        bn_coef(a.st1 + tid * 2, (double)a.n[g] * 64.0, a.g1[ch], a.b1[ch], sc[g][ch], sh[g][ch]);
    }
    for (int i = tid; i < ENCB_IMGS * ENCB_A2; i += ENC_T) a2[i] = 0.f;
    for (int i = tid; i < 192; i += ENC_T) (&ssum[0][0])[i] = (&ssq[0][0])[i] = 0.0;
    const int n_total = a.n[0] + a.n[1] + a.n[2];
    const int n_quads = (n_total + ENCB_IMGS - 1) / ENCB_IMGS;
    const int p = tid & 15, cg = tid >> 4, oy = p >> 2, ox = p & 3;      // output pixel (16) x channel group (8 x 8)
    for (int quad = blockIdx.x; quad < n_quads; quad += gridDim.x) {
        const int im0 = quad * ENCB_IMGS, nimg = min(ENCB_IMGS, n_total - im0);
        __syncthreads();                       // coefficients ready / previous quad consumed
        for (int k = tid; k < nimg * 2048; k += ENC_T) {
            const int i = k >> 11, r = k & 2047, ch = r >> 6;
            int g, li;
            enc_group(a, im0 + i, g, li);
            a2[i * ENCB_A2 + ch * 100 + (((r >> 3) & 7) + 1) * 10 + (r & 7) + 1] =
                lrelu_(a.y1[(size_t)(im0 + i) * 2048 + r] * sc[g][ch] + sh[g][ch]);
        }
        float acc[ENCB_IMGS][8];
#pragma unroll
        for (int i = 0; i < ENCB_IMGS; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 1
        for (int c8 = 0; c8 < 4; ++c8) {
            __syncthreads();                   // a2 staged / previous weight chunk consumed
            for (int i = tid; i < ENCB_W; i += ENC_T) w2s[i] = a.w2t[c8 * ENCB_W + i];
            __syncthreads();
#pragma unroll 1
            for (int cl = 0; cl < 8; ++cl)
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int off = (c8 * 8 + cl) * 100 + (2 * oy + (t >> 2)) * 10 + 2 * ox + (t & 3);
                    const float4* w = reinterpret_cast<const float4*>(w2s + (cl * 16 + t) * 64 + cg * 8);
                    const float4 wa = w[0], wb = w[1];
#pragma unroll
                    for (int i = 0; i < ENCB_IMGS; ++i) {
                        const float v = a2[i * ENCB_A2 + off];
                        acc[i][0] = fmaf(v, wa.x, acc[i][0]); acc[i][1] = fmaf(v, wa.y, acc[i][1]);
                        acc[i][2] = fmaf(v, wa.z, acc[i][2]); acc[i][3] = fmaf(v, wa.w, acc[i][3]);
                        acc[i][4] = fmaf(v, wb.x, acc[i][4]); acc[i][5] = fmaf(v, wb.y, acc[i][5]);
                        acc[i][6] = fmaf(v, wb.z, acc[i][6]); acc[i][7] = fmaf(v, wb.w, acc[i][7]);
                    }
                }
        }
#pragma unroll
        for (int i = 0; i < ENCB_IMGS; ++i) {
            if (i < nimg) {                    // uniform per CTA
                int g, li;
                enc_group(a, im0 + i, g, li);
                float* y = a.y2 + (size_t)(im0 + i) * 1024 + cg * 8 * 16 + p;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    y[j * 16] = acc[i][j];
                    float s = acc[i][j], q = acc[i][j] * acc[i][j];
#pragma unroll
                    for (int d = 8; d > 0; d >>= 1) {      // the 16 pixels of a channel sit in one half-warp
                        s += __shfl_xor_sync(0xffffffffu, s, d);
                        q += __shfl_xor_sync(0xffffffffu, q, d);
                    }
                    if (p == 0) {
                        atomicAdd(&ssum[g][cg * 8 + j], (double)s);
                        atomicAdd(&ssq[g][cg * 8 + j], (double)q);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 192; i += ENC_T)
        if ((&ssum[0][0])[i] != 0.0 || (&ssq[0][0])[i] != 0.0) {
            atomicAdd(a.st2 + i * 2, (&ssum[0][0])[i]);
            atomicAdd(a.st2 + i * 2 + 1, (&ssq[0][0])[i]);
        }
}

constexpr int ENCC_IMGS = 16;
constexpr int ENCC_SMEM = ENCC_IMGS * 1024 * 4;                                   // 65 536 B
__global__ void __launch_bounds__(ENC_T) enc_train_c_kernel(const EncTrainArgs a) {
    extern __shared__ float enc_sm[];          // a3 [img 16][k 1024], k = ch*16 + y*4 + x
    __shared__ float sc[3][64], sh[3][64];
    const int tid = threadIdx.x;
    for (int i = tid; i < 192; i += ENC_T) {
        const int g = i >> 6, ch = i & 63;
        bn_coef(a.st2 + i * 2, (double)a.n[g] * 16.0, a.g2[ch], a.b2[ch], sc[g][ch], sh[g][ch]);
    }
    const int n_total = a.n[0] + a.n[1] + a.n[2];
    const int n_blk = (n_total + ENCC_IMGS - 1) / ENCC_IMGS;
    for (int blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
        const int im0 = blk * ENCC_IMGS, nimg = min(ENCC_IMGS, n_total - im0);
        __syncthreads();
        for (int k = tid; k < ENCC_IMGS * 1024; k += ENC_T) {
            const int i = k >> 10, r = k & 1023;
            float v = 0.f;
            if (i < nimg) {
                int g, li;
                enc_group(a, im0 + i, g, li);
                v = lrelu_(a.y2[(size_t)(im0 + i) * 1024 + r] * sc[g][r >> 4] + sh[g][r >> 4]);
                if (g == 1) {
                    a.skip2[(size_t)li * 1024 + r] = v;
                    a.skip2_bf16[(size_t)li * 1024 + r] = __float2bfloat16_rn(v);
                }
            }
            enc_sm[k] = v;
        }
        __syncthreads();
        // head: 1024 -> 128 for 16 images; thread = output channel, 16 accumulators
        float acc[ENCC_IMGS];
#pragma unroll
        for (int i = 0; i < ENCC_IMGS; ++i) acc[i] = 0.f;
        const float* wp = a.w3t + tid;
#pragma unroll 1
        for (int k = 0; k < 1024; k += 4) {
            const float w0 = __ldg(wp + (size_t)k * 128), w1 = __ldg(wp + (size_t)(k + 1) * 128);
            const float w2 = __ldg(wp + (size_t)(k + 2) * 128), w3 = __ldg(wp + (size_t)(k + 3) * 128);
#pragma unroll
            for (int i = 0; i < ENCC_IMGS; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(enc_sm + i * 1024 + k);
                acc[i] = fmaf(v.x, w0, acc[i]);
                acc[i] = fmaf(v.y, w1, acc[i]);
                acc[i] = fmaf(v.z, w2, acc[i]);
                acc[i] = fmaf(v.w, w3, acc[i]);
            }
        }
        const float b = __ldg(a.W.b3 + tid);
#pragma unroll
        for (int i = 0; i < ENCC_IMGS; ++i) {
            if (i < nimg) {
                int g, li;
                enc_group(a, im0 + i, g, li);
                const float s = acc[i] + b;
                if (g == 0) {
                    a.enc_seq[(size_t)li * 128 + tid] = s;
                } else {
                    const size_t r = (size_t)(g == 1 ? a.row0_a : a.row0_b) + li;
                    a.lat_f32[r * 128 + tid] = s;
                    a.lat_bf16[r * 128 + tid] = __float2bfloat16_rn(s);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Conv1d(k=3, pad=1) over time, 128 output channels (ConvSeqEncodingModule, blox/torch/subnetworks.py:120-147).
// grid (T/8, B), 128 threads = output channels, 8 time steps per block.  wT = weights transposed to [k][cin][128].
//   add_time: the input gets the frame index as channel 128 (subnetworks.py:126-128)
//   gn_st   : non-null -> GroupNorm(8) + LReLU is applied to the INPUT while it is staged (statistics per sequence
//             and group over 16 channels x T steps, (sum, sumsq) in gn_st [B][8][2])
//   st_out  : non-null -> accumulate the same statistics of the raw OUTPUT
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv1d_k3_kernel(const float* __restrict__ in, int cin, int add_time,
                                                        const float* __restrict__ wT, const float* __restrict__ bias,
                                                        const double* __restrict__ gn_st, const float* __restrict__ gn_g,
                                                        const float* __restrict__ gn_b, int act, float* __restrict__ out,
                                                        double* st_out, int T) {
    __shared__ float tile[10][132];
    __shared__ float gmean[8], grstd[8];
    const int b = blockIdx.y, t0 = blockIdx.x * 8, tid = threadIdx.x;
    const int cw = cin + (add_time ? 1 : 0);
    if (gn_st != nullptr && tid < 8) {
        const double cnt = 16.0 * T;
        const double m = gn_st[(b * 8 + tid) * 2] / cnt;
        double var = gn_st[(b * 8 + tid) * 2 + 1] / cnt - m * m;
        if (var < 0.0) var = 0.0;
        gmean[tid] = (float)m;
        grstd[tid] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
    for (int r = 0; r < 10; ++r) {
        const int t = t0 - 1 + r;
        for (int c = tid; c < cw; c += 128) {
            float v = 0.f;
            if (t >= 0 && t < T) {
                if (c < cin) {
                    v = in[((size_t)b * T + t) * cin + c];
                    if (gn_st != nullptr) v = lrelu_((v - gmean[c >> 4]) * grstd[c >> 4] * __ldg(gn_g + c) + __ldg(gn_b + c));
                } else {
                    v = (float)t;
                }
            }
            tile[r][c] = v;
        }
    }
    __syncthreads();
    float acc[8];
    const float b0 = bias != nullptr ? __ldg(bias + tid) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = b0;
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < cw; ++c) {
            const float w = __ldg(wT + ((size_t)k * cw + c) * 128 + tid);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(tile[i + k][c], w, acc[i]);
        }
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int t = t0 + i;
        if (t < T) {
            const float v = act == ACT_LRELU ? lrelu_(acc[i]) : acc[i];
            out[((size_t)b * T + t) * 128 + tid] = v;
            s += v;
            q += v * v;
        }
    }
    if (st_out != nullptr) {
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            q += __shfl_xor_sync(0xffffffffu, q, d);
        }
        if ((tid & 15) == 0) {
            atomicAdd(st_out + (b * 8 + (tid >> 4)) * 2, (double)s);
            atomicAdd(st_out + (b * 8 + (tid >> 4)) * 2 + 1, (double)q);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Frame <-> node matching of the balanced binding (tree/frame_binding.py:42-65, tree/tree.py:26-37): the interval
// recursion of prune_map_kernel, additionally returning every node's matched frame (match_timesteps, used to index
// the inference encoding, tree/inference.py:27-34) and whether the node is bound to a frame (c_n_prime non-zero).
// ---------------------------------------------------------------------------------------------
__global__ void match_tables_kernel(const long long* __restrict__ end_ind, int n_cand, int depth, int lcap,
                                    int* __restrict__ frame_node, int* __restrict__ tstep, unsigned char* __restrict__ keep) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_nodes = (1 << depth) - 1;
    if (idx >= n_cand * n_nodes) return;
    const int c = idx / n_nodes, node = idx - c * n_nodes;
    const int slot = node + 1;
    const int tz = __ffs(slot) - 1;
    const int level = depth - 1 - tz;
    const int j = slot >> (tz + 1);
    int l = -1, r = (int)end_ind[c] + 1, t = 0;
    for (int lv = 0; lv <= level; ++lv) {
        t = (l + r) / 2;
        if (lv < level) {
            if ((j >> (level - 1 - lv)) & 1) l = t; else r = t;
        }
    }
    const bool k = t != l && t != r && t >= 0 && t < lcap;
    tstep[idx] = t;
    keep[idx] = k ? 1 : 0;
    if (k) frame_node[c * lcap + t] = node;
}

// e_tilde rows of one tree level: dst[j * Bp + cand][0:128] = inf_seq[cand][tstep[cand][node(level, j)]] (bf16)
__global__ void gather_etilde_kernel(const float* __restrict__ inf_seq, const int* __restrict__ tstep, LevelGeom g, int B,
                                     int T, bf16* __restrict__ dst, int ld) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t rows = (size_t)g.Bp << g.level;
    if (idx >= rows * 128) return;
    const int k = idx & 127;
    const size_t row = idx >> 7;
    const int j = (int)(row / g.Bp), cand = (int)(row - (size_t)j * g.Bp);
    float v = 0.f;
    if (cand < B) {
        const int node = slot_of(g, j, ROW_SELF) - 1;
        int t = tstep[cand * ((1 << g.depth) - 1) + node];
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        v = inf_seq[((size_t)cand * T + t) * 128 + k];
    }
    dst[row * ld + k] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// Batch-statistic BatchNorm of the decoder's three dense layers.  The GEMMs write the raw (pre-BN) layer output as
// bf16 rows [slot][cand][cols]; column -> channel depends on the packed layout of each layer (api.cu pack_decoder):
//   kind 1: n = ch * 16 + pixel                                  (64 channels, 4x4)
//   kind 2: n = iy * 256 + ch * 8 + ix                           (32 channels, 8x8)
//   kind 3: n = ((ch >> 3) * 256 + pixel) * 8 + (ch & 7)         (16 channels, 16x16)
// Rows of padded candidates (cand >= B) are excluded from the statistics.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int dec_col_channel(int kind, int n) {
    if (kind == 1) return n >> 4;
    if (kind == 2) return (n & 255) >> 3;
    return ((n >> 3) >> 8) * 8 + (n & 7);
}
// grid (ld / 256, ceil(rows / 64)), 256 threads: thread = one column over 64 rows
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ x, int rows, int ld, int Bp, int B, int kind,
                                                       double* st /*[C][2]*/) {
    __shared__ double ssum[64], ssq[64];
    const int tid = threadIdx.x, n = blockIdx.x * 256 + tid;
    if (tid < 64) ssum[tid] = ssq[tid] = 0.0;
    __syncthreads();
    const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
    float s = 0.f, q = 0.f;
    for (int r = r0; r < r1; ++r) {
        if (r % Bp >= B) continue;
        const float v = __bfloat162float(x[(size_t)r * ld + n]);
        s += v;
        q = fmaf(v, v, q);
    }
    const int ch = dec_col_channel(kind, n);
    atomicAdd(&ssum[ch], (double)s);
    atomicAdd(&ssq[ch], (double)q);
    __syncthreads();
    if (tid < 64 && (ssum[tid] != 0.0 || ssq[tid] != 0.0)) {
        atomicAdd(st + tid * 2, ssum[tid]);
        atomicAdd(st + tid * 2 + 1, ssq[tid]);
    }
}
// in place: x = relu(x * scale[ch] + shift[ch]); 8 bf16 per thread, grid-stride
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(bf16* x, size_t n_vec /* rows * ld / 8 */, int ld, int kind,
                                                            const double* __restrict__ st, double count,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int C) {
    __shared__ float sc[64], sh[64];
    if (threadIdx.x < C) bn_coef(st + threadIdx.x * 2, count, gamma[threadIdx.x], beta[threadIdx.x], sc[threadIdx.x], sh[threadIdx.x]);
    __syncthreads();
    uint4* x4 = reinterpret_cast<uint4*>(x);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 256) {
        const int n0 = (int)((i * 8) % ld);
        uint4 u = x4[i];
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = dec_col_channel(kind, n0 + k);
            f[k] = fmaxf(f[k] * sc[ch] + sh[ch], 0.f);
        }
        u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
        x4[i] = u;
    }
}

// ---------------------------------------------------------------------------------------------
// Decoder tail + reconstruction NLL of one ground-truth frame (BalancedBinding.reconstruction_loss,
// tree/frame_binding.py:88-100; ProbabilisticConvDecoder.nll, blox/torch/encoder_decoder.py:220-232; ImageDLM /
// DiscreteLogisticMixture / DiscreteLogistic, encoder_decoder.py:150-156, dist.py:87-130,178-197).
// One block per (sequence b, frame t): the node bound to the frame (frames past end_ind: the root, weight 0 unless
// the caller's pad_mask says otherwise) is decoded through the two full-resolution convolutions
//   up(x3) ++ up(skip) -> pad -> conv4x4(32->16) + tanh -> pad -> conv4x4(16->30)
// and the 30 output channels (5 mixture means through a sigmoid, 5 log-scales) give
//   nll = -log(mean_k [cdf((x_bin + 1/256 - mu_k) / s_k) - cdf((x_bin - mu_k) / s_k)] + 1e-7)   (edge bins open-ended)
// summed over the 3 x 32 x 32 sub-pixels.  fp32 SIMT with the operand rounding points of the tcgen05 tail kernel
// (bf16 activations and weights): thread = one image column x 4 rows, weights broadcast from shared memory.
// ---------------------------------------------------------------------------------------------
struct TailNllArgs {
    const bf16* x3;          // [255 * Bp][4096] post-BN layer-3 output, row (node, cand)
    const bf16* skip_up;     // [B][2][DT_PSTRIDE][8]
    const bf16 *w4p, *w5p;   // [16][32][16], [32][16][16] (plain layout, bf16)
    const float *b4, *b5;    // [16], [32]
    const float* traj;       // [B][T][3][32][32]
    const float* pad_mask;   // [B][T]
    const long long* end_ind;
    const int* frame_node;   // [B][lcap]
    int Bp, T, lcap, root_node;
    float* nll_bt;           // [B][T]  (already multiplied by pad_mask)
};
constexpr int TN_W4_FLOATS = 16 * 32 * 16, TN_W5_FLOATS = 16 * 16 * 32;
constexpr int TN_SMEM_BYTES = 6 * DT_PLANE_BYTES + (TN_W4_FLOATS + TN_W5_FLOATS) * 4 + 128 + 256;

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

#ifdef GCPB200_VERIFY
__global__ void __launch_bounds__(256) dec_tail_nll_kernel(const __grid_constant__ TailNllArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* in4 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* in5 = in4 + 4 * DT_PLANE_BYTES;
    float* w4s = reinterpret_cast<float*>(in5 + 2 * DT_PLANE_BYTES);   // [tap][ci 32][co 16]
    float* w5s = w4s + TN_W4_FLOATS;                                   // [tap][ci 16][2 halves x 16]
    double* red = reinterpret_cast<double*>(w5s + TN_W5_FLOATS);
    const int tid = threadIdx.x;
    const int b = blockIdx.x / a.T, t = blockIdx.x - b * a.T;
    const float pad = a.pad_mask[blockIdx.x];
    if (pad == 0.f) {                      // uniform per block
        if (tid == 0) a.nll_bt[blockIdx.x] = 0.f;
        return;
    }
    const int node = t <= (int)a.end_ind[b] ? a.frame_node[b * a.lcap + t] : a.root_node;
    for (int i = tid; i < 6 * DT_PLANE_BYTES / 16; i += 256) reinterpret_cast<uint4*>(in4)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < TN_W4_FLOATS; i += 256) {
        const int co = i & 15, ci = (i >> 4) & 31, tap = i >> 9;
        w4s[i] = __bfloat162float(a.w4p[(co * 32 + ci) * 16 + tap]);
    }
    for (int i = tid; i < TN_W5_FLOATS; i += 256) {
        const int col = i & 31, ci = (i >> 5) & 15, tap = i >> 9;   // column = half * 16 + j  <-  head channel half * 15 + j
        const int j = col & 15, co = (col >> 4) * 15 + j;
        w5s[i] = j < 15 ? __bfloat162float(a.w5p[(co * 16 + ci) * 16 + tap]) : 0.f;
    }
    __syncthreads();
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.skip_up + (size_t)b * 2 * DT_PSTRIDE * 8);
        uint4* dst = reinterpret_cast<uint4*>(in4 + 2 * DT_PLANE_BYTES);
        for (int i = tid; i < 2 * DT_PSTRIDE; i += 256) dst[i] = __ldg(src + i);
    }
    dt_build_up(in4, a.x3 + ((size_t)node * a.Bp + b) * 4096, tid, 256);
    __syncthreads();

    const int x = tid & 31, y0 = (tid >> 5) * 4;      // output pixels (y0 .. y0+3, x)
    // ---- conv 4: 32 -> 16 channels, tanh, into the padded planes of conv 5
    {
        float acc[4][16];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[p][co] = __ldg(a.b4 + co);
#pragma unroll 1
        for (int plane = 0; plane < 4; ++plane) {
            const uint8_t* pb = in4 + plane * DT_PLANE_BYTES;
#pragma unroll 1
            for (int tap = 0; tap < 16; ++tap) {               // weights of one tap are shared by the thread's 4 pixels
                const int ky = tap >> 2, kx = tap & 3;
                float v[4][8];
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    unpack8(*reinterpret_cast<const uint4*>(pb + ((y0 + p + ky) * DT_WP + x + kx) * 16), v[p]);
                const float* w = w4s + (tap * 32 + plane * 8) * 16;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const float4* w4 = reinterpret_cast<const float4*>(w + c8 * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ww = w4[q];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            acc[p][4 * q + 0] = fmaf(v[p][c8], ww.x, acc[p][4 * q + 0]);
                            acc[p][4 * q + 1] = fmaf(v[p][c8], ww.y, acc[p][4 * q + 1]);
                            acc[p][4 * q + 2] = fmaf(v[p][c8], ww.z, acc[p][4 * q + 2]);
                            acc[p][4 * q + 3] = fmaf(v[p][c8], ww.w, acc[p][4 * q + 3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int pix = (y0 + p + 1) * DT_WP + x + 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint4 u;
                u.x = pack_bf16x2(tanhf(acc[p][8 * h + 0]), tanhf(acc[p][8 * h + 1]));
                u.y = pack_bf16x2(tanhf(acc[p][8 * h + 2]), tanhf(acc[p][8 * h + 3]));
                u.z = pack_bf16x2(tanhf(acc[p][8 * h + 4]), tanhf(acc[p][8 * h + 5]));
                u.w = pack_bf16x2(tanhf(acc[p][8 * h + 6]), tanhf(acc[p][8 * h + 7]));
                *reinterpret_cast<uint4*>(in5 + h * DT_PLANE_BYTES + pix * 16) = u;
            }
        }
    }
    __syncthreads();
    // ---- conv 5: 16 -> 30 channels in two halves of 15 (mixture means, then log-scales), then the NLL
    float mu[4][15];
    double nll_sum = 0.0;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float acc[4][16];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[p][co] = co < 15 ? __ldg(a.b5 + half * 15 + co) : 0.f;
#pragma unroll 1
        for (int plane = 0; plane < 2; ++plane) {
            const uint8_t* pb = in5 + plane * DT_PLANE_BYTES;
#pragma unroll 1
            for (int tap = 0; tap < 16; ++tap) {
                const int ky = tap >> 2, kx = tap & 3;
                float v[4][8];
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    unpack8(*reinterpret_cast<const uint4*>(pb + ((y0 + p + ky) * DT_WP + x + kx) * 16), v[p]);
                const float* w = w5s + (tap * 16 + plane * 8) * 32 + half * 16;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const float4* w4 = reinterpret_cast<const float4*>(w + c8 * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ww = w4[q];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            acc[p][4 * q + 0] = fmaf(v[p][c8], ww.x, acc[p][4 * q + 0]);
                            acc[p][4 * q + 1] = fmaf(v[p][c8], ww.y, acc[p][4 * q + 1]);
                            acc[p][4 * q + 2] = fmaf(v[p][c8], ww.z, acc[p][4 * q + 2]);
                            acc[p][4 * q + 3] = fmaf(v[p][c8], ww.w, acc[p][4 * q + 3]);
                        }
                    }
                }
            }
        }
        if (half == 0) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int co = 0; co < 15; ++co) mu[p][co] = sigmoid_acc(acc[p][co]);
        } else {
            const float* tgt = a.traj + (size_t)blockIdx.x * 3072;
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float x01 = (tgt[ch * 1024 + (y0 + p) * 32 + x] + 1.0f) * 0.5f;
                    const float xb = floorf(x01 * 256.0f) * (1.0f / 256.0f);
                    float pm = 0.f;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const float inv = expf(-acc[p][k * 3 + ch]);                 // 1 / scale
                        const float xs = (xb - mu[p][k * 3 + ch]) * inv;
                        const float hi = sigmoid_acc(xs + (1.0f / 256.0f) * inv), lo = sigmoid_acc(xs);
                        float pr = hi - lo;
                        if (x01 == 0.f) pr = hi;
                        if (x01 == 1.f) pr = 1.0f - lo;
                        pm += pr;
                    }
                    nll_sum -= (double)logf(pm * 0.2f + 1e-7f);
                }
        }
    }
    const double s = block_sum_d(nll_sum, red);
    if (tid == 0) a.nll_bt[blockIdx.x] = (float)(s * (double)pad);
}
#endif

// Reconstruction NLL from the raw head outputs of the tcgen05 tail kernel (dec_tail3_raw_kernel run twice: mixture
// means, log-scales): the same arithmetic as the second half of dec_tail_nll_kernel.  One block per (sequence, frame);
// frames whose node is outside [node0, node0 + n_chunk) are left to the launch of their node chunk.
struct RawNllArgs {
    const float* raw_mu;     // [B][n_chunk][1024][16]: channel k*3 + colour = mean logit of mixture k
    const float* raw_ls;     // same layout: log-scales
    const float* traj;       // [B][T][3][32][32]
    const float* pad_mask;   // [B][T]
    const long long* end_ind;
    const int* frame_node;   // [B][lcap]
    int T, lcap, root_node, node0, n_chunk;
    float* nll_bt;           // [B][T]  (multiplied by pad_mask)
};
__global__ void __launch_bounds__(256) dlm_nll_raw_kernel(const RawNllArgs a) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int b = blockIdx.x / a.T, t = blockIdx.x - b * a.T;
    const float pad = a.pad_mask[blockIdx.x];
    if (pad == 0.f) {                      // uniform per block
        if (tid == 0 && a.node0 == 0) a.nll_bt[blockIdx.x] = 0.f;
        return;
    }
    const int node = t <= (int)a.end_ind[b] ? a.frame_node[b * a.lcap + t] : a.root_node;
    if (node < a.node0 || node >= a.node0 + a.n_chunk) return;
    const size_t base = (((size_t)b * a.n_chunk + (node - a.node0)) * 1024) * 16;
    const float* tgt = a.traj + (size_t)blockIdx.x * 3072;
    double nll_sum = 0.0;
    for (int p = tid; p < 1024; p += 256) {
        float m[16], ls[16];
        const float4* m4 = reinterpret_cast<const float4*>(a.raw_mu + base + (size_t)p * 16);
        const float4* s4 = reinterpret_cast<const float4*>(a.raw_ls + base + (size_t)p * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 x = __ldg(m4 + v), y = __ldg(s4 + v);
            m[4 * v] = x.x; m[4 * v + 1] = x.y; m[4 * v + 2] = x.z; m[4 * v + 3] = x.w;
            ls[4 * v] = y.x; ls[4 * v + 1] = y.y; ls[4 * v + 2] = y.z; ls[4 * v + 3] = y.w;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float x01 = (tgt[ch * 1024 + p] + 1.0f) * 0.5f;
            const float xb = floorf(x01 * 256.0f) * (1.0f / 256.0f);
            float pm = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float inv = expf(-ls[k * 3 + ch]);                 // 1 / scale
                const float xs = (xb - sigmoid_acc(m[k * 3 + ch])) * inv;
                const float hi = sigmoid_acc(xs + (1.0f / 256.0f) * inv), lo = sigmoid_acc(xs);
                float pr = hi - lo;
                if (x01 == 0.f) pr = hi;
                if (x01 == 1.f) pr = 1.0f - lo;
                pm += pr;
            }
            nll_sum -= (double)logf(pm * 0.2f + 1e-7f);
        }
    }
    const double s = block_sum_d(nll_sum, red);
    if (tid == 0) a.nll_bt[blockIdx.x] = (float)(s * (double)pad);
}

// ---------------------------------------------------------------------------------------------
// KL(q || p) of diagonal Gaussians summed over the 255 nodes x 256 dims of one sequence
// (Gaussian.kl_divergence, blox/torch/dist.py:249-252; KLDivLoss2, blox/torch/losses.py:75-109).  grid B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_seq_kernel(const float* __restrict__ q_mu, const float* __restrict__ q_ls,
                                                     const float* __restrict__ p_mu, const float* __restrict__ p_ls, int per_seq,
                                                     float* __restrict__ kl_b) {
    __shared__ double red[32];
    const size_t base = (size_t)blockIdx.x * per_seq;
    double s = 0.0;
    for (int i = threadIdx.x; i < per_seq; i += 256) {
        const float qm = q_mu[base + i], ql = q_ls[base + i], pm = p_mu[base + i], pl = p_ls[base + i];
        const float qs = expf(ql), ps = expf(pl), d = qm - pm;
        s += (double)((pl - ql) + (qs * qs + d * d) / (2.0f * ps * ps) - 0.5f);
    }
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) kl_b[blockIdx.x] = (float)s;
}

// Ground-truth cost of the cost model's frame pair when the model is configured with EuclideanPathLength
// (experiments/prediction/25room/gcp_tree/conf.py:33-35; gcp/planning/cem/cost_fcn.py:9-22,49-54; cost_mdl.py:112-113):
// sum over t in [start, end) and over (channel, image row) of || traj[t+1][c][y][:] - traj[t][c][y][:] ||_2
// (the last step compares frame `end` with itself).  grid B.
__global__ void __launch_bounds__(256) path_length_kernel(const float* __restrict__ traj, const long long* __restrict__ cs,
                                                          const long long* __restrict__ ce, int T, float* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    const int t0 = (int)cs[b], t1 = (int)ce[b];
    double s = 0.0;
    for (int r = threadIdx.x; r < (t1 - t0) * 96; r += 256) {
        const int t = t0 + r / 96, row = r % 96;
        const float4* p = reinterpret_cast<const float4*>(traj + ((size_t)b * T + t) * 3072 + row * 32);
        const float4* q = p + 768;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 u = __ldg(p + k), v = __ldg(q + k);
            const float dx = v.x - u.x, dy = v.y - u.y, dz = v.z - u.z, dw = v.w - u.w;
            acc += dx * dx + dy * dy + dz * dz + dw * dw;
        }
        s += (double)sqrtf(acc);
    }
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) out[b] = (float)s;
}

// pair rows of the training-time auxiliary heads (inverse_mdl.py:139-170; cost_mdl.py:55-67,100-113):
//   row b       = [enc_traj_seq[b][t0[b]] | model_enc_seq[b][t1[b]]]      (inverse model)
//   row 128 + b = [model_enc_seq[b][cs[b]] | model_enc_seq[b][ce[b]]]     (cost model)
__global__ void train_pairs_kernel(const float* __restrict__ enc_seq, const float* __restrict__ seq, const long long* t0,
                                   const long long* t1, const long long* cs, const long long* ce, int B, int T,
                                   bf16* __restrict__ pairs) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 256 * 256) return;
    const int k = idx & 255, row = idx >> 8, b = row & 127;
    float v = 0.f;
    if (b < B) {
        if (row < 128) v = k < 128 ? enc_seq[((size_t)b * T + t0[b]) * 128 + k] : seq[((size_t)b * T + t1[b]) * 128 + (k - 128)];
        else v = k < 128 ? seq[((size_t)b * T + cs[b]) * 128 + k] : seq[((size_t)b * T + ce[b]) * 128 + (k - 128)];
    }
    pairs[idx] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// Scalar losses (TreeModel.loss + get_total_loss, base_gcp.py:264-301; tree/tree_module.py:116-131), one block.
// Output order = GCPB200_LOSS_* in include/gcpb200.h.
// ---------------------------------------------------------------------------------------------
struct LossArgs {
    int B, T, n_nodes;
    const float* logits; int logits_ld;     // length predictor [B][200]
    const long long* end_ind;
    const float* nll_bt;                    // [B][T]
    const float* kl_b;                      // [B]
    const float* existence;                 // [B][255] logits, depth-first
    const unsigned char* keep;              // [B][255]
    const float* reg;                       // [B][T][2] state regressor on the zero-padded matched latents
    const float* states;                    // [B][T][2]
    const float* pad_mask;                  // [B][T]
    const float* inv_pred;                  // [128][2]
    const float* actions;                   // [B][T-1][2]
    const long long* inv_t0;
    const float* cost_pred;                 // [128]
    const float* cost_target;               // [B]
    double frame_elems;                     // T * 3 * 32 * 32
    float* losses;                          // [9]
    float* nll_seq;                         // optional [B]
};
__global__ void __launch_bounds__(256) loss_finalize_kernel(const LossArgs a) {
    __shared__ double red[32];
    __shared__ int lmax_s;
    const int tid = threadIdx.x;
    double out[9];
    // len_pred: cross entropy of the length logits against end_ind (misc.py:53-57)
    double v = 0.0;
    for (int b = tid; b < a.B; b += 256) {
        const float* lg = a.logits + (size_t)b * a.logits_ld;
        float m = -INFINITY;
        for (int k = 0; k < a.T; ++k) m = fmaxf(m, lg[k]);
        double se = 0.0;
        for (int k = 0; k < a.T; ++k) se += exp((double)(lg[k] - m));
        v += (double)m + log(se) - (double)lg[a.end_ind[b]];
    }
    out[0] = block_sum_d(v, red) / a.B;
    // action_reconst (inverse_mdl.py:180-190)
    v = 0.0;
    for (int i = tid; i < a.B * 2; i += 256) {
        const int b = i >> 1, k = i & 1;
        const double d = (double)a.inv_pred[b * 2 + k] - (double)a.actions[((size_t)b * (a.T - 1) + a.inv_t0[b]) * 2 + k];
        v += d * d;
    }
    out[1] = block_sum_d(v, red) / (a.B * 2);
    // cost_estimation (cost_mdl.py:69-72)
    v = 0.0;
    for (int b = tid; b < a.B; b += 256) {
        const double d = (double)a.cost_pred[b] - (double)a.cost_target[b];
        v += d * d;
    }
    out[2] = block_sum_d(v, red) / a.B;
    // state_regression over the first Lmax = max(end_ind) + 1 steps (base_gcp.py:284-288)
    if (tid == 0) {
        int lm = 0;
        for (int b = 0; b < a.B; ++b) lm = max(lm, (int)a.end_ind[b] + 1);
        lmax_s = min(lm, a.T);
    }
    __syncthreads();
    const int lmax = lmax_s;
    v = 0.0;
    for (int i = tid; i < a.B * lmax * 2; i += 256) {
        const int k = i & 1, t = (i >> 1) % lmax, b = (i >> 1) / lmax;
        const size_t o = ((size_t)b * a.T + t) * 2 + k;
        const double d = (double)a.reg[o] - (double)a.states[o];
        v += d * d * (double)a.pad_mask[b * a.T + t];
    }
    out[3] = block_sum_d(v, red) / ((double)a.B * lmax * 2);
    // dense_img_rec: sum over frames and sub-pixels, mean over the batch (encoder_decoder.py:220-232)
    v = 0.0;
    for (int i = tid; i < a.B * a.T; i += 256) v += (double)a.nll_bt[i];
    out[4] = block_sum_d(v, red) / a.B;
    if (a.nll_seq != nullptr)
        for (int b = tid; b < a.B; b += 256) {
            double s = 0.0;
            for (int t = 0; t < a.T; ++t) s += (double)a.nll_bt[b * a.T + t];
            a.nll_seq[b] = (float)s;
        }
    // kl
    v = 0.0;
    for (int b = tid; b < a.B; b += 256) v += (double)a.kl_b[b];
    out[5] = block_sum_d(v, red) / a.B;
    // existence_predictor: BCE with logits against "node is bound to a frame" (frame_binding.py:80-86)
    v = 0.0;
    for (int i = tid; i < a.B * a.n_nodes; i += 256) {
        const double x = (double)a.existence[i], y = a.keep[i] ? 1.0 : 0.0;
        v += fmax(x, 0.0) - x * y + log1p(exp(-fabs(x)));
    }
    out[6] = block_sum_d(v, red) / ((double)a.B * a.n_nodes);
    if (tid == 0) {
        out[7] = 0.0;   // entropy of the one-hot matching (tree_module.py:127,141), weight 0
        double tot = 0.0;
        for (int i = 0; i < 7; ++i) tot += out[i];
        out[8] = tot / a.frame_elems;
        for (int i = 0; i < 9; ++i) a.losses[i] = (float)out[i];
    }
}

}  // namespace gcp
