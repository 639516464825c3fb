// Decoder tail v3: the two full-resolution convolutions of the GCP decoder as quad-row implicit GEMMs.
//
//   x3 [16ch,16x16] --bilinear x2--> (cat with up(skip), split off as a per-candidate constant S4)
//      --ZeroPad(1,2,1,2)--> conv k4 (->16) + S4, tanh = feat [16,32,32]
//      --ZeroPad(1,2,1,2)--> conv k4 (16->30): only the 15 mixture-mean channels are needed for the image
//      --> image = mean_5 sigmoid(.)*2-1 = 0.2 * sum_5 tanh(./2)
//   (blox/torch/encoder_decoder.py:56-97,150-218; blox/torch/layers.py:128-150; blox/torch/dist.py:200-201)
//
// What the measurements say (profiles/r1_microbench_umma_ldtm_shfl.txt): a tcgen05.mma with K=16 costs
// ~46-48 cycles for ANY M <= 128 and N <= 64, one SM moves one warp-shuffle per cycle, and v2's shuffle
// epilogues (13 k cycles per image) bound that kernel, not its MMAs (4 k cycles).  v3 therefore removes the
// epilogue work instead of MMA work:
//   * GEMM row = QUAD of 4 horizontally adjacent output pixels; N = (pixel in quad, channel) = 64.  The 4
//     horizontal taps become a Toeplitz weight matrix: output quad c' needs input pixels 4c'..4c'+6 = input
//     quad c' (phases 0..3) and quad c'+1 (phases 0..2), i.e. 7 MMAs (K = 16 channels) per filter row.  All 7
//     weight matrices are 64-row windows of one 160-row array Z = [0 0 0 W0 W1 W2 W3 0 0 0] when N is ordered
//     with the pixel index reversed, so the B operand is again "the same smem array with a shifted start".
//   * Input planes are stored per pixel phase: P[phase j][channel half h][quad-row = Y*9 + C][8ch] (16 B
//     cells, un-swizzled K-major core matrices).  Padded rows are 36 px = 9 quads, but only 8 output quads
//     per row exist: the A descriptor's 8-row-group stride (SBO) is set to 9 cells = 144 B, so one group of 8
//     GEMM rows is exactly one image row and a 128-row tile is exactly 16 image rows -- no dead rows, two
//     tiles per image, filter taps are start-address shifts (ky*9 + t cells).
//   * No shuffles, no cross-thread exchange: thread = quad, accumulator drain = output size.  Epilogue 1 adds
//     the skip term, applies tanh and writes bf16 feature quads straight into the second conv's phase planes;
//     epilogue 2 sums the mixture and writes three float4 per thread, 512 B contiguous per warp and colour.
// Tensor pipe: 2 tiles x 2 convs x 28 MMAs x 48 cycles = 5.4 k cycles per image (v2: 18.9 k measured).
#pragma once
#include "dec_tail.cuh"

namespace gcp {

constexpr int D3_QP = 9;                        // quads per padded row (36 px)
constexpr int D3_ROWS = 35 * D3_QP;             // quad-rows per plane (padded rows Y = 0..34)
constexpr int D3_PLANE_BYTES = D3_ROWS * 16;    // 5040
constexpr int D3_BUF_BYTES = 8 * D3_PLANE_BYTES;   // planes [phase 4][half 2]
constexpr int D3_Z_ROWS = 160;                  // 10 blocks of 16 rows
constexpr int D3_Z_CHUNK = D3_Z_ROWS * 16;      // one 8-channel K chunk of Z
constexpr int D3_Z_KY = 2 * D3_Z_CHUNK;         // per filter row
constexpr int D3_W_BYTES = 4 * D3_Z_KY;         // 20480 per convolution
constexpr int D3_S4_BYTES = 256 * 64 * 2;       // bf16 [quad 256][n 64]
constexpr int D3_X3_BYTES = 8192;
constexpr int D3_NB = 4;                        // accumulator buffers per convolution (4 x 64 columns each)
constexpr int D3_OFF_IN4 = 0;
constexpr int D3_OFF_IN5 = D3_OFF_IN4 + 2 * D3_BUF_BYTES;
constexpr int D3_OFF_W4 = D3_OFF_IN5 + 2 * D3_BUF_BYTES;
constexpr int D3_OFF_W5 = D3_OFF_W4 + D3_W_BYTES;
constexpr int D3_OFF_X3 = D3_OFF_W5 + D3_W_BYTES;
constexpr int D3_OFF_BIAS = D3_OFF_X3 + D3_X3_BYTES;
constexpr int D3_OFF_BAR = D3_OFF_BIAS + 64;
constexpr int D3_SMEM_BYTES = D3_OFF_BAR + 256 + 128;
constexpr int D3_THREADS = 18 * 32;             // warps 0-3 epilogue 1 (tile 0), 4-7 epilogue 2, 8-11 up-sampler,
                                                // 12 UMMA issuer conv 4, 13-16 epilogue 1 (tile 1), 17 UMMA issuer conv 5

struct DecTail3Args {
    const bf16* x3;        // [n_slots * Bp][4096]  rows (slot_local, cand); [plane 2][y16][x16][8]
    const bf16* s4;        // [n_cand or 1][256 quads][64] bf16: skip half of conv 32->16 incl. its bias
    int s4_stride;         // elements between candidates (0: shared)
    const bf16* w4;        // D3_W_BYTES: Z arrays of the x-half of the 32->16 conv
    const bf16* w5;        // D3_W_BYTES: Z arrays of the 15 mixture-mean channels, pre-scaled by 1/2
    const float* b5h;      // [16] 0.5 * bias of those channels (entry 15 = 0)
    float* images;         // [B][n_nodes][3][32][32]
    int Bp, n_cand, slot0, n_slots, n_nodes;
    int slot_extra;        // x3 row block sl holds tree slot slot0 + sl * (1 + slot_extra); node = slot - 1, and a block
                           // whose node is -1 (slot 0 = the start frame) is computed but not stored
    unsigned long long* prof;   // optional [16]: per-role wait / work cycles
    // pixel-copy head (dec_tail3_pc_kernel; PixelCopyDecoder, blox/torch/encoder_decoder.py:235-259): w5 holds gen_head
    // (channels 0-2, unscaled) and mask_head (3-5), b5h their biases
    // raw head (dec_tail3_raw_kernel): w5 / b5h hold 15 head channels + one zero slot, unscaled; the pre-activation
    // outputs go to raw [n_cand][n_slots][1024 px][16] fp32 (block sl = slot slot0 + sl) for the training-phase NLL
    float* raw;
    const float* src0;     // [n_cand or 1][3][32][32] start image
    const float* srcg;     // goal image
    int src_stride;        // elements between candidates (0: shared)
    // ---- list mode (planner mode "decode only the nodes balanced pruning keeps"): image i of this launch is x3 row i,
    // its (candidate, node) come from row_cand / row_node at index img_base + i, and the number of images is known to the
    // device only: min(n_slots * Bp, max(0, *n_img_dev - img_base)).  n_cand / slot0 / slot_extra are unused.
    const int* n_img_dev;
    const int* row_cand;
    const int* row_node;
    int img_base;
    // ---- fused L2 image cost (HEAD 0; L2ImageCost._compute, gcp/planning/cem/cost_fcn.py:65-72): sum over the image's 3072
    // sub-pixels of (image - goal)^2, reduced in a fixed order (thread: 24 values; warp butterfly; 4 warps in order) and
    // written to frame_sq[img_base + i] (list mode) or frame_sq[cand * n_nodes + node]; images may then be NULL (no image
    // is written at all).
    float* frame_sq;
    const float* l2_goal;  // [3][32][32] in [-1,1]
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// D[tmem] (+)= A * B with the descriptors given as (lo, hi) words: only the lo word (start address) changes
// between the MMAs of a tile, so the issuing warp spends one 32-bit add per operand.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 28 MMAs of one 128-quad tile: filter row ky x Toeplitz offset s = 4t + j (input quad c'+t, pixel phase j).
// a_lo / b_lo: lo words of the descriptors of the tile's first operand cells (start address in units of 16 B).
__device__ __forceinline__ void d3_conv_tile(uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t d_tmem) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int s = 0; s < 7; ++s) {
            const uint32_t da = a_lo + (uint32_t)(((s & 3) * 2 * D3_PLANE_BYTES + (ky * D3_QP + (s >> 2)) * 16) >> 4);
            const uint32_t db = b_lo + (uint32_t)((ky * D3_Z_KY + s * 256) >> 4);
            if (elect_one()) umma_bf16_lohi(d_tmem, da, a_hi, db, b_hi, idesc, (ky | s) != 0);
        }
}

#define D3_T() (prof_on ? clock64() : 0ll)

// HEAD 0: discrete-logistic-mixture mean (15 channels); HEAD 1: pixel-copy head (3 generated + 3 mask channels);
// HEAD 2: raw pre-activation output of 16 head channels
template <int HEAD>
__device__ __forceinline__ void dec_tail3_body(const DecTail3Args& a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* in4 = smem + D3_OFF_IN4;
    uint8_t* in5 = smem + D3_OFF_IN5;
    uint8_t* w4 = smem + D3_OFF_W4;
    uint8_t* w5 = smem + D3_OFF_W5;
    uint8_t* x3s = smem + D3_OFF_X3;
    float* bias5 = reinterpret_cast<float*>(smem + D3_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D3_OFF_BAR);
    uint64_t* x3_full = bars + 0;       // tx
    uint64_t* in4_full = bars + 1;      // [2] count 128 (up-sampler threads)
    uint64_t* in4_empty = bars + 3;     // [2] tcgen05.commit after the image's last conv-4 MMA
    uint64_t* d4_full = bars + 5;       // [4] commit
    uint64_t* d4_empty = bars + 9;      // [4] count 128
    uint64_t* d5_full = bars + 13;      // [4] commit
    uint64_t* d5_empty = bars + 17;     // [4] count 128
    uint64_t* feat_full = bars + 21;    // [2] count 256: both feature tiles of an image are in in5[buf]
    uint64_t* m5_done = bars + 23;      // [2] commit: both conv-5 tiles of an image have been read from in5[buf]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 25);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform role index

    for (int i = tid; i < 4 * D3_BUF_BYTES / 16; i += D3_THREADS)
        reinterpret_cast<uint4*>(in4)[i] = make_uint4(0, 0, 0, 0);     // padding cells stay zero forever
    for (int i = tid; i < D3_W_BYTES / 16; i += D3_THREADS) {
        reinterpret_cast<uint4*>(w4)[i] = __ldg(reinterpret_cast<const uint4*>(a.w4) + i);
        reinterpret_cast<uint4*>(w5)[i] = __ldg(reinterpret_cast<const uint4*>(a.w5) + i);
    }
    if (tid < 16) bias5[tid] = a.b5h[tid];
    if (tid == 0) {
        mbar_init(x3_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&in4_full[i], 128);
            mbar_init(&in4_empty[i], 1);
            mbar_init(&feat_full[i], 256);
            mbar_init(&m5_done[i], 1);
        }
        for (int i = 0; i < D3_NB; ++i) {
            mbar_init(&d4_full[i], 1);
            mbar_init(&d4_empty[i], 128);
            mbar_init(&d5_full[i], 1);
            mbar_init(&d5_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc(tmem_holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    constexpr uint32_t TM_D5 = D3_NB * 64;

    // this CTA's contiguous run of images; image index = cand * n_slots + slot_local (candidate-major, so
    // a per-candidate skip term is reloaded rarely)
    const bool list = a.row_cand != nullptr;
    const int n_total = list ? min(a.n_slots * a.Bp, max(__ldg(a.n_img_dev) - a.img_base, 0)) : a.n_cand * a.n_slots;
    const int per = (n_total + gridDim.x - 1) / gridDim.x;
    const int img0 = blockIdx.x * per;
    const int img1 = min(img0 + per, n_total);
    const int n_img = max(img1 - img0, 0);
    const bool prof_on = (a.prof != nullptr) && lane == 0 && (warp == 0 || warp == 4 || warp == 8 || warp == 12 || warp == 13 || warp == 17);
    long long pc[4] = {0, 0, 0, 0};

    if (warp >= 8 && warp < 12) {
        // =================== up-sampler: x3 (bulk copy -> staging) -> bilinear x2 -> phase planes of in4[buf] ===========
        const int bt = tid - 256;            // 0..127
        if (bt == 0 && img0 < img1) {
            const int cand = img0 / a.n_slots, sl = img0 - cand * a.n_slots;
            mbar_arrive_expect_tx(x3_full, D3_X3_BYTES);
            bulk_load_1d(x3s, a.x3 + (list ? (size_t)img0 : (size_t)sl * a.Bp + cand) * 4096, D3_X3_BYTES, x3_full);
        }
        uint32_t n = 0;
        for (int img = img0; img < img1; ++img, ++n) {
            const int buf = n & 1;
            long long t0 = D3_T();
            mbar_wait(x3_full, n & 1);
            if (n >= 2) mbar_wait(&in4_empty[buf], ((n >> 1) - 1) & 1);
            long long t1 = D3_T();
            pc[0] += t1 - t0;
            uint8_t* dst = in4 + buf * D3_BUF_BYTES;
            const uint4* src = reinterpret_cast<const uint4*>(x3s);
            // work item = (channel half h, low-res pixel (i,k)): a 2x2 block of outputs from its 3x3 neighbourhood
#pragma unroll 1
            for (int it = bt; it < 512; it += 128) {
                const int h = it >> 8, i = (it >> 4) & 15, k = it & 15;
                const int im = max(i - 1, 0), ip = min(i + 1, 15), km = max(k - 1, 0), kp = min(k + 1, 15);
                const uint4* sp = src + h * 256;
                float hz[3][2][8];      // horizontally interpolated: rows (i-1,i,i+1) x (left,right output column)
                const int rows[3] = {im, i, ip};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float l[8], c[8], rr[8];
                    unpack8(sp[rows[r] * 16 + km], l);
                    unpack8(sp[rows[r] * 16 + k], c);
                    unpack8(sp[rows[r] * 16 + kp], rr);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        hz[r][0][e] = 0.25f * l[e] + 0.75f * c[e];
                        hz[r][1][e] = 0.75f * c[e] + 0.25f * rr[e];
                    }
                }
#pragma unroll
                for (int ay = 0; ay < 2; ++ay)
#pragma unroll
                    for (int bx = 0; bx < 2; ++bx) {
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            o[e] = ay == 0 ? 0.25f * hz[0][bx][e] + 0.75f * hz[1][bx][e]
                                           : 0.75f * hz[1][bx][e] + 0.25f * hz[2][bx][e];
                        uint4 u;
                        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
                        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
                        const int Y = 2 * i + ay + 1, X = 2 * k + bx + 1;
                        *reinterpret_cast<uint4*>(dst + ((X & 3) * 2 + h) * D3_PLANE_BYTES + (Y * D3_QP + (X >> 2)) * 16) = u;
                    }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);               // every up-sampler thread is done with the staging tile
            if (bt == 0 && img + 1 < img1) {
                const int ni = img + 1, cand = ni / a.n_slots, sl = ni - cand * a.n_slots;
                mbar_arrive_expect_tx(x3_full, D3_X3_BYTES);
                bulk_load_1d(x3s, a.x3 + (list ? (size_t)ni : (size_t)sl * a.Bp + cand) * 4096, D3_X3_BYTES, x3_full);
            }
            mbar_arrive(&in4_full[buf]);
            pc[1] += D3_T() - t1;
        }
        if (prof_on) { atomicAdd(a.prof + 0, (unsigned long long)pc[0]); atomicAdd(a.prof + 1, (unsigned long long)pc[1]); }
    } else if (warp == 12) {
        // =================== UMMA issuer, conv 4 ===================
        // The whole warp runs this loop with warp-uniform control flow and values; each tcgen05 instruction is
        // issued by one elected lane.  (Issuing from inside an `if (lane == 0)` region makes the compiler wrap
        // every MMA in a uniform-register broadcast loop: 83 instead of 48 cycles per MMA, measured.)  The two
        // convolutions have separate issuing warps on different SM sub-partitions so that neither the issue
        // rate nor a wait of one stream can starve the tensor pipe.
        const uint64_t dw = umma_desc_nosw(smem_u32(w4), D3_Z_CHUNK, 128);
        const uint64_t din = umma_desc_nosw(smem_u32(in4), D3_PLANE_BYTES, D3_QP * 16);
        const uint32_t a_hi = (uint32_t)(din >> 32), b_hi = (uint32_t)(dw >> 32), b_lo = (uint32_t)dw;
        const long long mma_t0 = D3_T();
        uint32_t n4 = 0;
        for (int n = 0; n < n_img; ++n) {
            const int buf = n & 1;
            long long w0 = D3_T();
            mbar_wait(&in4_full[buf], (n >> 1) & 1);
            pc[0] += D3_T() - w0;
            tc_fence_after();
            const uint32_t a_lo = (uint32_t)din + (uint32_t)((buf * D3_BUF_BYTES) >> 4);
#pragma unroll
            for (int T = 0; T < 2; ++T, ++n4) {
                const int b4 = n4 & (D3_NB - 1);
                long long w1 = D3_T();
                if (n4 >= D3_NB) mbar_wait(&d4_empty[b4], ((n4 / D3_NB) - 1) & 1);
                pc[1] += D3_T() - w1;
                tc_fence_after();
                d3_conv_tile(a_lo + T * 16 * D3_QP, a_hi, b_lo, b_hi, tmem + b4 * 64);
                if (elect_one()) umma_commit(&d4_full[b4]);
            }
            if (elect_one()) umma_commit(&in4_empty[buf]);
            __syncwarp();
        }
        if (prof_on) {
            atomicAdd(a.prof + 2, (unsigned long long)pc[0]); atomicAdd(a.prof + 3, (unsigned long long)pc[1]);
            atomicAdd(a.prof + 4, (unsigned long long)(D3_T() - mma_t0)); atomicAdd(a.prof + 10, (unsigned long long)n_img);
        }
    } else if (warp == 17) {
        // =================== UMMA issuer, conv 5 ===================
        const uint64_t dw = umma_desc_nosw(smem_u32(w5), D3_Z_CHUNK, 128);
        const uint64_t din = umma_desc_nosw(smem_u32(in5), D3_PLANE_BYTES, D3_QP * 16);
        const uint32_t a_hi = (uint32_t)(din >> 32), b_hi = (uint32_t)(dw >> 32), b_lo = (uint32_t)dw;
        const long long mma_t0 = D3_T();
        uint32_t n5 = 0;
        for (int n = 0; n < n_img; ++n) {
            const int buf = n & 1;
            long long w0 = D3_T();
            mbar_wait(&feat_full[buf], (n >> 1) & 1);
            pc[0] += D3_T() - w0;
            tc_fence_after();
            const uint32_t a_lo = (uint32_t)din + (uint32_t)((buf * D3_BUF_BYTES) >> 4);
#pragma unroll
            for (int T = 0; T < 2; ++T, ++n5) {
                const int b5 = n5 & (D3_NB - 1);
                long long w1 = D3_T();
                if (n5 >= D3_NB) mbar_wait(&d5_empty[b5], ((n5 / D3_NB) - 1) & 1);
                pc[1] += D3_T() - w1;
                tc_fence_after();
                d3_conv_tile(a_lo + T * 16 * D3_QP, a_hi, b_lo, b_hi, tmem + TM_D5 + b5 * 64);
                if (elect_one()) umma_commit(&d5_full[b5]);
            }
            if (elect_one()) umma_commit(&m5_done[buf]);
            __syncwarp();
        }
        if (prof_on) {
            atomicAdd(a.prof + 13, (unsigned long long)pc[0]); atomicAdd(a.prof + 14, (unsigned long long)pc[1]);
            atomicAdd(a.prof + 15, (unsigned long long)(D3_T() - mma_t0));
        }
    } else if (warp < 4 || warp >= 13) {
        // =================== epilogue 1: + skip term, tanh -> bf16 feature quads into in5's phase planes ===============
        // two groups of four warps: warps 0-3 take tile 0 of every image, warps 13-16 tile 1.  A thread always
        // owns the same quad position, so its 64 skip-term values live in registers (32 x bf16x2).
        const int T = warp < 4 ? 0 : 1;
        const int q = warp & 3;                      // TMEM lane quadrant this warp may read
        const int m = q * 32 + lane;                 // GEMM row inside the tile: image row m/8 of the tile, quad m%8
        const int oy = 16 * T + (m >> 3), c = m & 7;
        uint4 sk[8];
        int loaded_cand = -1;
        for (int n = 0; n < n_img; ++n) {
            const int want = a.s4_stride == 0 ? 0 : (list ? __ldg(a.row_cand + a.img_base + img0 + n) : (img0 + n) / a.n_slots);
            if (want != loaded_cand) {
                const uint4* src = reinterpret_cast<const uint4*>(a.s4 + (size_t)want * a.s4_stride + (size_t)(T * 128 + m) * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i) sk[i] = __ldg(src + i);
                loaded_cand = want;
            }
            const uint32_t n4 = 2 * n + T;
            const int b4 = n4 & (D3_NB - 1);
            long long w0 = D3_T();
            mbar_wait(&d4_full[b4], (n4 / D3_NB) & 1);
            if (n >= 2) mbar_wait(&m5_done[n & 1], ((n >> 1) - 1) & 1);   // conv 5 of image n-2 is done with this in5 buffer
            long long w1 = D3_T();
            pc[0] += w1 - w0;
            tc_fence_after();
            uint8_t* dst = in5 + (n & 1) * D3_BUF_BYTES;
            const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + b4 * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float d[32];
                __syncwarp();
                tmem_ld32(t0 + 32 * half, d);
                if (half == 1) {
                    tc_fence_before();
                    mbar_arrive(&d4_empty[b4]);
                }
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2) {          // columns 16jj..16jj+15 = output pixel j' = 3 - jj of the quad
                    const int jj = 2 * half + j2;
                    float sa[8], sb[8];
                    unpack8(sk[2 * jj], sa);
                    unpack8(sk[2 * jj + 1], sb);
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        f[e] = tanh_approx(d[16 * j2 + e] + sa[e]);
                        f[8 + e] = tanh_approx(d[16 * j2 + 8 + e] + sb[e]);
                    }
                    uint4 u0, u1;
                    u0.x = pack_bf16x2(f[0], f[1]);   u0.y = pack_bf16x2(f[2], f[3]);
                    u0.z = pack_bf16x2(f[4], f[5]);   u0.w = pack_bf16x2(f[6], f[7]);
                    u1.x = pack_bf16x2(f[8], f[9]);   u1.y = pack_bf16x2(f[10], f[11]);
                    u1.z = pack_bf16x2(f[12], f[13]); u1.w = pack_bf16x2(f[14], f[15]);
                    // output pixel x = 4c + j' sits at padded (Y, X) = (oy+1, x+1): phase (j'+1)&3, quad c + (j'==3)
                    const int jp = 3 - jj;
                    const int ph = (jp + 1) & 3, row = (oy + 1) * D3_QP + c + (jp == 3 ? 1 : 0);
                    *reinterpret_cast<uint4*>(dst + (ph * 2) * D3_PLANE_BYTES + row * 16) = u0;
                    *reinterpret_cast<uint4*>(dst + (ph * 2 + 1) * D3_PLANE_BYTES + row * 16) = u1;
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&feat_full[n & 1]);
            pc[1] += D3_T() - w1;
        }
        if (prof_on) { atomicAdd(a.prof + (T ? 11 : 6), (unsigned long long)pc[0]); atomicAdd(a.prof + (T ? 12 : 7), (unsigned long long)pc[1]); }
    } else {
        // =================== epilogue 2: mixture mean -> image ===================
        const int q = warp - 4;
        const int m = q * 32 + lane;
        uint32_t n5 = 0;
        constexpr int NCH = HEAD == 0 ? 15 : (HEAD == 1 ? 6 : 16);
        float bh[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) bh[i] = bias5[i];
        // fused L2 cost: this thread always owns the same 4-pixel quad of each of the two tiles -> its 24 goal values
        // stay in registers; per-warp partial sums meet in shared memory (two slots, alternating by image)
        const bool l2 = HEAD == 0 && a.frame_sq != nullptr;
        float gl[2][3][4];
        float* sq_part = reinterpret_cast<float*>(smem + D3_OFF_BAR + 208);          // 8 floats behind the barriers / TMEM holder
        if (l2) {
#pragma unroll
            for (int T = 0; T < 2; ++T)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(a.l2_goal + k * 1024 + (16 * T + (m >> 3)) * 32 + 4 * (m & 7)));
                    gl[T][k][0] = g.x; gl[T][k][1] = g.y; gl[T][k][2] = g.z; gl[T][k][3] = g.w;
                }
        }
        for (int n = 0; n < n_img; ++n) {
            const int img = img0 + n;
            int cand, sl, node;
            if (list) {
                cand = __ldg(a.row_cand + a.img_base + img);
                node = __ldg(a.row_node + a.img_base + img);
                sl = 0;
            } else {
                cand = img / a.n_slots;
                sl = img - cand * a.n_slots;
                node = a.slot0 + sl * (1 + a.slot_extra) - 1;
            }
            float* out = a.images != nullptr ? a.images + ((size_t)cand * a.n_nodes + max(node, 0)) * 3072 : nullptr;
            float sq = 0.f;
            const float* src0 = HEAD == 1 ? a.src0 + (size_t)cand * a.src_stride : nullptr;
            const float* srcg = HEAD == 1 ? a.srcg + (size_t)cand * a.src_stride : nullptr;
            for (int T = 0; T < 2; ++T, ++n5) {
                const int b5 = n5 & (D3_NB - 1);
                long long w0 = D3_T();
                mbar_wait(&d5_full[b5], (n5 / D3_NB) & 1);
                long long w1 = D3_T();
                pc[0] += w1 - w0;
                tc_fence_after();
                const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + TM_D5 + b5 * 64;
                float rgb[3][4];
                float mk[2][4];          // HEAD 1: softmax weights of the start / goal image per pixel of the quad
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float d[32];
                    __syncwarp();
                    tmem_ld32(t0 + 32 * half, d);
                    if (half == 1) {
                        tc_fence_before();
                        mbar_arrive(&d5_empty[b5]);
                    }
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        if (HEAD == 2) {
                            // pixel 4c + px of image row oy; 16 channels = 64 contiguous bytes
                            const int px = 3 - (2 * half + j2);
                            float4* rp = reinterpret_cast<float4*>(
                                a.raw + ((((size_t)cand * a.n_slots + sl) * 1024) + (16 * T + (m >> 3)) * 32 + 4 * (m & 7) + px) * 16);
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                rp[v] = make_float4(d[16 * j2 + 4 * v] + bh[4 * v], d[16 * j2 + 4 * v + 1] + bh[4 * v + 1],
                                                    d[16 * j2 + 4 * v + 2] + bh[4 * v + 2], d[16 * j2 + 4 * v + 3] + bh[4 * v + 3]);
                        } else if (HEAD == 0) {
                            float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
                            for (int ch = 0; ch < 15; ++ch) s[ch % 3] += tanh_approx(d[16 * j2 + ch] + bh[ch]);
#pragma unroll
                            for (int k = 0; k < 3; ++k) rgb[k][3 - (2 * half + j2)] = 0.2f * s[k];
                        } else {
                            const int px = 3 - (2 * half + j2);
                            const float l0 = d[16 * j2 + 3] + bh[3], l1 = d[16 * j2 + 4] + bh[4], l2 = d[16 * j2 + 5] + bh[5];
                            const float mx = fmaxf(l0, fmaxf(l1, l2));
                            const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
                            const float inv = __fdividef(1.0f, e0 + e1 + e2);
                            mk[0][px] = e0 * inv;
                            mk[1][px] = e1 * inv;
#pragma unroll
                            for (int k = 0; k < 3; ++k) rgb[k][px] = e2 * inv * tanh_approx(d[16 * j2 + k] + bh[k]);
                        }
                    }
                }
                if (HEAD == 2) {
                    pc[1] += D3_T() - w1;
                    continue;
                }
                const int oy = 16 * T + (m >> 3), c = m & 7;
                if (HEAD == 1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float4 p0 = __ldg(reinterpret_cast<const float4*>(src0 + k * 1024 + oy * 32 + 4 * c));
                        const float4 pg = __ldg(reinterpret_cast<const float4*>(srcg + k * 1024 + oy * 32 + 4 * c));
                        rgb[k][0] += mk[0][0] * p0.x + mk[1][0] * pg.x;
                        rgb[k][1] += mk[0][1] * p0.y + mk[1][1] * pg.y;
                        rgb[k][2] += mk[0][2] * p0.z + mk[1][2] * pg.z;
                        rgb[k][3] += mk[0][3] * p0.w + mk[1][3] * pg.w;
                    }
                }
                if (node >= 0 && out != nullptr) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        *reinterpret_cast<float4*>(out + k * 1024 + oy * 32 + 4 * c) = make_float4(rgb[k][0], rgb[k][1], rgb[k][2], rgb[k][3]);
                }
                if (l2) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const float d = rgb[k][x] - (T == 0 ? gl[0][k][x] : gl[1][k][x]);
                            sq = fmaf(d, d, sq);
                        }
                }
                pc[1] += D3_T() - w1;
            }
            if (l2) {
                // fixed-order reduction: butterfly inside the warp, then the four warps in order
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                float* slot = sq_part + 4 * (n & 1);
                if (lane == 0) slot[q] = sq;
                named_bar_sync(2, 128);
                if (m == 0 && node >= 0)
                    a.frame_sq[list ? (size_t)(a.img_base + img) : (size_t)cand * a.n_nodes + node] =
                        ((slot[0] + slot[1]) + slot[2]) + slot[3];
            }
        }
        if (prof_on) { atomicAdd(a.prof + 8, (unsigned long long)pc[0]); atomicAdd(a.prof + 9, (unsigned long long)pc[1]); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

__global__ void __launch_bounds__(D3_THREADS, 1) dec_tail3_kernel(const __grid_constant__ DecTail3Args a) {
    dec_tail3_body<0>(a);
}
__global__ void __launch_bounds__(D3_THREADS, 1) dec_tail3_pc_kernel(const __grid_constant__ DecTail3Args a) {
    dec_tail3_body<1>(a);
}
__global__ void __launch_bounds__(D3_THREADS, 1) dec_tail3_raw_kernel(const __grid_constant__ DecTail3Args a) {
    dec_tail3_body<2>(a);
}

// per-candidate skip term of the 32->16 conv in the quad layout of dec_tail3:
//   S4[quad m = oy*8 + c][n = (3-j')*16 + co] = b4[co] + conv4_skip(up(s0))[oy][4c + j'],  bf16.
// skip_up: [n][2][DT_PSTRIDE][8] (skip_prep_kernel, 35-wide padded pixels); w4p: plain [16][32][4][4]; grid (8, n_cand).
__global__ void __launch_bounds__(128) skip_term3_kernel(const bf16* __restrict__ skip_up, const bf16* __restrict__ w4p,
                                                         const float* __restrict__ b4, bf16* __restrict__ s4) {
    __shared__ float w[16 * 16 * 16];     // [co][ci][tap] of the skip half
    const int cand = blockIdx.y;
    for (int i = threadIdx.x; i < 4096; i += 128) {
        const int co = i >> 8, ci = (i >> 4) & 15, tap = i & 15;
        w[i] = __bfloat162float(w4p[(co * 32 + 16 + ci) * 16 + tap]);
    }
    __syncthreads();
    const int px = blockIdx.x * 128 + threadIdx.x;        // output pixel 0..1023
    const int oy = px >> 5, ox = px & 31;
    const bf16* src = skip_up + (size_t)cand * 2 * DT_PSTRIDE * 8;
    float acc[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) acc[co] = b4[co];
    for (int tap = 0; tap < 16; ++tap) {
        const int pix = (oy + (tap >> 2)) * DT_WP + ox + (tap & 3);
        for (int ci = 0; ci < 16; ++ci) {
            const float v = __bfloat162float(src[((size_t)(ci >> 3) * DT_PSTRIDE + pix) * 8 + (ci & 7)]);
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[co] = fmaf(v, w[(co * 16 + ci) * 16 + tap], acc[co]);
        }
    }
    bf16* out = s4 + ((size_t)cand * 256 + oy * 8 + (ox >> 2)) * 64 + (3 - (ox & 3)) * 16;
    for (int co = 0; co < 16; ++co) out[co] = __float2bfloat16_rn(acc[co]);
}

}  // namespace gcp
