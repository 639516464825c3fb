// Decoder tail v2: warp-specialised, tile-pipelined implicit-GEMM kernel for the two full-resolution
// convolutions of the GCP decoder (see dec_tail.cuh for the math and the operand-plane layout).
//
// What v1 taught (profiles/r1_dec_tail_v1.md): one tcgen05.mma with M=128 costs >= ~64 cycles for its
// A-operand read no matter how small N is, so 16 taps x N=16 wastes the tensor pipe.  v2 therefore
//   * folds the 4 horizontal taps into N:   D[p][(kx,co)] = sum_{ky,ci} X[p + ky*35][ci] W[co][ci][ky][kx]
//     (4 MMAs per 128-pixel tile instead of 32 / 16), and finishes the convolution in the epilogue with a
//     row-shifted sum   out[p][co] = sum_kx D[p + kx][(kx,co)]   done with warp shuffles (tiles carry a
//     3-row halo, quadrant boundaries are patched through a few shared-memory rows);
//   * splits the skip half of the 32->16 conv off as a per-candidate constant S4[p][co] (the conv is linear
//     in its input and the skip is the same for all 255 nodes of a candidate), halving K;
//   * pipelines at tile granularity across four warp roles (up-sampler, UMMA issuer, epilogue-1,
//     epilogue-2) with mbarriers, double-buffered TMEM accumulators and double-buffered input planes, so
//     up-sampling, both convolutions and both epilogues of neighbouring tiles / images overlap.
#pragma once
#include "dec_tail.cuh"

namespace gcp {

constexpr int D2_TILE = 125;                   // output pixels finished per 128-row tile (3-row halo)
constexpr int D2_TILES = 9;                    // 9 * 125 = 1125 >= 1120
constexpr int D2_IN4_BYTES = 2 * DT_PLANE_BYTES;   // 16 channels
constexpr int D2_IN5_BYTES = 2 * DT_PLANE_BYTES;
constexpr int D2_S4_BYTES = 1152 * 16 * 2;     // bf16 [p][co]
constexpr int D2_W4_BYTES = 4 * 2048;          // [ky][kchunk 2][n 64 = (kx,co)][8]
constexpr int D2_W5_BYTES = 4 * 4096;          // [ky][kchunk 2][n 128 = (kx,co32)][8]
constexpr int D2_X3_BYTES = 8192;
constexpr int D2_XCH_BYTES = 2 * 3 * 3 * 48 * 4;   // [parity][warp boundary][lane][48 floats]
constexpr int D2_OFF_IN4 = 0;
constexpr int D2_OFF_IN5 = D2_OFF_IN4 + 2 * D2_IN4_BYTES;
constexpr int D2_OFF_S4 = D2_OFF_IN5 + D2_IN5_BYTES;
constexpr int D2_OFF_W4 = D2_OFF_S4 + D2_S4_BYTES;
constexpr int D2_OFF_W5 = D2_OFF_W4 + D2_W4_BYTES;
constexpr int D2_OFF_X3 = D2_OFF_W5 + D2_W5_BYTES;
constexpr int D2_OFF_XCH4 = D2_OFF_X3 + D2_X3_BYTES;
constexpr int D2_OFF_XCH5 = D2_OFF_XCH4 + D2_XCH_BYTES;
constexpr int D2_OFF_BIAS = D2_OFF_XCH5 + D2_XCH_BYTES;
constexpr int D2_OFF_BAR = D2_OFF_BIAS + 64;
constexpr int D2_SMEM_BYTES = D2_OFF_BAR + 256 + 128;
constexpr int D2_THREADS = 13 * 32;            // warps 0-3 epilogue-1, 4-7 epilogue-2, 8-11 up-sampler, 12 UMMA

struct DecTail2Args {
    const bf16* x3;        // [n_slots * Bp][4096]  rows (slot_local, cand); [plane 2][y16][x16][8]
    const bf16* s4;        // [n_cand or 1][1152][16] bf16: skip half of conv 32->16 incl. its bias
    int s4_stride;         // elements between candidates (0: shared)
    const bf16* w4;        // D2_W4_BYTES, x-half of the 32->16 conv, kx folded into N
    const bf16* w5;        // D2_W5_BYTES
    const float* b5;       // [32]
    float* images;         // [B][n_nodes][3][32][32]
    int Bp, n_cand, slot0, n_slots, n_nodes, slots_per_unit;
    unsigned long long* prof;   // optional [8]: issuer wait cycles etc.
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

#define D2_T() (prof_on ? clock64() : 0ll)

__global__ void __launch_bounds__(D2_THREADS, 1) dec_tail2_kernel(const __grid_constant__ DecTail2Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* in4 = smem + D2_OFF_IN4;
    uint8_t* in5 = smem + D2_OFF_IN5;
    bf16* s4 = reinterpret_cast<bf16*>(smem + D2_OFF_S4);
    uint8_t* w4 = smem + D2_OFF_W4;
    uint8_t* w5 = smem + D2_OFF_W5;
    uint8_t* x3s = smem + D2_OFF_X3;
    float* xch4 = reinterpret_cast<float*>(smem + D2_OFF_XCH4);
    float* xch5 = reinterpret_cast<float*>(smem + D2_OFF_XCH5);
    float* bias5 = reinterpret_cast<float*>(smem + D2_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D2_OFF_BAR);
    uint64_t* x3_full = bars + 0;       // tx
    uint64_t* in4_full = bars + 1;      // [2] count 128 (up-sampler threads)
    uint64_t* in4_empty = bars + 3;     // [2] tcgen05.commit
    uint64_t* d4_full = bars + 5;       // [2] commit
    uint64_t* d4_empty = bars + 7;      // [2] count 128
    uint64_t* d5_full = bars + 9;       // [2] commit
    uint64_t* d5_empty = bars + 11;     // [2] count 128
    uint64_t* f_full = bars + 13;       // [4] feature tile g written (count 128): barrier g&3, phase g>>2
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (2 * D2_IN4_BYTES + D2_IN5_BYTES) / 16; i += D2_THREADS)
        reinterpret_cast<uint4*>(in4)[i] = make_uint4(0, 0, 0, 0);     // padding rings stay zero forever
    for (int i = tid; i < D2_W4_BYTES / 16; i += D2_THREADS) reinterpret_cast<uint4*>(w4)[i] = __ldg(reinterpret_cast<const uint4*>(a.w4) + i);
    for (int i = tid; i < D2_W5_BYTES / 16; i += D2_THREADS) reinterpret_cast<uint4*>(w5)[i] = __ldg(reinterpret_cast<const uint4*>(a.w5) + i);
    if (tid < 16) bias5[tid] = a.b5[tid];
    if (tid == 0) {
        mbar_init(x3_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&in4_full[i], 128);
            mbar_init(&in4_empty[i], 1);
            mbar_init(&d4_full[i], 1);
            mbar_init(&d4_empty[i], 128);
            mbar_init(&d5_full[i], 1);
            mbar_init(&d5_empty[i], 128);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&f_full[i], 128);
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc(tmem_holder, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    constexpr uint32_t TM_D5 = 128;     // D4: 2 x 64 columns, D5: 2 x 128 columns

    const int units_per_cand = (a.n_slots + a.slots_per_unit - 1) / a.slots_per_unit;
    const int n_units = a.n_cand * units_per_cand;
    const bool prof_on = (a.prof != nullptr) && lane == 0 && (warp == 0 || warp == 4 || warp == 8 || warp == 12);
    long long pc[5] = {0, 0, 0, 0, 0};

    if (warp >= 8 && warp < 12) {
        // =================== up-sampler: x3 (TMA bulk -> staging) -> bilinear x2 -> in4[buf] ===================
        const int bt = tid - 256;            // 0..127
        uint32_t img = 0;
        // prefetch of the very first image
        if (bt == 0 && blockIdx.x < n_units) {
            const int cand = blockIdx.x / units_per_cand;
            const int sl = (blockIdx.x - cand * units_per_cand) * a.slots_per_unit;
            mbar_arrive_expect_tx(x3_full, D2_X3_BYTES);
            bulk_load_1d(x3s, a.x3 + ((size_t)sl * a.Bp + cand) * 4096, D2_X3_BYTES, x3_full);
        }
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int cand = unit / units_per_cand;
            const int s_begin = (unit - cand * units_per_cand) * a.slots_per_unit;
            const int s_end = min(s_begin + a.slots_per_unit, a.n_slots);
            for (int sl = s_begin; sl < s_end; ++sl, ++img) {
                const int buf = img & 1;
                long long t0 = D2_T();
                mbar_wait(x3_full, img & 1);
                long long t1 = D2_T();
                if (img >= 2) mbar_wait(&in4_empty[buf], ((img >> 1) - 1) & 1);
                long long t2 = D2_T();
                pc[0] += t1 - t0; pc[1] += t2 - t1;
                uint8_t* dst = in4 + buf * D2_IN4_BYTES;
                const uint4* src = reinterpret_cast<const uint4*>(x3s);
                for (int it = bt; it < 2048; it += 128) {
                    const int plane = it >> 10, oy = (it >> 5) & 31, ox = it & 31;
                    int y0, y1, x0, x1;
                    float wy0, wy1, wx0, wx1;
                    up2_src(oy, 16, y0, y1, wy0, wy1);
                    up2_src(ox, 16, x0, x1, wx0, wx1);
                    float p00[8], p01[8], p10[8], p11[8];
                    unpack8(src[plane * 256 + y0 * 16 + x0], p00);
                    unpack8(src[plane * 256 + y0 * 16 + x1], p01);
                    unpack8(src[plane * 256 + y1 * 16 + x0], p10);
                    unpack8(src[plane * 256 + y1 * 16 + x1], p11);
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] = wy0 * (wx0 * p00[k] + wx1 * p01[k]) + wy1 * (wx0 * p10[k] + wx1 * p11[k]);
                    uint4 u;
                    u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
                    u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(dst + plane * DT_PLANE_BYTES + ((oy + 1) * DT_WP + ox + 1) * 16) = u;
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);               // every up-sampler thread is done with the staging tile
                if (bt == 0) {
                    // prefetch the next image of this CTA (if any)
                    int nsl = sl + 1, ncand = cand;
                    bool more = true;
                    if (nsl >= s_end) {
                        const int nunit = unit + gridDim.x;
                        if (nunit < n_units) {
                            ncand = nunit / units_per_cand;
                            nsl = (nunit - ncand * units_per_cand) * a.slots_per_unit;
                        } else {
                            more = false;
                        }
                    }
                    if (more) {
                        mbar_arrive_expect_tx(x3_full, D2_X3_BYTES);
                        bulk_load_1d(x3s, a.x3 + ((size_t)nsl * a.Bp + ncand) * 4096, D2_X3_BYTES, x3_full);
                    }
                }
                mbar_arrive(&in4_full[buf]);
                pc[2] += D2_T() - t2;
            }
        }
        if (prof_on) { atomicAdd(a.prof + 0, (unsigned long long)pc[0]); atomicAdd(a.prof + 1, (unsigned long long)pc[1]); atomicAdd(a.prof + 2, (unsigned long long)pc[2]); }
    } else if (warp == 12) {
        if (lane == 0) {
            // =================== UMMA issuer ===================
            constexpr uint32_t idesc4 = umma_idesc_bf16(128, 64);
            constexpr uint32_t idesc5 = umma_idesc_bf16(128, 128);
            const uint32_t w4_base = smem_u32(w4), w5_base = smem_u32(w5), in5_base = smem_u32(in5);
            uint32_t img = 0, n4 = 0, n5 = 0, feat_seen = 0;
            auto issue5 = [&](int t) {
                const int b5 = n5 & 1;
                long long w0 = D2_T();
                if (n5 >= 2) mbar_wait(&d5_empty[b5], ((n5 >> 1) - 1) & 1);
                pc[3] += D2_T() - w0;
                tc_fence_after();
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const uint64_t da = umma_desc_nosw(in5_base + (t * D2_TILE + ky * DT_WP) * 16, DT_PLANE_BYTES, 128);
                    const uint64_t db = umma_desc_nosw(w5_base + ky * 4096, 128 * 16, 128);
                    umma_bf16(tmem + TM_D5 + b5 * 128, da, db, idesc5, ky != 0);
                }
                umma_commit(&d5_full[b5]);
                ++n5;
            };
            auto wait_feat = [&](uint32_t upto) {      // feature tiles [0, upto) of the whole stream are in in5
                long long w0 = D2_T();
                while (feat_seen < upto) {
                    mbar_wait(&f_full[feat_seen & 3], (feat_seen >> 2) & 1);
                    ++feat_seen;
                }
                pc[2] += D2_T() - w0;
                tc_fence_after();
            };
            const long long mma_t0 = D2_T();
            // Issue order: M4(g), then M5(g-2) (two tiles behind).  The lag gives epilogue 1 a full tile period to turn
            // D4(g-1) into feature rows before the second conv needs them, so the tensor pipe never waits on it.
            uint32_t g = 0;
            auto issue5_lagged = [&](uint32_t g5) {
                const uint32_t i5 = g5 / D2_TILES, t5 = g5 - i5 * D2_TILES;
                wait_feat(i5 * D2_TILES + min(t5 + 1, (uint32_t)(D2_TILES - 1)) + 1);
                issue5((int)t5);
            };
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const int cand = unit / units_per_cand;
                const int s_begin = (unit - cand * units_per_cand) * a.slots_per_unit;
                const int s_end = min(s_begin + a.slots_per_unit, a.n_slots);
                for (int sl = s_begin; sl < s_end; ++sl, ++img) {
                    const int buf = img & 1;
                    long long w0 = D2_T();
                    mbar_wait(&in4_full[buf], (img >> 1) & 1);
                    pc[0] += D2_T() - w0;
                    tc_fence_after();
                    const uint32_t in4_base = smem_u32(in4 + buf * D2_IN4_BYTES);
                    for (int t = 0; t < D2_TILES; ++t) {
                        const int b4 = n4 & 1;
                        long long w1 = D2_T();
                        if (n4 >= 2) mbar_wait(&d4_empty[b4], ((n4 >> 1) - 1) & 1);
                        pc[1] += D2_T() - w1;
                        tc_fence_after();
#pragma unroll
                        for (int ky = 0; ky < 4; ++ky) {
                            const uint64_t da = umma_desc_nosw(in4_base + (t * D2_TILE + ky * DT_WP) * 16, DT_PLANE_BYTES, 128);
                            const uint64_t db = umma_desc_nosw(w4_base + ky * 2048, 64 * 16, 128);
                            umma_bf16(tmem + b4 * 64, da, db, idesc4, ky != 0);
                        }
                        umma_commit(&d4_full[b4]);
                        ++n4;
                        if (t == D2_TILES - 1) umma_commit(&in4_empty[buf]);
                        ++g;                                   // M4(g-1) just issued
                        if (g >= 3) issue5_lagged(g - 3);
                    }
                }
            }
            if (g >= 2) issue5_lagged(g - 2);      // inside the loop M5(0..g-3) were issued;
            if (g >= 1) issue5_lagged(g - 1);      // the last two tiles drain here
            if (prof_on) {
                atomicAdd(a.prof + 3, (unsigned long long)pc[0]); atomicAdd(a.prof + 4, (unsigned long long)pc[1]);
                atomicAdd(a.prof + 5, (unsigned long long)pc[2]); atomicAdd(a.prof + 6, (unsigned long long)pc[3]);
                atomicAdd(a.prof + 7, (unsigned long long)(D2_T() - mma_t0)); atomicAdd(a.prof + 13, (unsigned long long)img);
            }
        }
    } else if (warp < 4) {
        // =================== epilogue 1: shifted sum + skip term + tanh -> in5 planes ===================
        const int q = warp;
        uint32_t n4 = 0;
        int loaded_cand = -1;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int cand = unit / units_per_cand;
            const int s_begin = (unit - cand * units_per_cand) * a.slots_per_unit;
            const int s_end = min(s_begin + a.slots_per_unit, a.n_slots);
            const int want = a.s4_stride == 0 ? 0 : cand;
            if (want != loaded_cand) {
                named_bar_sync(2, 128);       // previous unit's tiles no longer read s4
                const uint4* src = reinterpret_cast<const uint4*>(a.s4 + (size_t)want * a.s4_stride);
                for (int i = tid; i < D2_S4_BYTES / 16; i += 128) reinterpret_cast<uint4*>(s4)[i] = __ldg(src + i);
                named_bar_sync(2, 128);
                loaded_cand = want;
            }
            for (int sl = s_begin; sl < s_end; ++sl) {
                for (int t = 0; t < D2_TILES; ++t, ++n4) {
                    const int b4 = n4 & 1;
                    long long w0 = D2_T();
                    mbar_wait(&d4_full[b4], (n4 >> 1) & 1);
                    long long w1 = D2_T();
                    pc[0] += w1 - w0;
                    tc_fence_after();
                    float d[64];
                    {
                        float lo[32], hi[32];
                        __syncwarp();
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + b4 * 64, lo);
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + b4 * 64 + 32, hi);
#pragma unroll
                        for (int i = 0; i < 32; ++i) { d[i] = lo[i]; d[32 + i] = hi[i]; }
                    }
                    tc_fence_before();
                    mbar_arrive(&d4_empty[b4]);
                    float* xc = xch4 + (n4 & 1) * (3 * 3 * 48);
                    if (q > 0 && lane < 3) {
#pragma unroll
                        for (int i = 0; i < 48; ++i) xc[((q - 1) * 3 + lane) * 48 + i] = d[16 + i];
                    }
                    long long w2 = D2_T();
                    named_bar_sync(3, 128);
                    pc[2] += D2_T() - w2;
                    float v1[16], v2[16], v3[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        v1[c] = __shfl_down_sync(0xffffffffu, d[16 + c], 1);
                        v2[c] = __shfl_down_sync(0xffffffffu, d[32 + c], 2);
                        v3[c] = __shfl_down_sync(0xffffffffu, d[48 + c], 3);
                    }
                    if (q < 3 && lane >= 29) {       // rows whose +kx neighbours live in the next warp's quadrant
                        const float4* x3p = reinterpret_cast<const float4*>(xc + (q * 3 + lane - 29) * 48 + 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float4 f = x3p[i]; v3[4 * i] = f.x; v3[4 * i + 1] = f.y; v3[4 * i + 2] = f.z; v3[4 * i + 3] = f.w; }
                        if (lane >= 30) {
                            const float4* x2p = reinterpret_cast<const float4*>(xc + (q * 3 + lane - 30) * 48 + 16);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float4 f = x2p[i]; v2[4 * i] = f.x; v2[4 * i + 1] = f.y; v2[4 * i + 2] = f.z; v2[4 * i + 3] = f.w; }
                        }
                        if (lane == 31) {
                            const float4* x1p = reinterpret_cast<const float4*>(xc + (q * 3) * 48);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float4 f = x1p[i]; v1[4 * i] = f.x; v1[4 * i + 1] = f.y; v1[4 * i + 2] = f.z; v1[4 * i + 3] = f.w; }
                        }
                    }
                    __syncwarp();
                    float acc[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c] = d[c] + v1[c] + v2[c] + v3[c];
                    const int r = q * 32 + lane;
                    const int p = t * D2_TILE + r;
                    const int x = p % DT_WP;
                    if (r < D2_TILE && p < DT_NPIX && x < 32) {
                        const uint4* sp = reinterpret_cast<const uint4*>(s4 + p * 16);
                        float sa[8], sb[8];
                        unpack8(sp[0], sa);
                        unpack8(sp[1], sb);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            acc[c] = tanh_approx(acc[c] + sa[c]);
                            acc[8 + c] = tanh_approx(acc[8 + c] + sb[c]);
                        }
                        uint4 u0, u1;
                        u0.x = pack_bf16x2(acc[0], acc[1]);   u0.y = pack_bf16x2(acc[2], acc[3]);
                        u0.z = pack_bf16x2(acc[4], acc[5]);   u0.w = pack_bf16x2(acc[6], acc[7]);
                        u1.x = pack_bf16x2(acc[8], acc[9]);   u1.y = pack_bf16x2(acc[10], acc[11]);
                        u1.z = pack_bf16x2(acc[12], acc[13]); u1.w = pack_bf16x2(acc[14], acc[15]);
                        *reinterpret_cast<uint4*>(in5 + (p + DT_WP + 1) * 16) = u0;
                        *reinterpret_cast<uint4*>(in5 + DT_PLANE_BYTES + (p + DT_WP + 1) * 16) = u1;
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&f_full[n4 & 3]);
                    pc[1] += D2_T() - w1;
                }
            }
        }
        if (prof_on) { atomicAdd(a.prof + 8, (unsigned long long)pc[0]); atomicAdd(a.prof + 9, (unsigned long long)pc[1]); atomicAdd(a.prof + 10, (unsigned long long)pc[2]); }
    } else {
        // =================== epilogue 2: shifted sum + sigmoid mixture mean -> image ===================
        const int q = warp - 4;
        uint32_t n5 = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int cand = unit / units_per_cand;
            const int s_begin = (unit - cand * units_per_cand) * a.slots_per_unit;
            const int s_end = min(s_begin + a.slots_per_unit, a.n_slots);
            for (int sl = s_begin; sl < s_end; ++sl) {
                const int node = a.slot0 + sl - 1;
                float* img = a.images + ((size_t)cand * a.n_nodes + node) * 3072;
                for (int t = 0; t < D2_TILES; ++t, ++n5) {
                    const int b5 = n5 & 1;
                    long long w0 = D2_T();
                    mbar_wait(&d5_full[b5], (n5 >> 1) & 1);
                    long long w1 = D2_T();
                    pc[0] += w1 - w0;
                    tc_fence_after();
                    float e[4][16];
                    __syncwarp();
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx)
                        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + TM_D5 + b5 * 128 + kx * 32, e[kx]);
                    tc_fence_before();
                    mbar_arrive(&d5_empty[b5]);
                    float* xc = xch5 + (n5 & 1) * (3 * 3 * 48);
                    if (q > 0 && lane < 3) {
#pragma unroll
                        for (int kx = 1; kx < 4; ++kx)
#pragma unroll
                            for (int c = 0; c < 16; ++c) xc[((q - 1) * 3 + lane) * 48 + (kx - 1) * 16 + c] = e[kx][c];
                    }
                    named_bar_sync(4, 128);
                    float v1[16], v2[16], v3[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        v1[c] = __shfl_down_sync(0xffffffffu, e[1][c], 1);
                        v2[c] = __shfl_down_sync(0xffffffffu, e[2][c], 2);
                        v3[c] = __shfl_down_sync(0xffffffffu, e[3][c], 3);
                    }
                    if (q < 3 && lane >= 29) {
                        const float4* x3p = reinterpret_cast<const float4*>(xc + (q * 3 + lane - 29) * 48 + 32);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float4 f = x3p[i]; v3[4 * i] = f.x; v3[4 * i + 1] = f.y; v3[4 * i + 2] = f.z; v3[4 * i + 3] = f.w; }
                        if (lane >= 30) {
                            const float4* x2p = reinterpret_cast<const float4*>(xc + (q * 3 + lane - 30) * 48 + 16);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float4 f = x2p[i]; v2[4 * i] = f.x; v2[4 * i + 1] = f.y; v2[4 * i + 2] = f.z; v2[4 * i + 3] = f.w; }
                        }
                        if (lane == 31) {
                            const float4* x1p = reinterpret_cast<const float4*>(xc + (q * 3) * 48);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float4 f = x1p[i]; v1[4 * i] = f.x; v1[4 * i + 1] = f.y; v1[4 * i + 2] = f.z; v1[4 * i + 3] = f.w; }
                        }
                    }
                    __syncwarp();
                    float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 15; ++c) {
                        const float sacc = e[0][c] + v1[c] + v2[c] + v3[c] + bias5[c];
                        rgb[c % 3] += 0.5f * tanh_approx(0.5f * sacc) + 0.5f;      // sigmoid
                    }
                    const int r = q * 32 + lane;
                    const int p = t * D2_TILE + r;
                    const int y = p / DT_WP, x = p - y * DT_WP;
                    if (r < D2_TILE && p < DT_NPIX && x < 32) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) img[k * 1024 + y * 32 + x] = rgb[k] * 0.4f - 1.0f;
                    }
                    pc[1] += D2_T() - w1;
                }
            }
        }
        if (prof_on) { atomicAdd(a.prof + 11, (unsigned long long)pc[0]); atomicAdd(a.prof + 12, (unsigned long long)pc[1]); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// per-candidate skip term of the 32->16 conv: S4[p][co] = b4[co] + conv4_skip(up(s0))[p], bf16.
// skip_up: [n][2][DT_PSTRIDE][8] (from skip_prep_kernel); w4p: plain [16][32][4][4]; grid (9, n_cand).
__global__ void __launch_bounds__(128) skip_term_kernel(const bf16* __restrict__ skip_up, const bf16* __restrict__ w4p,
                                                        const float* __restrict__ b4, bf16* __restrict__ s4) {
    __shared__ float w[16 * 16 * 16];     // [co][ci][tap] of the skip half
    const int cand = blockIdx.y;
    for (int i = threadIdx.x; i < 4096; i += 128) {
        const int co = i >> 8, ci = (i >> 4) & 15, tap = i & 15;
        w[i] = __bfloat162float(w4p[(co * 32 + 16 + ci) * 16 + tap]);
    }
    __syncthreads();
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= 1152) return;
    bf16* out = s4 + ((size_t)cand * 1152 + p) * 16;
    const int x = p % DT_WP;
    if (p >= DT_NPIX || x >= 32) {
        for (int co = 0; co < 16; ++co) out[co] = __float2bfloat16_rn(0.f);
        return;
    }
    const bf16* src = skip_up + (size_t)cand * 2 * DT_PSTRIDE * 8;
    float acc[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) acc[co] = b4[co];
    for (int tap = 0; tap < 16; ++tap) {
        const int pix = p + (tap >> 2) * DT_WP + (tap & 3);
        for (int ci = 0; ci < 16; ++ci) {
            const float v = __bfloat162float(src[((size_t)(ci >> 3) * DT_PSTRIDE + pix) * 8 + (ci & 7)]);
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[co] = fmaf(v, w[(co * 16 + ci) * 16 + tap], acc[co]);
        }
    }
    for (int co = 0; co < 16; ++co) out[co] = __float2bfloat16_rn(acc[co]);
}

}  // namespace gcp
