// Fused row-MLP body: the `in` layer (+bias, LeakyReLU) and the three [linear, GroupNorm(8 groups of 16), LeakyReLU] layers
// of a 128-wide BaseProcessingNet (blox/torch/subnetworks.py Predictor; used for the prior, the length / existence /
// distance predictors, the inverse model, the state regressor and the cost model) in ONE launch.
//
// Why: as four gemm_tc_kernel launches the body moves its 128-wide activations through HBM three times and, worse, is a
// chain of four dependent launches -- at tree levels 0-4, in the 199-step sequential rollout, in the hierarchical
// planner (10 candidates) and in the batch-16 training forward that chain is pure launch latency (~9 us per launch).
// Here a CTA keeps the activations of its 128-row tile on chip: layer l's accumulator (TMEM, 128 fp32 columns) is drained
// by the epilogue warps, normalised, rounded to bf16 and written straight into shared memory in the 128B-swizzled K-major
// layout tcgen05.mma reads its A operand from; only the last layer's output goes to HBM.  Rounding points are exactly
// those of the unfused path (bf16 activations between layers), so results are bit-identical to it.
//
// Roles (as gemm_tc_kernel): warp 0 = TMA producer, warp 1 = UMMA issuer + TMEM owner, warps 2-9 = epilogue
// (thread = row, warp half = 64-column half).  The producer streams "steps" of 64 K-columns through a 3-stage ring:
// layer 0 steps carry an A block (input rows, any of the GEMM's row modes / segments) and a weight block, layer 1-3
// steps only a weight block (A is the activation buffer).  One mbarrier (`epi_done`, 256 arrivals) orders everything
// between layers: it says the previous accumulator has been read AND the activation buffer has been rewritten.
#pragma once
#include "gemm.cuh"

namespace gcp {

constexpr int MLPF_STAGES = 3;
constexpr int MLPF_BLK = GEMM_BM * GEMM_BK * 2;          // 16 KB: one 128 x 64 bf16 block
constexpr int MLPF_STAGE_BYTES = 2 * MLPF_BLK;           // A block | W block
constexpr int MLPF_ACT_BYTES = 2 * MLPF_BLK;             // 128 rows x 128 columns bf16
constexpr int MLPF_OFF_ACT = MLPF_STAGES * MLPF_STAGE_BYTES;
constexpr int MLPF_OFF_GN = MLPF_OFF_ACT + MLPF_ACT_BYTES;           // [3 layers][gamma 128 | beta 128] fp32
constexpr int MLPF_OFF_BIAS = MLPF_OFF_GN + 3 * 256 * 4;             // [128] fp32 bias of the in layer
constexpr int MLPF_OFF_BAR = MLPF_OFF_BIAS + 128 * 4;
constexpr int MLPF_SMEM_BYTES = MLPF_OFF_BAR + 128 + 1024 /*align slack*/;
constexpr int MLPF_LAYERS = 4;

struct MlpFusedArgs {
    GemmArgs in;              // layer 0: A maps / segments / row geometry, w_map = in-layer weights (box 128 rows), rows, K
    CUtensorMap w_mid[3];     // mid-layer weights [128][128], box 64 columns x 128 rows
    const float* bias_in;     // [128] or null
    const float* gam[3];
    const float* bet[3];
    bf16* out;                // [rows][out_ld], all 128 columns
    int out_ld;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) mlp_fused_kernel(const __grid_constant__ MlpFusedArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem + MLPF_OFF_ACT;
    float* gn = reinterpret_cast<float*>(smem + MLPF_OFF_GN);
    float* bias_s = reinterpret_cast<float*>(smem + MLPF_OFF_BIAS);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MLPF_OFF_BAR);
    uint64_t* empty_bar = full_bar + MLPF_STAGES;
    uint64_t* acc_full = empty_bar + MLPF_STAGES;
    uint64_t* epi_done = acc_full + 1;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(epi_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmArgs& ia = args.in;
    const int kb_in = ia.K / GEMM_BK;

    // weights only: safe before pdl_wait()
    for (int i = threadIdx.x; i < 3 * 128; i += GEMM_THREADS) {
        const int l = i >> 7, ch = i & 127;
        gn[l * 256 + ch] = __ldg(args.gam[l] + ch);
        gn[l * 256 + 128 + ch] = __ldg(args.bet[l] + ch);
    }
    for (int i = threadIdx.x; i < 128; i += GEMM_THREADS) bias_s[i] = args.bias_in != nullptr ? __ldg(args.bias_in + i) : 0.f;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < ia.n_seg; ++s) tma_prefetch_desc(&ia.a_map[s]);
        tma_prefetch_desc(&ia.w_map);
        for (int l = 0; l < 3; ++l) tma_prefetch_desc(&args.w_mid[l]);
        for (int s = 0; s < MLPF_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(epi_done, 32 * GEMM_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_holder, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();
    pdl_wait();
    int rows = ia.rows;
    if (ia.rows_dev != nullptr)               // pruned tree level: the live row count is on the device (GemmArgs::rows_dev)
        rows = min(rows, (max(__ldg(ia.rows_dev) - ia.rows_dev_base, 0) + GEMM_BM - 1) / GEMM_BM * GEMM_BM);
    const int tiles_m = rows / GEMM_BM;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: the step sequence of every tile of this CTA =====
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
                int kb = 0;
                for (int s = 0; s < ia.n_seg; ++s) {
                    const ASeg& sg = ia.seg[s];
                    const int row0 = tile_row0(ia.g, sg.row_mode, sg.row_mode == ROW_LEVEL ? tile : listed_tile(ia.g, tile)) + sg.row_base;
                    for (int kk = 0; kk < sg.k_len; kk += GEMM_BK, ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = smem + stage * MLPF_STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], 2 * MLPF_BLK);
                        tma_load_2d(st, &ia.a_map[s], &full_bar[stage], sg.col0 + kk, row0);
                        tma_load_2d(st + MLPF_BLK, &ia.w_map, &full_bar[stage], kb * GEMM_BK, 0);
                        if (++stage == MLPF_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                for (int l = 0; l < 3; ++l)
                    for (int k2 = 0; k2 < 2; ++k2) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = smem + stage * MLPF_STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], MLPF_BLK);
                        tma_load_2d(st + MLPF_BLK, &args.w_mid[l], &full_bar[stage], k2 * GEMM_BK, 0);
                        if (++stage == MLPF_STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== UMMA issuer =====
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, 128);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t n_layer = 0;          // layers issued so far (all tiles): layer n waits for epilogue n - 1
            for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
                for (int l = 0; l < MLPF_LAYERS; ++l, ++n_layer) {
                    if (n_layer > 0) {
                        mbar_wait(epi_done, (n_layer - 1) & 1);     // accumulator drained, activation buffer rewritten
                        tc_fence_after();
                    }
                    const int nkb = l == 0 ? kb_in : 2;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * MLPF_STAGE_BYTES);
                        const uint64_t da = umma_desc_sw128(l == 0 ? sa : smem_u32(act + kb * MLPF_BLK));
                        const uint64_t db = umma_desc_sw128(sa + MLPF_BLK);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit(&empty_bar[stage]);
                        if (++stage == MLPF_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(acc_full);
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int r = q * 32 + lane;                      // row inside the tile
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + half * 64;
        uint8_t* arow = act + half * MLPF_BLK + r * 128;  // this thread's 128-byte row piece of the activation buffer
        uint32_t n_layer = 0;
        for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
            for (int l = 0; l < MLPF_LAYERS; ++l, ++n_layer) {
                mbar_wait(acc_full, n_layer & 1);
                tc_fence_after();
                float a0[32], a1[32];
                __syncwarp();
                tmem_ld32_pair(t0, t0 + 32, a0, a1);
                if (l == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        a0[i] = lrelu_(a0[i] + bias_s[half * 64 + i]);
                        a1[i] = lrelu_(a1[i] + bias_s[half * 64 + 32 + i]);
                    }
                } else {
                    const float* g = gn + (l - 1) * 256 + half * 64;
                    group_norm16_chunk_smem(g, g + 128, a0);
                    group_norm16_chunk_smem(g + 32, g + 128 + 32, a1);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        a0[i] = lrelu_(a0[i]);
                        a1[i] = lrelu_(a1[i]);
                    }
                }
                uint4 u[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u[j].x = pack_bf16x2(a0[8 * j + 0], a0[8 * j + 1]); u[j].y = pack_bf16x2(a0[8 * j + 2], a0[8 * j + 3]);
                    u[j].z = pack_bf16x2(a0[8 * j + 4], a0[8 * j + 5]); u[j].w = pack_bf16x2(a0[8 * j + 6], a0[8 * j + 7]);
                    u[4 + j].x = pack_bf16x2(a1[8 * j + 0], a1[8 * j + 1]); u[4 + j].y = pack_bf16x2(a1[8 * j + 2], a1[8 * j + 3]);
                    u[4 + j].z = pack_bf16x2(a1[8 * j + 4], a1[8 * j + 5]); u[4 + j].w = pack_bf16x2(a1[8 * j + 6], a1[8 * j + 7]);
                }
                if (l < MLPF_LAYERS - 1) {
                    // A operand of the next layer: 16-byte chunk j of row r sits at chunk j ^ (r % 8) (128B swizzle)
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) = u[j];
                    fence_proxy_async_smem();
                } else {
                    uint4* o = reinterpret_cast<uint4*>(args.out + ((size_t)tile * GEMM_BM + r) * args.out_ld + half * 64);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = u[j];
                }
                tc_fence_before();
                mbar_arrive(epi_done);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace gcp
