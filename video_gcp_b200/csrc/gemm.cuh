// Row-tiled GEMM with fused epilogues for the TreeLSTM recursion and every row-MLP of the rollout.
//
//   D[rows, N] = concat_K(A_seg0, A_seg1, ...)[rows, K] * Wp[N, K]^T      (bf16 in, fp32 accumulate)
//
// Two implementations share the same argument block and the same epilogue code:
//   * gemm_tc_kernel  -- the product path: persistent, warp-specialised tcgen05 kernel
//                        (warp 0 = TMA producer, warp 1 = UMMA issuer + TMEM owner, warps 2-9 =
//                        epilogue, thread = accumulator row), 128B-swizzled K-major operands staged by
//                        TMA through a 4/6-deep mbarrier ring, double-buffered accumulators in TMEM.
//   * gemm_ref_kernel -- a deliberately simple SIMT kernel used only by tests / the verification
//                        mode to isolate tensor-core/TMA bugs from epilogue / packing bugs.
//
// A-operand rows are addressed through "row modes" because the tree recursion stores node state in a
// slot-major array [257 slots][Bp candidates][features]: slot 0 = start frame, slot 256 = goal frame,
// slot s = in-order (depth-first) node s-1.  Level l, node j sits at slot (2j+1)*2^(7-l); its parents
// are the slots +-2^(7-l) away (gcp/prediction/utils/tree_utils.py:21-44,202-208 restated as indexing).
#pragma once
#include "common.cuh"

namespace gcp {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_MAX_SEGS = 5;
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quadrant, each takes half the columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

// ROW_CAND: row (node j, candidate c) of a level reads array row row_base + c -- a per-candidate array shared by all nodes
// (the start / goal encodings in slots 0 and n_nodes + 1 of the latent array)
enum RowMode { ROW_LEVEL = 0, ROW_SELF = 1, ROW_LEFT = 2, ROW_RIGHT = 3, ROW_CAND = 4 };
enum EpiKind { EPI_LINEAR = 0, EPI_GN = 1, EPI_REPARAM = 2, EPI_LSTM = 3 };
enum ActKind { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2 };

struct LevelGeom {
    int Bp;     // candidates padded to a multiple of 128
    int level;  // tree level 0..depth-1 (0 for plain GEMMs)
    int depth;  // 8
    // Work list of a pruned tree level (planner mode): tiles[i] = the 128-row tile (node j, candidate tile) = j * (Bp / 128) +
    // ctile that the i-th tile of this launch stands for; null = every tile in order.  Slot-addressed operands / outputs
    // (ROW_LEFT / ROW_RIGHT / ROW_SELF, the candidate index, the node of a row) use the LISTED tile, level-row scratch
    // arrays (ROW_LEVEL) are addressed compactly by the launch's own tile index, so a level's kernels only touch and
    // compute the (node, candidate tile) pairs some candidate keeps.
    const int* tiles;
};

struct ASeg {
    const bf16* ptr;   // base of the source array (for the SIMT reference path)
    int ld;            // leading dimension in elements
    int col0;          // first column of this K-segment in the source array
    int k_len;         // multiple of 64
    int row_mode;      // RowMode
    int row_base;      // ROW_LEVEL only: first array row of level-row 0 (views into bigger arrays)
    int group_cols;    // if > 0: output columns are split in groups of this many, each group reads a
    int group_col[16]; //          different column window of the source: col0 + group_col[n / group_cols]
};

struct EpiParams {
    const float* bias;      // [N] in packed column order (may be null)
    const float* rowbias;   // per-candidate additive term [Bp][rowbias_ld] (may be null)
    int rowbias_ld;
    const int* rowbias_idx; // non-null: the candidate of output row r is rowbias_idx[r] (compacted rows of the pruned
                            // decoder) instead of r % Bp
    const float* gn_gamma;  // EPI_GN
    const float* gn_beta;
    int gn_group;           // channels per group (16 or 4)
    int act;                // ActKind
    bf16* out_bf16;
    int out_bf16_ld, out_bf16_mode;
    float* out_f32;
    int out_f32_ld, out_f32_mode;
    int out_f32_df;         // >0 (= n_nodes): out_f32 is candidate-major depth-first [n_cand][n_nodes][ld]; a row of tree
                            // level g.level goes to (cand, node = slot - 1), padded candidates are dropped
    int split_col;          // >0: columns < split go to out_bf16, the rest to out_f32 (col - split)
    int n_valid;            // columns >= n_valid are dropped
    // EPI_REPARAM: zeta = exp(log_sigma) * eps + mu (blox/torch/dist.py:285-287)
    const float* z;         // [B][n_nodes][nz] candidate-major, depth-first nodes (reference layout)
    int n_cand;             // valid candidates B (<= Bp)
    int nz;                 // 256
    float* mu_out;          // optional [B][n_nodes][nz]
    float* ls_out;
    int z_node, z_nodes;    // z_nodes > 0: every row reads noise row z_node of z_nodes (sequential rollout, step t)
    // EPI_LSTM: torch.nn.LSTMCell update, gate order i,f,g,o
    const bf16* c_prev;     // projected cell state (bf16) [rows][c_prev_ld], column window c_prev_col0
    int c_prev_ld, c_prev_col0;
    bf16* hid;              // slot-major hidden state [slots*Bp][hid_ld]; h at hid_col0+u, c at +H
    int hid_ld, hid_col0, hidden;  // hidden = 512
    int write_hid;
    float* c_f32;           // non-null: fp32 cell state [rows][c_f32_ld], window c_prev_col0, updated in place
    int c_f32_ld;           //           (sequential rollout: 199 chained steps keep c in fp32)
};

struct GemmArgs {
    CUtensorMap a_map[GEMM_MAX_SEGS];
    CUtensorMap w_map;       // box = 64 columns x (BN / cluster size) rows
    ASeg seg[GEMM_MAX_SEGS];
    int n_seg;
    const bf16* w;   // packed weights [N][K] (K contiguous)
    int w_ld;
    int rows, N, K;  // rows % 128 == 0, N % BN == 0, K % 64 == 0
    // Device-side row count (pruned decoder: the number of kept node rows is known to the device only): if non-null, the
    // kernel processes min(rows, max(0, *rows_dev - rows_dev_base)) rows, rounded up to whole (cluster x 128)-row groups
    // (the caller zero-fills the padding rows of the A operand).
    const int* rows_dev;
    int rows_dev_base;
    LevelGeom g;
    EpiParams epi;
};

// ---------------------------------------------------------------------------------------------
// row addressing
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int slot_of(const LevelGeom& g, int j, int mode) {
    const int half = 1 << (g.depth - 1 - g.level);
    int slot = (2 * j + 1) * half;
    if (mode == ROW_LEFT) slot -= half;
    if (mode == ROW_RIGHT) slot += half;
    return slot;
}
// first array row of a 128-row tile (tiles never straddle nodes because Bp % 128 == 0)
__device__ __forceinline__ int tile_row0(const LevelGeom& g, int mode, int tile_m) {
    if (mode == ROW_LEVEL) return tile_m * GEMM_BM;   // caller adds row_base
    const int tpn = g.Bp >> 7;
    const int j = tile_m / tpn;
    const int c0 = (tile_m - j * tpn) << 7;
    if (mode == ROW_CAND) return c0;                   // caller adds row_base
    return slot_of(g, j, mode) * g.Bp + c0;
}
__device__ __forceinline__ int map_row(const LevelGeom& g, int mode, int row) {
    if (mode == ROW_LEVEL) return row;
    const int j = row / g.Bp;
    const int c = row - j * g.Bp;
    if (mode == ROW_CAND) return c;
    return slot_of(g, j, mode) * g.Bp + c;
}
__device__ __forceinline__ int listed_tile(const LevelGeom& g, int tile_m) {
    return g.tiles != nullptr ? __ldg(g.tiles + tile_m) : tile_m;
}
// array row of the launch's row `row` (compact) whose listed ("logical") row is `lrow`
__device__ __forceinline__ int map_row2(const LevelGeom& g, int mode, int row, int lrow) {
    return mode == ROW_LEVEL ? row : map_row(g, mode, lrow);
}
__device__ __forceinline__ int seg_col0(const ASeg& s, int n0) {
    return s.col0 + (s.group_cols > 0 ? s.group_col[n0 / s.group_cols] : 0);
}

// ---------------------------------------------------------------------------------------------
// epilogues: one thread = one output row, 32 consecutive (packed) columns at a time
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x32(bf16* dst, const float (&v)[32], int nvalid) {
    if (nvalid >= 32) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            d4[i] = u;
        }
    } else {
        // fully unrolled + predicated: a run-time index into v[] would move the accumulators to local memory
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) dst[i] = __float2bfloat16_rn(v[i]);
    }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32], int nvalid) {
    if (nvalid >= 32) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) dst[i] = v[i];
    }
}

// 32 rows x 32 bf16 columns of one warp, written as full 64-byte row segments: every thread parks its row piece in a
// per-warp 2 KB shared-memory tile (16-byte cells XOR-swizzled by row), then lanes 4k..4k+3 store the four cells of
// one row, 8 rows per instruction.  Storing straight from the accumulator registers (thread = row) issues four
// half-sector writes per row and caps the GEMMs with wide bf16 outputs at ~0.8 TB/s (profiles/r1_gemm_store_path.txt).
// `row0_ptr` = address of column col0 in the output row of lane 0; the 32 rows of a warp are consecutive.
__device__ __forceinline__ void store_bf16x32_staged(uint4* stage, bf16* row0_ptr, size_t ld, const float (&v)[32]) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
        u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
        u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        stage[lane * 4 + (i ^ ((lane >> 1) & 3))] = u;
    }
    __syncwarp();
    const int p = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        const uint4 u = stage[r * 4 + (p ^ ((r >> 1) & 3))];
        *reinterpret_cast<uint4*>(row0_ptr + (size_t)r * ld + p * 8) = u;
    }
}

template <int G>
__device__ __forceinline__ void group_norm_chunk(const EpiParams& p, int col0, float (&acc)[32]) {
#pragma unroll
    for (int g0 = 0; g0 < 32; g0 += G) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < G; ++i) s += acc[g0 + i];
        const float mean = s * (1.0f / (float)G);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const float d = acc[g0 + i] - mean;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(q * (1.0f / (float)G) + 1e-5f);
#pragma unroll
        for (int i = 0; i < G; ++i)
            acc[g0 + i] = (acc[g0 + i] - mean) * rstd * __ldg(p.gn_gamma + col0 + g0 + i) + __ldg(p.gn_beta + col0 + g0 + i);
    }
}

// GroupNorm over groups of 16 channels with the affine parameters staged in shared memory (gam / bet point at the
// chunk's first column; all lanes read the same address -> broadcast LDS.128) and pairwise sums: the tcgen05 kernel's
// row-MLP epilogue is latency-bound (8 warps per SM), so it is the dependent-chain length and the instruction count per
// element that matter (profiles/r1h_gemm_ncu_full.txt: 548 instructions per 32-column chunk before, 6.6 cycles each).
__device__ __forceinline__ void group_norm16_chunk_smem(const float* gam, const float* bet, float (&acc)[32]) {
#pragma unroll
    for (int g0 = 0; g0 < 32; g0 += 16) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = acc[g0 + 2 * i] + acc[g0 + 2 * i + 1];
        const float mean = (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) * (1.0f / 16.0f);
        float d[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = acc[g0 + i] - mean;
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = fmaf(d[2 * i + 1], d[2 * i + 1], d[2 * i] * d[2 * i]);
        const float var = (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]))) * (1.0f / 16.0f);
        const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 ga = *reinterpret_cast<const float4*>(gam + g0 + 4 * q);
            const float4 be = *reinterpret_cast<const float4*>(bet + g0 + 4 * q);
            acc[g0 + 4 * q + 0] = fmaf(d[4 * q + 0] * rstd, ga.x, be.x);
            acc[g0 + 4 * q + 1] = fmaf(d[4 * q + 1] * rstd, ga.y, be.y);
            acc[g0 + 4 * q + 2] = fmaf(d[4 * q + 2] * rstd, ga.z, be.z);
            acc[g0 + 4 * q + 3] = fmaf(d[4 * q + 3] * rstd, ga.w, be.w);
        }
    }
}

// Fast path of the 128-wide row-MLP layers (EPI_LINEAR / EPI_GN, all columns valid, bf16 output only, no row bias): the
// arithmetic of one chunk without any store, so that the kernel can run the two chunks of a thread as two independent
// instruction streams (the layer is bound by the latency of the epilogue, not by its instruction count).
template <int EPI>
__device__ __forceinline__ void epilogue_math_fast(const EpiParams& p, int col0, float (&acc)[32], const float* gn_sm) {
    if (p.bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
        }
    }
    if (EPI == EPI_GN) group_norm16_chunk_smem(gn_sm + col0, gn_sm + 256 + col0, acc);
    if (p.act == ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = lrelu_(acc[i]);
    } else if (p.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f);
    }
}

// torch.nn.LSTMCell update of 8 hidden units from their packed gate pre-activations [i(8) | f(8) | g(8) | o(8)].
// One MUFU op per gate (tanh.approx.f32; sigmoid(x) = 0.5 * tanh(x / 2) + 0.5) instead of ex2 + rcp: the cell update
// is MUFU-bound (5 instead of 10 ops per hidden unit), and a gate GEMM tile's epilogue is as long as its MMAs.
// tanh.approx: max relative error 2^-11, below the bf16 rounding (2^-9) of the h / c that leave this epilogue.
__device__ __forceinline__ void lstm_cell_update(const float (&acc)[32], const float (&cprev)[8], float (&h)[8], float (&c)[8],
                                                 uint4& hv, uint4& cv) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float ig = fmaf(0.5f, tanh_mufu(0.5f * acc[u]), 0.5f);
        const float fg = fmaf(0.5f, tanh_mufu(0.5f * acc[8 + u]), 0.5f);
        const float gg = tanh_mufu(acc[16 + u]);
        const float og = fmaf(0.5f, tanh_mufu(0.5f * acc[24 + u]), 0.5f);
        c[u] = fg * cprev[u] + ig * gg;
        h[u] = og * tanh_mufu(c[u]);
    }
    hv.x = pack_bf16x2(h[0], h[1]); hv.y = pack_bf16x2(h[2], h[3]);
    hv.z = pack_bf16x2(h[4], h[5]); hv.w = pack_bf16x2(h[6], h[7]);
    cv.x = pack_bf16x2(c[0], c[1]); cv.y = pack_bf16x2(c[2], c[3]);
    cv.z = pack_bf16x2(c[4], c[5]); cv.w = pack_bf16x2(c[6], c[7]);
}

// row: row of this launch (compact when the level is pruned); lrow: the row it stands for in the full level
// bias_sm: the chunk's 32 biases staged in shared memory, rb: the row's 32 row-bias values already in registers (both
// fetched by the caller before it waited for the accumulator); null: read them from global memory here.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& p, const LevelGeom& g, int row, int lrow, int col0,
                                               float (&acc)[32], uint4* stage = nullptr, const float* gn_sm = nullptr,
                                               const float* bias_sm = nullptr, const float4* rb = nullptr) {
    const int cand = lrow % g.Bp;
    if (p.bias != nullptr) {
        if (bias_sm != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_sm);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = b4[i];
                acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
            }
        } else {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = __ldg(b4 + i);
                acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
            }
        }
    }
    if (EPI == EPI_LINEAR || EPI == EPI_GN) {
        if (col0 >= p.n_valid) return;
        const int nvalid = min(32, p.n_valid - col0);
        if (p.rowbias != nullptr) {
            if (rb != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = rb[i];
                    acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
                }
            } else {
                const int rc = p.rowbias_idx != nullptr ? __ldg(p.rowbias_idx + row) : cand;
                const float4* rg = reinterpret_cast<const float4*>(p.rowbias + (size_t)rc * p.rowbias_ld + col0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = __ldg(rg + i);
                    acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
                }
            }
        }
        if (EPI == EPI_GN) {
            // torch GroupNorm: biased variance over the channels of one group, eps 1e-5
            if (p.gn_group == 16) {
                if (gn_sm != nullptr) group_norm16_chunk_smem(gn_sm + col0, gn_sm + 256 + col0, acc);
                else group_norm_chunk<16>(p, col0, acc);
            } else {
                group_norm_chunk<4>(p, col0, acc);
            }
        }
        if (p.act == ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = lrelu_(acc[i]);
        } else if (p.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f);
        }
        if (p.split_col > 0) {
            if (col0 < p.split_col) {
                const size_t r = (size_t)map_row2(g, p.out_bf16_mode, row, lrow);
                if (stage != nullptr && nvalid >= 32)
                    store_bf16x32_staged(stage, p.out_bf16 + (r - (threadIdx.x & 31)) * p.out_bf16_ld + col0, p.out_bf16_ld, acc);
                else
                    store_bf16x32(p.out_bf16 + r * p.out_bf16_ld + col0, acc, nvalid);
            } else {
                const size_t r = (size_t)map_row2(g, p.out_f32_mode, row, lrow);
                store_f32x32(p.out_f32 + r * p.out_f32_ld + (col0 - p.split_col), acc, nvalid);
            }
        } else {
            if (p.out_bf16 != nullptr) {
                const size_t r = (size_t)map_row2(g, p.out_bf16_mode, row, lrow);
                if (stage != nullptr && nvalid >= 32)
                    store_bf16x32_staged(stage, p.out_bf16 + (r - (threadIdx.x & 31)) * p.out_bf16_ld + col0, p.out_bf16_ld, acc);
                else
                    store_bf16x32(p.out_bf16 + r * p.out_bf16_ld + col0, acc, nvalid);
            }
            if (p.out_f32 != nullptr) {
                if (p.out_f32_df > 0) {
                    if (cand < p.n_cand) {
                        const size_t r = (size_t)cand * p.out_f32_df + slot_of(g, lrow / g.Bp, ROW_SELF) - 1;
                        store_f32x32(p.out_f32 + r * p.out_f32_ld + col0, acc, nvalid);
                    }
                } else {
                    const size_t r = (size_t)map_row2(g, p.out_f32_mode, row, lrow);
                    store_f32x32(p.out_f32 + r * p.out_f32_ld + col0, acc, nvalid);
                }
            }
        }
    } else if (EPI == EPI_REPARAM) {
        // packed columns: [mu(16) | log_sigma(16)] for latent dims d0 .. d0+15
        const int d0 = col0 >> 1;
        const int j = lrow / g.Bp;
        const int node = p.z_nodes > 0 ? p.z_node : slot_of(g, j, ROW_SELF) - 1;  // depth-first node index / time step
        const int n_nodes = p.z_nodes > 0 ? p.z_nodes : (1 << g.depth) - 1;
        float zeta[16];
        if (cand < p.n_cand) {
            const size_t zoff = ((size_t)cand * n_nodes + node) * p.nz + d0;
            const float4* z4 = reinterpret_cast<const float4*>(p.z + zoff);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 e = __ldg(z4 + i);
                zeta[4 * i + 0] = __expf(acc[16 + 4 * i + 0]) * e.x + acc[4 * i + 0];
                zeta[4 * i + 1] = __expf(acc[16 + 4 * i + 1]) * e.y + acc[4 * i + 1];
                zeta[4 * i + 2] = __expf(acc[16 + 4 * i + 2]) * e.z + acc[4 * i + 2];
                zeta[4 * i + 3] = __expf(acc[16 + 4 * i + 3]) * e.w + acc[4 * i + 3];
            }
            if (p.mu_out != nullptr) {
                for (int i = 0; i < 16; ++i) {
                    p.mu_out[zoff + i] = acc[i];
                    p.ls_out[zoff + i] = acc[16 + i];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) zeta[i] = acc[i];  // padded candidates: eps = 0
        }
        uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.out_bf16_ld + d0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint4 u;
            u.x = pack_bf16x2(zeta[8 * i + 0], zeta[8 * i + 1]);
            u.y = pack_bf16x2(zeta[8 * i + 2], zeta[8 * i + 3]);
            u.z = pack_bf16x2(zeta[8 * i + 4], zeta[8 * i + 5]);
            u.w = pack_bf16x2(zeta[8 * i + 6], zeta[8 * i + 7]);
            o[i] = u;
        }
    } else if (EPI == EPI_LSTM) {
        // packed columns: [i(8) | f(8) | g(8) | o(8)] for hidden units u0 .. u0+7
        const int u0 = col0 >> 2;
        float cprev[8];
        float4* cf = nullptr;
        if (p.c_f32 != nullptr) {
            cf = reinterpret_cast<float4*>(p.c_f32 + (size_t)row * p.c_f32_ld + p.c_prev_col0 + u0);
            const float4 a = cf[0], b = cf[1];
            cprev[0] = a.x; cprev[1] = a.y; cprev[2] = a.z; cprev[3] = a.w;
            cprev[4] = b.x; cprev[5] = b.y; cprev[6] = b.z; cprev[7] = b.w;
        } else {
            const uint4 cp = __ldg(reinterpret_cast<const uint4*>(p.c_prev + (size_t)row * p.c_prev_ld + p.c_prev_col0 + u0));
            const __nv_bfloat162* cp2 = reinterpret_cast<const __nv_bfloat162*>(&cp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = __bfloat1622float2(cp2[i]);
                cprev[2 * i] = t.x;
                cprev[2 * i + 1] = t.y;
            }
        }
        float h[8], c[8];
        uint4 hv, cv;
        lstm_cell_update(acc, cprev, h, c, hv, cv);
        *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.out_bf16_ld + u0) = hv;
        if (cf != nullptr) {
            cf[0] = make_float4(c[0], c[1], c[2], c[3]);
            cf[1] = make_float4(c[4], c[5], c[6], c[7]);
        }
        if (p.write_hid) {
            const size_t r = (size_t)map_row(g, ROW_SELF, lrow);
            bf16* hrow = p.hid + r * p.hid_ld + p.hid_col0 + u0;
            *reinterpret_cast<uint4*>(hrow) = hv;
            *reinterpret_cast<uint4*>(hrow + p.hidden) = cv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------
// PAIR: the CTAs of a 2-CTA cluster run one M = 256 MMA (tcgen05 cta_group::2, see common.cuh): a CTA stages its own 128
// rows of A and half of the B tile, so a stage is 32 KB instead of 48 KB at BN = 256 and six of them fit.
template <int BN, bool PAIR = false>
struct GemmCfg {
    static constexpr int STAGES = (BN == 256 && !PAIR) ? 4 : 6;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STORE_STAGE_BYTES = GEMM_EPI_WARPS * 2048;   // per-warp transpose tile of the bf16 store path
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STORE_STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, int EPI, bool PAIR = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmArgs args) {
    using Cfg = GemmCfg<BN, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tmem_full = empty_bar + Cfg::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint4* store_stage = reinterpret_cast<uint4*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);

    // per epilogue warp: the biases of the columns it drains (up to 128), staged before the accumulator wait
    __shared__ __align__(16) float bias_stage[GEMM_EPI_WARPS * 128];
    // EPI_GN: GroupNorm scale / shift of all N <= 256 columns (weights, so reading them before pdl_wait() is safe)
    __shared__ __align__(16) float gn_sm[EPI == EPI_GN ? 512 : 4];
    if (EPI == EPI_GN && args.N <= 256) {
        for (int i = threadIdx.x; i < args.N; i += GEMM_THREADS) {
            gn_sm[i] = __ldg(args.epi.gn_gamma + i);
            gn_sm[256 + i] = __ldg(args.epi.gn_beta + i);
        }
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_n = args.N / BN;
    const int num_kb = args.K / GEMM_BK;
    // Thread-block clusters of C CTAs along M share the weight tile: every CTA fetches 1/C of it and TMA
    // multicasts the slice into all C shared memories, cutting L2->SM operand traffic (the bound of this
    // kernel at BN=256: 48 KB per k-block and CTA without multicast, 16 + 32/C KB with it).
    const uint32_t C = cluster_nctarank();
    const uint32_t crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);
    const int work0 = blockIdx.x / (int)C, work_stride = gridDim.x / (int)C;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < args.n_seg; ++s) tma_prefetch_desc(&args.a_map[s]);
        tma_prefetch_desc(&args.w_map);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR ? 1 : C);   // released by the UMMA issuer of every CTA in the cluster (PAIR: the leader's)
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            // PAIR: one arrival per epilogue warp of BOTH CTAs, on the leader's barrier
            mbar_init(&tmem_empty[s], PAIR ? 2 * GEMM_EPI_WARPS : 32 * GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair(tmem_holder, Cfg::TMEM_COLS);
        else tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();             // peers' barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch, cluster rendezvous) overlaps the tail of
    // the previous kernel in the stream; nothing below may run before that kernel's writes are visible.
    pdl_launch_dependents();
    pdl_wait();
    int rows = args.rows;
    if (args.rows_dev != nullptr) {           // produced by an earlier kernel of the stream: read only after pdl_wait()
        const int grp = GEMM_BM * (int)C;
        const int left = max(__ldg(args.rows_dev) - args.rows_dev_base, 0);
        rows = min(rows, (left + grp - 1) / grp * grp);
    }
    const int tiles_m = rows / GEMM_BM;
    const int n_work = (tiles_m / (int)C) * tiles_n;          // work item = (tile_n, group of C consecutive tile_m)

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            const int slice_rows = BN / (int)C;
            for (int work = work0; work < n_work; work += work_stride) {
                const int grp = work / tiles_n, tile_n = work - grp * tiles_n;
                const int tile_m = grp * (int)C + (int)crank;
                int kb = 0;
                for (int s = 0; s < args.n_seg; ++s) {
                    const ASeg& sg = args.seg[s];
                    const int row0 = tile_row0(args.g, sg.row_mode, sg.row_mode == ROW_LEVEL ? tile_m : listed_tile(args.g, tile_m)) + sg.row_base;
                    const int c0 = seg_col0(sg, tile_n * BN);
                    for (int kk = 0; kk < sg.k_len; kk += GEMM_BK, ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                        if (PAIR) {
                            // both CTAs' bytes are counted on the LEADER's barrier: it expects two stages' worth
                            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                            const uint32_t lead_bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
                            tma_load_2d_pair(sa, &args.a_map[s], lead_bar, c0 + kk, row0);
                            tma_load_2d_pair(sa + Cfg::A_BYTES, &args.w_map, lead_bar, kb * GEMM_BK, tile_n * BN + (int)crank * (BN / 2));
                            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(sa, &args.a_map[s], &full_bar[stage], c0 + kk, row0);
                        if (C == 1)
                            tma_load_2d(sa + Cfg::A_BYTES, &args.w_map, &full_bar[stage], kb * GEMM_BK, tile_n * BN);
                        else
                            tma_load_2d_mcast(sa + Cfg::A_BYTES + crank * slice_rows * 128, &args.w_map, &full_bar[stage],
                                              kb * GEMM_BK, tile_n * BN + (int)crank * slice_rows, cmask);
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && (!PAIR || crank == 0)) {
            // ===== UMMA issuer (single thread; PAIR: of the leader CTA, for both) =====
            constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int work = work0; work < n_work; work += work_stride) {
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(sa);
                    const uint64_t db = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        if (PAIR) umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    if (PAIR) umma_commit_pair(&empty_bar[stage], 3);
                    else if (C == 1) umma_commit(&empty_bar[stage]);
                    else umma_commit_mcast(&empty_bar[stage], cmask);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (PAIR) umma_commit_pair(&tmem_full[as], 3);
                else umma_commit(&tmem_full[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // which half of the tile's columns this warp drains
        constexpr int CH_PER_WARP = BN / 32 / (GEMM_EPI_WARPS / 4);
        constexpr bool kPair = (BN == 128) && (EPI == EPI_LINEAR || EPI == EPI_GN);
        const bool pair_fast = kPair && args.epi.rowbias == nullptr && args.epi.split_col == 0 && args.epi.out_f32 == nullptr &&
                               args.epi.out_bf16 != nullptr && args.epi.n_valid >= args.N &&
                               (EPI != EPI_GN || (args.epi.gn_group == 16 && args.N <= 256));
        // LSTM gate GEMMs with bf16 previous cell state (what the rollouts run): drain with prefetched operands, see below
        constexpr bool kLstmFast = (EPI == EPI_LSTM) && (CH_PER_WARP * 32 == 128);
        const bool lstm_fast = kLstmFast && args.epi.c_f32 == nullptr && args.epi.bias != nullptr;
        float* lstm_bias_sm = bias_stage + (warp - 2) * 128;
        float* const bias_sm_w = lstm_bias_sm;
        const int row_in_tile = q * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        for (int work = work0; work < n_work; work += work_stride) {
            const int grp = work / tiles_n, tile_n = work - grp * tiles_n;
            const int tile_m = grp * (int)C + (int)crank;
            const int row = tile_m * GEMM_BM + row_in_tile;
            const int lrow = listed_tile(args.g, tile_m) * GEMM_BM + row_in_tile;
            // LSTM epilogue: everything a tile's cell update needs besides the accumulator -- this warp's 128 gate biases
            // (staged in its 2 KB of shared memory: every row adds the same ones), the row's previous cell state, its
            // slot row in the state array (an integer division) -- is fetched BEFORE waiting for the accumulator, so the
            // global-load latency overlaps the tile's MMAs instead of sitting in the drain (measured: the bias and c_prev
            // loads were half of the drain's stall samples, profiles/r2aj_lstm_epilogue_ncu.txt)
            uint4 lstm_cp[kLstmFast ? CH_PER_WARP : 1];
            bf16 *lstm_out = nullptr, *lstm_hid = nullptr;
            if (kLstmFast && lstm_fast) {
                const int colw = tile_n * BN + half * CH_PER_WARP * 32;      // this warp's first packed column
                const float4 bl = __ldg(reinterpret_cast<const float4*>(args.epi.bias + colw) + lane);
                const bf16* cpr = args.epi.c_prev + (size_t)row * args.epi.c_prev_ld + args.epi.c_prev_col0 + (colw >> 2);
#pragma unroll
                for (int k = 0; k < CH_PER_WARP; ++k) lstm_cp[k] = __ldg(reinterpret_cast<const uint4*>(cpr + k * 8));
                lstm_out = args.epi.out_bf16 + (size_t)row * args.epi.out_bf16_ld + (colw >> 2);
                if (args.epi.write_hid)
                    lstm_hid = args.epi.hid + (size_t)map_row(args.g, ROW_SELF, lrow) * args.epi.hid_ld + args.epi.hid_col0 + (colw >> 2);
                __syncwarp();                                   // the previous tile's reads of the staged biases are done
                reinterpret_cast<float4*>(lstm_bias_sm)[lane] = bl;
                __syncwarp();
            }
            // generic drain: the same idea -- this warp's biases staged in shared memory and the row bias of its first
            // chunk in registers before the wait; the row bias of chunk k + 1 is requested before chunk k is processed
            const bool gen_pref = !(kLstmFast && lstm_fast) && !(kPair && pair_fast) && EPI != EPI_LSTM;
            const bool rb_pref = gen_pref && BN == 256 && (EPI == EPI_LINEAR || EPI == EPI_GN) && args.epi.rowbias != nullptr;
            float4 rbn[8];
            const float* rb_row = nullptr;
            if (gen_pref) {
                constexpr int WCOLS = CH_PER_WARP * 32;
                const int colw = tile_n * BN + half * WCOLS;
                if (args.epi.bias != nullptr) {
                    float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane < WCOLS / 4) bl = __ldg(reinterpret_cast<const float4*>(args.epi.bias + colw) + lane);
                    __syncwarp();
                    if (lane < WCOLS / 4) reinterpret_cast<float4*>(bias_sm_w)[lane] = bl;
                    __syncwarp();
                }
                if (rb_pref) {
                    const int rc = args.epi.rowbias_idx != nullptr ? __ldg(args.epi.rowbias_idx + row) : lrow % args.g.Bp;
                    rb_row = args.epi.rowbias + (size_t)rc * args.epi.rowbias_ld;
                    if (colw < args.epi.n_valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rbn[i] = __ldg(reinterpret_cast<const float4*>(rb_row + colw) + i);
                    }
                }
            }
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
            if (kLstmFast && lstm_fast) {
                // (operands fetched above, before the wait) drain: bias from shared memory, cell update, stores
#pragma unroll
                for (int k = 0; k < CH_PER_WARP; ++k) {
                    float acc[32];
                    __syncwarp();
                    tmem_ld32(t0 + (half * CH_PER_WARP + k) * 32, acc);
                    if (k == CH_PER_WARP - 1) {      // the accumulator buffer goes back as soon as it is in registers
                        tc_fence_before();
                        if (PAIR) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                        } else {
                            mbar_arrive_relaxed(&tmem_empty[as]);
                        }
                    }
                    const float4* b4 = reinterpret_cast<const float4*>(lstm_bias_sm) + k * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        acc[4 * i] += b.x; acc[4 * i + 1] += b.y; acc[4 * i + 2] += b.z; acc[4 * i + 3] += b.w;
                    }
                    float cprev[8], h[8], c[8];
                    const __nv_bfloat162* cp2 = reinterpret_cast<const __nv_bfloat162*>(&lstm_cp[k]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 t = __bfloat1622float2(cp2[i]);
                        cprev[2 * i] = t.x;
                        cprev[2 * i + 1] = t.y;
                    }
                    uint4 hv, cv;
                    lstm_cell_update(acc, cprev, h, c, hv, cv);
                    *reinterpret_cast<uint4*>(lstm_out + k * 8) = hv;
                    if (args.epi.write_hid) {
                        *reinterpret_cast<uint4*>(lstm_hid + k * 8) = hv;
                        *reinterpret_cast<uint4*>(lstm_hid + k * 8 + args.epi.hidden) = cv;
                    }
                }
                if (++as == 2) { as = 0; aphase ^= 1; }
                continue;
            }
            if (kPair && pair_fast) {
                // both chunks of this thread at once; the accumulator buffer is released as soon as they are in registers
                float a0[32], a1[32];
                const int ch0 = half * 2;
                __syncwarp();
                tmem_ld32_pair(t0 + ch0 * 32, t0 + (ch0 + 1) * 32, a0, a1);
                tc_fence_before();
                if (PAIR) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[as]), 0));
                } else {
                    mbar_arrive_relaxed(&tmem_empty[as]);
                }
                const int c0 = tile_n * BN + ch0 * 32;
                epilogue_math_fast<EPI>(args.epi, c0, a0, gn_sm);
                epilogue_math_fast<EPI>(args.epi, c0 + 32, a1, gn_sm);
                const size_t r = (size_t)map_row2(args.g, args.epi.out_bf16_mode, row, lrow);
                bf16* orow = args.epi.out_bf16 + (r - lane) * args.epi.out_bf16_ld + c0;
                uint4* stg = store_stage + (warp - 2) * 128;
                store_bf16x32_staged(stg, orow, args.epi.out_bf16_ld, a0);
                store_bf16x32_staged(stg, orow + 32, args.epi.out_bf16_ld, a1);
                if (++as == 2) { as = 0; aphase ^= 1; }
                continue;
            }
#pragma unroll 1
            for (int ch = half * CH_PER_WARP; ch < (half + 1) * CH_PER_WARP; ++ch) {
                float acc[32];
                float4 rbc[8];
                const int col0 = tile_n * BN + ch * 32;
                if (rb_pref) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rbc[i] = rbn[i];
                    if (ch + 1 < (half + 1) * CH_PER_WARP && col0 + 32 < args.epi.n_valid) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rbn[i] = __ldg(reinterpret_cast<const float4*>(rb_row + col0 + 32) + i);
                    }
                }
                __syncwarp();
                tmem_ld32(t0 + ch * 32, acc);
                epilogue_chunk<EPI>(args.epi, args.g, row, lrow, col0, acc, store_stage + (warp - 2) * 128,
                                    (EPI == EPI_GN && args.N <= 256) ? gn_sm : nullptr,
                                    (gen_pref && args.epi.bias != nullptr) ? bias_sm_w + (ch - half * CH_PER_WARP) * 32 : nullptr,
                                    rb_pref ? rbc : nullptr);
            }
            tc_fence_before();
            if (PAIR) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[as]), 0));
            } else {
                mbar_arrive_relaxed(&tmem_empty[as]);
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();             // no CTA exits while peers may still multicast into it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT verification kernel (same args, same epilogue); GCPB200_VERIFY builds only
// ---------------------------------------------------------------------------------------------
#ifdef GCPB200_VERIFY
template <int EPI>
__global__ void __launch_bounds__(128) gemm_ref_kernel(const __grid_constant__ GemmArgs args, int BN) {
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int row = tile_m * GEMM_BM + threadIdx.x;
    for (int ch = 0; ch < BN / 32; ++ch) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        const int n0 = tile_n * BN + ch * 32;
        int kb = 0;
        for (int s = 0; s < args.n_seg; ++s) {
            const ASeg& sg = args.seg[s];
            const bf16* arow = sg.ptr + ((size_t)map_row(args.g, sg.row_mode, row) + sg.row_base) * sg.ld + seg_col0(sg, tile_n * BN);
            for (int k = 0; k < sg.k_len; ++k) {
                const float a = __bfloat162float(arow[k]);
                const bf16* wp = args.w + (size_t)n0 * args.w_ld + kb + k;
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = fmaf(a, __bfloat162float(wp[(size_t)i * args.w_ld]), acc[i]);
            }
            kb += sg.k_len;
        }
        epilogue_chunk<EPI>(args.epi, args.g, row, row, n0, acc);
    }
}
#endif

}  // namespace gcp
