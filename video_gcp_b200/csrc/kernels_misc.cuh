// Small kernels around the GEMM / decoder cores: image encoder, balanced pruning, gathers, trajectory
// costs, elite selection, refit, on-device noise.  These are HBM- or latency-bound; they use coalesced,
// vectorised global access and warp-shuffle reductions.
#pragma once
#include "common.cuh"

namespace gcp {

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Conv encoder (blox/torch/encoder_decoder.py:31-53): k4 s2 p1 x3 (+BN eval, LeakyReLU .2) then k4
// valid head.  One block per image; activations stay in shared memory.  BatchNorm is folded to a
// per-channel scale/shift on the host.  1.4 MMAC per image, < 0.1 % of the rollout: SIMT is enough.
// ---------------------------------------------------------------------------------------------
struct EncoderWeights {
    const float* w0;   // [16][3][4][4]
    const float* b0;   // [16]
    const float* w1;   // [32][16][4][4]
    const float* sc1;  // [32] BN scale
    const float* sh1;  // [32] BN shift
    const float* w2;   // [64][32][4][4]
    const float* sc2;
    const float* sh2;
    const float* w3;   // [128][64][4][4]
    const float* b3;   // [128]
};

// One strided convolution layer (k4 s2 p1) as work items (output element, part): `P` consecutive lanes share one output, each
// sums CI / P input channels in (ci, ky, kx) order, an xor-shuffle tree adds the parts, part 0 applies bias / folded BN /
// LeakyReLU and hands the value to `store(o, v)`.  The arithmetic of an output element depends on nothing but the layer and
// P, so any distribution of the items over threads / CTAs gives the same bits (the per-image kernel and the cluster kernel
// below share this function).  [item0, item1) must be a multiple of 32 long and `nthreads` a multiple of 32.
template <int CI, int CO, int HIN, int P, class Store>
__device__ __forceinline__ void enc_conv_items(const float* in, const float* w, const float* sc, const float* sh,
                                               const float* bias, int item0, int item1, int tid, int nthreads, Store store) {
    constexpr int HO = HIN / 2, CPP = CI / P;
    for (int it = item0 + tid; it < item1; it += nthreads) {
        const int o = it / P, part = it % P;
        const int co = o / (HO * HO), oy = (o / HO) % HO, ox = o % HO;
        float s = 0.f;
        for (int ci = part * CPP; ci < (part + 1) * CPP; ++ci) {
            const float* wp = w + ((size_t)co * CI + ci) * 16;
            const float* ip = in + ci * HIN * HIN;
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const int iy = 2 * oy - 1 + ky;
                if (iy < 0 || iy >= HIN) continue;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const int ix = 2 * ox - 1 + kx;
                    if (ix < 0 || ix >= HIN) continue;
                    s = fmaf(ip[iy * HIN + ix], __ldg(wp + ky * 4 + kx), s);
                }
            }
        }
#pragma unroll
        for (int d = P / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (part == 0) {
            if (bias != nullptr) s += __ldg(bias + co);
            if (sc != nullptr) s = s * __ldg(sc + co) + __ldg(sh + co);
            store(o, lrelu_(s));
        }
    }
}
// head: conv k4 valid on the 4x4 map = 1024 -> 128 linear; 8 lanes per output (128 inputs each), fixed summation order
template <class Store>
__device__ __forceinline__ void enc_head_items(const float* a3, const float* w3, const float* b3, int item0, int item1, int tid,
                                               int nthreads, Store store) {
    for (int it = item0 + tid; it < item1; it += nthreads) {
        const int o = it >> 3, part = it & 7;
        const float* wp = w3 + (size_t)o * 1024 + part * 128;
        const float* ap = a3 + part * 128;
        float s = 0.f;
        for (int k = 0; k < 128; ++k) s = fmaf(ap[k], __ldg(wp + k), s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (part == 0) store(o, s + __ldg(b3 + o));
    }
}

// grid = (n_images, 2): blockIdx.y = 0 encodes the start images (latent rows row0_a + i, skips stored), 1 the goal
// images (latent rows row0_b + i).  1024 threads per image; activations stay in shared memory.  Skip outputs are the
// post-activation layer-0 / layer-2 maps (GetIntermediatesSequential, stride 2).
constexpr int ENC_THREADS = 1024;
__global__ void __launch_bounds__(ENC_THREADS) encoder_kernel(const float* __restrict__ img_a, const float* __restrict__ img_b,
                                                              EncoderWeights W, float* lat_f32, bf16* lat_bf16, int row0_a,
                                                              int row0_b, float* skip0, float* skip2, bf16* skip2_bf16) {
    __shared__ float a0[3 * 32 * 32];
    __shared__ float a1[16 * 16 * 16];
    __shared__ float a2[32 * 8 * 8];
    __shared__ float a3[64 * 4 * 4];
    const int tid = threadIdx.x, i = blockIdx.x;
    const bool goal = blockIdx.y != 0;
    const float* img = goal ? img_b : img_a;
    for (int k = tid; k < 3072; k += ENC_THREADS) a0[k] = img[(size_t)i * 3072 + k];
    __syncthreads();
    enc_conv_items<3, 16, 32, 1>(a0, W.w0, nullptr, nullptr, W.b0, 0, 4096, tid, ENC_THREADS, [&](int o, float v) { a1[o] = v; });
    __syncthreads();
    enc_conv_items<16, 32, 16, 4>(a1, W.w1, W.sc1, W.sh1, nullptr, 0, 4 * 2048, tid, ENC_THREADS, [&](int o, float v) { a2[o] = v; });
    __syncthreads();
    enc_conv_items<32, 64, 8, 4>(a2, W.w2, W.sc2, W.sh2, nullptr, 0, 4 * 1024, tid, ENC_THREADS, [&](int o, float v) { a3[o] = v; });
    __syncthreads();
    if (!goal && skip0 != nullptr) {
        for (int k = tid; k < 4096; k += ENC_THREADS) skip0[(size_t)i * 4096 + k] = a1[k];
        for (int k = tid; k < 1024; k += ENC_THREADS) {
            skip2[(size_t)i * 1024 + k] = a3[k];
            skip2_bf16[(size_t)i * 1024 + k] = __float2bfloat16_rn(a3[k]);
        }
    }
    const size_t r = (size_t)(goal ? row0_b : row0_a) + i;
    enc_head_items(a3, W.w3, W.b3, 0, 1024, tid, ENC_THREADS, [&](int o, float v) {
        lat_f32[r * 128 + o] = v;
        lat_bf16[r * 128 + o] = __float2bfloat16_rn(v);
    });
}

// The same encoder for ONE start / goal pair (a CEM call: every candidate shares them), where the per-image kernel above is
// two CTAs and 170 us of pure latency at the head of every rollout: a cluster of 8 CTAs per image takes an eighth of every
// layer's work items and broadcasts its outputs into the activation arrays of all 8 CTAs through distributed shared memory
// (st.shared::cluster), one cluster barrier per layer.  Same work items, same bits.  grid = 16 (2 clusters of 8).
constexpr int ENCC_CTAS = 8, ENCC_THREADS = 512;
__device__ __forceinline__ void st_cluster_all(float* local, float v) {
    const uint32_t a = smem_u32(local);
#pragma unroll
    for (uint32_t r = 0; r < ENCC_CTAS; ++r) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(mapa_shared(a, r)), "f"(v) : "memory");
}
__global__ void __launch_bounds__(ENCC_THREADS) encoder_cluster_kernel(const float* __restrict__ img_a, const float* __restrict__ img_b,
                                                                       EncoderWeights W, float* lat_f32, bf16* lat_bf16, int row0_a,
                                                                       int row0_b, float* skip0, float* skip2, bf16* skip2_bf16) {
    __shared__ float a0[3 * 32 * 32];
    __shared__ float a1[16 * 16 * 16];
    __shared__ float a2[32 * 8 * 8];
    __shared__ float a3[64 * 4 * 4];
    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const bool goal = blockIdx.x >= ENCC_CTAS;
    const float* img = goal ? img_b : img_a;
    for (int k = tid; k < 3072; k += ENCC_THREADS) a0[k] = img[k];
    __syncthreads();
    cluster_sync_all();          // every CTA of the cluster is running before anyone writes into its shared memory
    const bool keep = !goal && skip0 != nullptr;
    enc_conv_items<3, 16, 32, 1>(a0, W.w0, nullptr, nullptr, W.b0, rank * 512, (rank + 1) * 512, tid, ENCC_THREADS, [&](int o, float v) {
        st_cluster_all(&a1[o], v);
        if (keep) skip0[o] = v;
    });
    cluster_sync_all();
    enc_conv_items<16, 32, 16, 4>(a1, W.w1, W.sc1, W.sh1, nullptr, rank * 1024, (rank + 1) * 1024, tid, ENCC_THREADS,
                                  [&](int o, float v) { st_cluster_all(&a2[o], v); });
    cluster_sync_all();
    enc_conv_items<32, 64, 8, 4>(a2, W.w2, W.sc2, W.sh2, nullptr, rank * 512, (rank + 1) * 512, tid, ENCC_THREADS, [&](int o, float v) {
        st_cluster_all(&a3[o], v);
        if (keep) {
            skip2[o] = v;
            skip2_bf16[o] = __float2bfloat16_rn(v);
        }
    });
    cluster_sync_all();
    const size_t r = (size_t)(goal ? row0_b : row0_a);
    enc_head_items(a3, W.w3, W.b3, rank * 128, (rank + 1) * 128, tid, ENCC_THREADS, [&](int o, float v) {
        lat_f32[r * 128 + o] = v;
        lat_bf16[r * 128 + o] = __float2bfloat16_rn(v);
    });
    cluster_sync_all();          // no CTA exits while a peer may still store into it
}

// broadcast row 0 of a slot to all candidates (CEM: every candidate shares the start / goal image)
__global__ void broadcast_rows_kernel(float* f32, bf16* b16, int row0, int n_rows, int cols) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (n_rows - 1) * cols) return;
    const int r = 1 + idx / cols, c = idx % cols;
    if (f32 != nullptr) f32[((size_t)row0 + r) * cols + c] = f32[(size_t)row0 * cols + c];
    if (b16 != nullptr) b16[((size_t)row0 + r) * cols + c] = b16[(size_t)row0 * cols + c];
}

// ---------------------------------------------------------------------------------------------
// Rollout length: OneHotCategorical(logits).sample() -> argmax -> clamp(min=2)
// (gcp/prediction/models/base_gcp.py:215-229).  Gumbel-max with a counter-based hash RNG.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void sample_length_kernel(const float* __restrict__ logits, int ld, int n_len, int n_cand,
                                     unsigned long long seed, long long* end_ind) {
    // one warp per candidate: lanes stride over the lengths, then a warp arg-max (first maximum wins, as np.argmax)
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    float best = -INFINITY;
    int arg = 0;
    for (int k = lane; k < n_len; k += 32) {
        const uint32_t h = mix32(mix32((uint32_t)seed ^ (uint32_t)(seed >> 32)) + 0x9E3779B9u * (uint32_t)(c * n_len + k + 1));
        const float u = ((h >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float v = logits[(size_t)c * ld + k] - __logf(-__logf(u));
        if (v > best) { best = v; arg = k; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) end_ind[c] = arg < 2 ? 2 : arg;
}

// Sampled rollout lengths of one CEM call, reassigned in descending order (counting sort of values in [0, 256); one CTA).
// With shared start / goal images every candidate's length is an i.i.d. draw from the SAME distribution, independent of its
// noise, so handing the sorted draws to candidates 0, 1, 2, ... leaves the joint law of the (noise, length) pairs -- and
// with it costs, elites and refit -- unchanged, while 128-candidate tiles become homogeneous in length: that is what makes
// the per-level work lists of the pruned tree recursion short (tree_worklists_kernel).
__global__ void __launch_bounds__(1024) sort_lengths_desc_kernel(long long* __restrict__ end_ind, int n) {
    __shared__ int hist[256];
    __shared__ int start[257];          // start[v] = number of values > v  (first position of value v in descending order)
    const int tid = threadIdx.x;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) atomicAdd(&hist[min(max((int)end_ind[i], 0), 255)], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int v = 255; v >= 0; --v) { start[v] = acc; acc += hist[v]; }
        start[256] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        int v = 255;
        while (v > 0 && start[v] + hist[v] <= i) --v;      // the value whose run [start, start + count) holds position i
        end_ind[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Balanced pruning (gcp/evaluation/evaluation_matching.py:192-206; gcp/prediction/models/tree/
// frame_binding.py:42-65): integer interval recursion from (-1, end_ind+1); midpoint with truncating
// division; a node is kept iff its timestep differs from both interval ends.  Kept nodes have the
// distinct timesteps 0..end_ind, so frame t of candidate c is node frame_node[c][t].
// ---------------------------------------------------------------------------------------------
__global__ void prune_map_kernel(const long long* __restrict__ end_ind, int n_cand, int depth, int lcap,
                                 int* __restrict__ frame_node) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_nodes = (1 << depth) - 1;
    if (idx >= n_cand * n_nodes) return;
    const int c = idx / n_nodes, node = idx - c * n_nodes;
    // node (in-order index) -> level / position: slot = node+1 = (2j+1) * 2^(depth-1-level)
    const int slot = node + 1;
    const int tz = __ffs(slot) - 1;
    const int level = depth - 1 - tz;
    const int j = slot >> (tz + 1);
    int l = -1, r = (int)end_ind[c] + 1, t = 0;
    for (int lv = 0; lv <= level; ++lv) {
        t = (l + r) / 2;  // C integer division truncates toward zero, as torch 1.3 did on int64
        if (lv < level) {
            if ((j >> (level - 1 - lv)) & 1) l = t; else r = t;
        }
    }
    if (t != l && t != r && t >= 0 && t < lcap) frame_node[c * lcap + t] = node;
}

// dst[c][t][:] = src[c][frame_node[c][t]][:] for t <= end_ind[c], zeros after (pad_sequence).  dst_b16 (optional): the
// same sequence as bf16 rows [c][lcap + 1][4 * d4] -- one extra, always-zero row per candidate, so that the GEMM A operand
// "row r | row r + 1" is the consecutive-frame pair of the inverse model without crossing into the next candidate.
__global__ void gather_frames_kernel(const float* __restrict__ src, const int* __restrict__ frame_node,
                                     const long long* __restrict__ end_ind, int n_cand, int n_nodes, int lcap,
                                     int d4 /* row length in float4 */, float* __restrict__ dst,
                                     bf16* __restrict__ dst_b16 = nullptr) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n_cand * lcap * d4;
    if (idx >= total) return;
    const int k = idx % d4;
    const int t = (idx / d4) % lcap;
    const int c = idx / ((size_t)d4 * lcap);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t <= (int)end_ind[c]) {
        const int node = frame_node[c * lcap + t];
        v = __ldg(reinterpret_cast<const float4*>(src) + ((size_t)c * n_nodes + node) * d4 + k);
    }
    reinterpret_cast<float4*>(dst)[idx] = v;
    if (dst_b16 != nullptr) {
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(dst_b16)[((size_t)c * (lcap + 1) + t) * d4 + k] = u;
    }
}

// fp32 sequences [c][lcap][4 * d4] -> bf16 rows [c][lcap + 1][4 * d4] (the layout gather_frames_kernel's dst_b16 has)
__global__ void seq_to_b16_rows_kernel(const float* __restrict__ seq, int n_cand, int lcap, int d4, bf16* __restrict__ dst) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * lcap * d4) return;
    const int k = idx % d4;
    const int t = (idx / d4) % lcap;
    const int c = idx / ((size_t)d4 * lcap);
    const float4 v = __ldg(reinterpret_cast<const float4*>(seq) + idx);
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[((size_t)c * (lcap + 1) + t) * d4 + k] = u;
}

// pair rows for the inverse model / learned cost: row (c,t) = [seq[c][t] | seq[c][t+1]] as bf16,
// seq = zero-padded pruned latent sequence [n_cand][lcap][128].  `goal0` (optional, [128]) replaces the
// element just past a candidate's end (the first frame of the goal sequence, learned cost).
__global__ void make_pairs_kernel(const float* __restrict__ seq, const long long* __restrict__ end_ind,
                                  const float* __restrict__ goal0, int n_cand, int lcap, int rows_padded,
                                  bf16* __restrict__ pairs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows_padded * 256) return;
    const int k = idx & 255;
    const size_t row = idx >> 8;
    float v = 0.f;
    if (row < (size_t)n_cand * lcap) {
        const int c = row / lcap, t = row - (size_t)c * lcap;
        const int tt = t + (k >> 7);
        if (tt < lcap) v = seq[((size_t)c * lcap + tt) * 128 + (k & 127)];
        if (goal0 != nullptr && tt == (int)end_ind[c] + 1) v = goal0[k & 127];
    }
    pairs[idx] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// L2 image cost (gcp/planning/cem/cost_fcn.py:9-22,65-72): per kept frame sqrt(sum((img-goal)^2)),
// last frame weighted, dense sum or last only.  One block per candidate, one warp per frame,
// float4 coalesced reads; algorithmic bytes = (end_ind+1) * 12 KB per candidate.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cost_l2_kernel(const float* __restrict__ images, const int* __restrict__ frame_node,
                                                      const long long* __restrict__ end_ind, const float* __restrict__ goal,
                                                      int n_nodes, int lcap, int dense, float final_w,
                                                      float* __restrict__ cost, int min_last = 1) {
    __shared__ float part[8];
    const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int last = max((int)end_ind[c], min_last);   // cem_simulator.py:31: end_ind = max(end_ind, 1)
    const float4* g4 = reinterpret_cast<const float4*>(goal);
    float acc = 0.f;
    for (int t = warp; t <= last; t += 8) {
        if (!dense && t != last) continue;
        const int node = frame_node != nullptr ? frame_node[c * lcap + t] : t;   // null map: frames stored in time order
        const float4* p = reinterpret_cast<const float4*>(images) + ((size_t)c * n_nodes + node) * 768;
        float s = 0.f;
#pragma unroll 4
        for (int k = lane; k < 768; k += 32) {
            const float4 a = __ldg(p + k), b = __ldg(g4 + k);
            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
            s += dx * dx + dy * dy + dz * dz + dw * dw;
        }
        s = warp_sum(s);
        acc += sqrtf(s) * (t == last ? final_w : 1.f);
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[w];
        cost[c] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Planner mode "decode only what balanced pruning keeps": the kept (candidate, frame) pairs are compacted into
// consecutive rows, candidate-major, frame order inside a candidate: row_off[c] = sum_{c' < c} n(c'),
// n(c) = max(end_ind[c], min_last) + 1 (the frames the cost reads, cem_simulator.py:31), row_off[B] = *n_rows.
// One CTA; B is a few thousand at most.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) kept_offsets_kernel(const long long* __restrict__ end_ind, int n_cand, int min_last,
                                                            int* __restrict__ row_off, int* __restrict__ n_rows) {
    __shared__ int wsum[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_cand; c0 += 1024) {
        const int c = c0 + tid;
        const int v = c < n_cand ? max((int)end_ind[c], min_last) + 1 : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (c < n_cand) row_off[c] = off + inc - v;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += wsum[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        row_off[n_cand] = s_base;
        *n_rows = s_base;
    }
}

// latc[row_off[c] + t][:] = lat[(frame_node[c][t] + 1) * Bp + c][:] (bf16 rows of 128 = 16 x 16 B), + the row's
// (candidate, node).  Thread per 16-byte piece.
__global__ void gather_kept_rows_kernel(const bf16* __restrict__ lat, const int* __restrict__ frame_node,
                                        const long long* __restrict__ end_ind, const int* __restrict__ row_off, int n_cand,
                                        int Bp, int lcap, int min_last, bf16* __restrict__ latc, int* __restrict__ row_cand,
                                        int* __restrict__ row_node) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * lcap * 16) return;
    const int k = idx & 15;
    const int t = (idx >> 4) % lcap;
    const int c = idx / ((size_t)16 * lcap);
    if (t > max((int)end_ind[c], min_last)) return;
    const int node = frame_node[c * lcap + t];
    const int r = row_off[c] + t;
    reinterpret_cast<uint4*>(latc)[(size_t)r * 16 + k] = __ldg(reinterpret_cast<const uint4*>(lat) + ((size_t)(node + 1) * Bp + c) * 16 + k);
    if (k == 0) {
        row_cand[r] = c;
        row_node[r] = node;
    }
}

// cost[c] = sum over the kept frames t of sqrt(frame_sq[..]) (last frame weighted; dense = 0: last frame only), from the
// per-image sums of squares the decoder tail wrote (dec_tail3.cuh, fused L2 cost).  frame_node != null: frame_sq is
// [c][n_nodes] (all nodes decoded); null: compact rows row_off[c] + t.  One warp per candidate, lanes stride the frames,
// then the butterfly: the same order in both layouts, so the two decode modes give bit-identical costs.
__global__ void cost_from_frames_kernel(const float* __restrict__ frame_sq, const int* __restrict__ frame_node,
                                        const int* __restrict__ row_off, const long long* __restrict__ end_ind, int n_nodes,
                                        int lcap, int dense, float final_w, int min_last, float* __restrict__ cost) {
    const int c = blockIdx.x, lane = threadIdx.x;
    const int last = max((int)end_ind[c], min_last);
    float s = 0.f;
    for (int t = lane; t <= last; t += 32) {
        if (!dense && t != last) continue;
        const float q = frame_node != nullptr ? frame_sq[(size_t)c * n_nodes + frame_node[c * lcap + t]] : frame_sq[row_off[c] + t];
        s += sqrtf(q) * (t == last ? final_w : 1.f);
    }
    s = warp_sum(s);
    if (lane == 0) cost[c] = s;
}

// ---------------------------------------------------------------------------------------------
// Planner mode, tree side: which (node, 128-candidate tile) pairs of every tree level some candidate keeps.  A node is
// kept iff its interval in the balanced-pruning recursion still has an interior point (frame_binding.py:42-65) -- the same
// integer recursion as prune_map_kernel -- and a kept node's ancestors and interval ends are kept, so a level only
// needs the listed tiles and every operand it reads was computed.  "Kept" is monotone in the rollout length (a node kept
// at length L is kept at every longer one: tests/test_host_logic.py checks it exhaustively for depth <= 8), so a tile
// needs a node iff its LONGEST candidate keeps it.  One CTA per level: the warps reduce the tiles' maximum lengths into
// shared memory, a thread decides one (node, tile) pair, an ordered ballot / prefix compaction writes the list (ascending
// tile index; padded to an even count with a copy of the last entry, for the CTA-pair GEMMs); rows[l] = listed tiles * 128.
// tiles: [depth] lists at offsets off(l) = (2^l - 1) * tpn + 2 * l.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool node_kept(int end, int level, int j) {
    int l = -1, r = end + 1, t = 0;
    for (int lv = 0; lv <= level; ++lv) {
        t = (l + r) / 2;
        if (lv < level) {
            if ((j >> (level - 1 - lv)) & 1) l = t; else r = t;
        }
    }
    return t != l && t != r;
}
constexpr int TREE_TPN_MAX = 512;      // 128-candidate tiles per call the work lists are built for (65 536 candidates)
__host__ __device__ __forceinline__ int tree_tiles_offset(int level, int tpn) { return ((1 << level) - 1) * tpn + 2 * level; }

__global__ void __launch_bounds__(1024) tree_worklists_kernel(const long long* __restrict__ end_ind, int n_cand, int Bp,
                                                              int min_last, int* __restrict__ tiles, int* __restrict__ rows) {
    __shared__ int wsum[32];
    __shared__ int s_base;
    const int level = blockIdx.x, tpn = Bp >> 7;
    const int n_tiles = tpn << level;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* list = tiles + tree_tiles_offset(level, tpn);
    __shared__ int tmax[TREE_TPN_MAX];      // longest rollout of each 128-candidate tile (-1: no candidate)
    if (tid == 0) s_base = 0;
    for (int t = warp; t < tpn; t += 32) {
        int m = -1;
        for (int c = (t << 7) + lane; c < min((t + 1) << 7, n_cand); c += 32) m = max(m, max((int)end_ind[c], min_last));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (lane == 0) tmax[t] = m;
    }
    __syncthreads();
    for (int t0 = 0; t0 < n_tiles; t0 += 1024) {
        const int tile = t0 + tid;
        bool act = false;
        if (tile < n_tiles) {
            const int j = tile / tpn, m = tmax[tile - j * tpn];
            act = m >= 0 && node_kept(m, level, j);
        }
        const unsigned b = __ballot_sync(0xffffffffu, act);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (act) list[off + __popc(b & ((1u << lane) - 1u))] = tile;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += wsum[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int n = s_base;
        if (n & 1) { list[n] = list[n - 1]; ++n; }
        rows[level] = n * 128;
    }
}

// Pair rows for the learned pairwise networks from two row tables: pairs[r] = cat(a[ia[r]], b[ib[r]]) (a null index
// list means row r itself).  Used by gcpb200_cost_pairs (LearnedCostEstimate ndarray branch, cost_fcn.py:84-87) and
// gcpb200_infer_action (InverseModel.run_single, inverse_mdl.py:221-224).  Rows >= n are zero padding.
__global__ void make_pairs_idx_kernel(const float* __restrict__ a, const int* __restrict__ ia, const float* __restrict__ b,
                                      const int* __restrict__ ib, int n, int rows_padded, bf16* __restrict__ pairs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows_padded * 256) return;
    const int k = idx & 255;
    const int row = (int)(idx >> 8);
    float v = 0.f;
    if (row < n) {
        if (k < 128) v = a[(size_t)(ia != nullptr ? ia[row] : row) * 128 + k];
        else v = b[(size_t)(ib != nullptr ? ib[row] : row) * 128 + (k - 128)];
    }
    pairs[idx] = __float2bfloat16_rn(v);
}

// cost[s] = sum of rowcost[seg_off[s] .. seg_off[s+1]) in a fixed (lane-strided, then butterfly) order: the summed
// sequence cost of LearnedCostEstimate's list branch (cost_fcn.py:88-97).  One warp per segment.
__global__ void seg_sum_kernel(const float* __restrict__ rowcost, const int* __restrict__ seg_off, float* __restrict__ cost) {
    const int sgm = blockIdx.x, lane = threadIdx.x;
    float s = 0.f;
    for (int t = seg_off[sgm] + lane; t < seg_off[sgm + 1]; t += 32) s += rowcost[t];
    s = warp_sum(s);
    if (lane == 0) cost[sgm] = s;
}

// learned cost reduction: cost[c] = sum_{t <= end_ind[c]} rowcost[c*lcap + t] + tail
__global__ void cost_sum_kernel(const float* __restrict__ rowcost, const long long* __restrict__ end_ind, int lcap,
                                const float* __restrict__ tail, int n_tail, float* __restrict__ cost) {
    const int c = blockIdx.x, lane = threadIdx.x;
    float s = 0.f;
    for (int t = lane; t <= (int)end_ind[c] && t < lcap; t += 32) s += rowcost[(size_t)c * lcap + t];
    for (int t = lane; t < n_tail; t += 32) s += tail[t];
    s = warp_sum(s);
    if (lane == 0) cost[c] = s;
}

// ---------------------------------------------------------------------------------------------
// Elite selection (gcp/planning/cem/cem_planner.py:124-135): `scores.argsort()[:k]`, the k lowest costs in ascending
// order.  Two launches, exact and deterministic, O(N) + O(k^2 / chip):
//   1. topk_select_kernel (one CTA): MSB-first radix select (8 bits per pass, warp-aggregated shared-memory
//      histograms: __match_any_sync groups the lanes of a warp by bin, one atomic per group) of the k-th smallest
//      64-bit composite key (order-preserving cost bits << 32 | index), then an ordered ballot / prefix-sum
//      compaction of the k composites that are <= it.
//   2. topk_rank_kernel: rank of every selected composite among the k by counting (tiles in shared memory) -> its
//      output position.
// Order rule = numpy's: ascending, every NaN after +inf, -0.0 == +0.0; ties (numpy's quicksort leaves them
// unspecified) are broken by index, so every rank of a sharded planner selects the same elites.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long topk_composite(float x, int i) {
    uint32_t key;
    if (x != x) {
        key = 0xFFFFFFFFu;                       // NaN sorts last
    } else {
        const uint32_t b = __float_as_uint(x + 0.0f);      // -0.0 + 0.0 = +0.0
        key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    return ((unsigned long long)key << 32) | (uint32_t)i;
}

constexpr int TOPK_SELECT_THREADS = 1024;
__global__ void __launch_bounds__(TOPK_SELECT_THREADS) topk_select_kernel(const float* __restrict__ cost, int n, int k,
                                                                          unsigned long long* __restrict__ sel) {
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_krem;
    __shared__ int wsum[TOPK_SELECT_THREADS / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_prefix = 0ull; s_krem = k; s_base = 0; }
    unsigned long long mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i0 = 0; i0 < n; i0 += TOPK_SELECT_THREADS) {
            const int i = i0 + tid;
            int bin = -1;
            if (i < n) {
                const unsigned long long c = topk_composite(cost[i], i);
                if ((c & mask) == prefix) bin = (int)((c >> shift) & 255ull);
            }
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            // 8 bins per lane: inclusive warp scan of the lane totals, then the lane that crosses k_rem picks its bin
            int h[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q] = hist[lane * 8 + q]; tot += h[q]; }
            int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int krem = s_krem, before = inc - tot;
            __syncwarp();                  // every lane has read s_krem / s_prefix before the selecting lane rewrites them
            if (before < krem && krem <= inc) {
                int cum = before;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (cum < krem && krem <= cum + h[q]) {
                        s_prefix = prefix | ((unsigned long long)(lane * 8 + q) << shift);
                        s_krem = krem - cum;
                    }
                    cum += h[q];
                }
            }
        }
        mask |= 255ull << shift;
        __syncthreads();
    }
    // composites are unique, so exactly k of them are <= the k-th smallest: compact them in index order
    const unsigned long long kth = s_prefix;
    for (int i0 = 0; i0 < n; i0 += TOPK_SELECT_THREADS) {
        const int i = i0 + tid;
        const unsigned long long c = i < n ? topk_composite(cost[i], i) : ~0ull;
        const bool take = i < n && c <= kth;
        const unsigned b = __ballot_sync(0xffffffffu, take);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (take) sel[off + __popc(b & ((1u << lane) - 1u))] = c;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < TOPK_SELECT_THREADS / 32; ++w) t += wsum[w];
            s_base += t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) topk_rank_kernel(const unsigned long long* __restrict__ sel, int k,
                                                        const float* __restrict__ cost, int* __restrict__ idx_out,
                                                        float* __restrict__ val_out) {
    __shared__ unsigned long long tile[1024];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const unsigned long long ci = i < k ? sel[i] : ~0ull;
    int rank = 0;
    for (int base = 0; base < k; base += 1024) {
        for (int t = threadIdx.x; t < 1024; t += 256) tile[t] = base + t < k ? sel[base + t] : ~0ull;
        __syncthreads();
        const int lim = min(1024, k - base);
#pragma unroll 8
        for (int t = 0; t < lim; ++t) rank += tile[t] < ci;
        __syncthreads();
    }
    if (i < k) {
        const int id = (int)(uint32_t)ci;
        idx_out[rank] = id;
        if (val_out != nullptr) val_out[rank] = cost[id];
    }
}

// ---------------------------------------------------------------------------------------------
// Refit (gcp/planning/cem/sampler.py:44-46): mean / std (ddof 0) over the elite samples per (node, dim), float64
// accumulation like numpy.  HBM-bound: k elite rows of `per_cand` floats are read exactly once (float4, coalesced);
// the k elites are split over gridDim.y CTAs per column block so that the whole chip streams, and every split writes
// its (sum d, sum d^2) of the shifted values d = x - x_first (exact in float64; identical elites give std == 0
// exactly) to `part`; refit_final_kernel adds the splits in a fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) refit_partial_kernel(const float* __restrict__ z, const int* __restrict__ elite, int k,
                                                            int per_cand, double* __restrict__ part) {
    const int e4 = blockIdx.x * blockDim.x + threadIdx.x;        // float4 column
    if (e4 >= per_cand / 4) return;
    const int S = gridDim.y, s = blockIdx.y;
    const int i0 = (int)((long long)k * s / S), i1 = (int)((long long)k * (s + 1) / S);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(z + (size_t)elite[0] * per_cand) + e4);
    double a[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(z + (size_t)elite[i] * per_cand) + e4);
        const double d0 = (double)v.x - (double)x0.x, d1 = (double)v.y - (double)x0.y;
        const double d2 = (double)v.z - (double)x0.z, d3 = (double)v.w - (double)x0.w;
        a[0] += d0; a[1] += d1; a[2] += d2; a[3] += d3;
        q[0] = fma(d0, d0, q[0]); q[1] = fma(d1, d1, q[1]); q[2] = fma(d2, d2, q[2]); q[3] = fma(d3, d3, q[3]);
    }
    double* o = part + ((size_t)s * per_cand + (size_t)e4 * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = a[j]; o[2 * j + 1] = q[j]; }
}
__global__ void refit_final_kernel(const float* __restrict__ z, const int* __restrict__ elite, int k, int per_cand, int S,
                                   const double* __restrict__ part, float* __restrict__ mean, float* __restrict__ stdv) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_cand) return;
    double a = 0.0, q = 0.0;
    for (int s = 0; s < S; ++s) {
        a += part[((size_t)s * per_cand + e) * 2];
        q += part[((size_t)s * per_cand + e) * 2 + 1];
    }
    const double x0 = (double)__ldg(z + (size_t)elite[0] * per_cand + e);
    const double md = a / k;
    const double var = q / k - md * md;
    mean[e] = (float)(x0 + md);
    stdv[e] = (float)sqrt(var > 0.0 ? var : 0.0);
}

// ---------------------------------------------------------------------------------------------
// Noise: z = clip(mean + std * n, +-clip), n ~ N(0,1) from Philox4x32-10 keyed by (seed, global
// candidate id) so any rank can regenerate any candidate (FlatCEMSampler.sample, sampler.py:40-42).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__global__ void sample_noise_kernel(const float* __restrict__ mean, const float* __restrict__ stdv, float std_scalar,
                                    unsigned long long seed, unsigned long long cand0, const int* __restrict__ ids,
                                    int n_cand, int per_cand, float clip, float* __restrict__ z) {
    // grid = (candidate, blocks of per_cand / 4): one thread = 4 outputs, no division
    const int per4 = per_cand >> 2;
    const int c = blockIdx.x, e4 = blockIdx.y * blockDim.x + threadIdx.x;
    if (e4 >= per4) return;
    const size_t idx = (size_t)c * per4 + e4;
    const unsigned long long gid = ids != nullptr ? (unsigned long long)ids[c] : cand0 + c;
    uint32_t ctr[4] = {(uint32_t)e4, 0u, (uint32_t)gid, (uint32_t)(gid >> 32)};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    float n[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = ((ctr[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((ctr[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-2.0f * __logf(u1));
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        n[2 * h] = r * cs;
        n[2 * h + 1] = r * sn;
    }
    const float4 m4 = mean != nullptr ? __ldg(reinterpret_cast<const float4*>(mean) + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 s4 = stdv != nullptr ? __ldg(reinterpret_cast<const float4*>(stdv) + e4)
                                      : make_float4(std_scalar, std_scalar, std_scalar, std_scalar);
    const float* mp = reinterpret_cast<const float*>(&m4);
    const float* sp = reinterpret_cast<const float*>(&s4);
    float4 o;
    float* op = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) op[q] = fminf(fmaxf(mp[q] + sp[q] * n[q], -clip), clip);
    reinterpret_cast<float4*>(z)[idx] = o;
}

// Upload of the noise of one set of tree levels straight from pinned host memory (zero-copy reads over PCIe):
// rows (cand, node = a*k + b), k < cnt, of the [B][n_nodes][row4 float4] array.  Depth-first node index of level l,
// position j is (2j+1)*2^(7-l) - 1, so "levels <= L" is (a,b,cnt) = (2^(7-L), a-1, 2^(L+1)-1) and level l alone is
// (2^(8-l), 2^(7-l)-1, 2^l).
// Launched with few, small blocks (128 threads, ~40 registers) so that it co-resides with the persistent GEMM / decoder
// CTAs of the rollout stream instead of taking SMs away from them; PCIe needs < 1 MB of reads in flight.
__global__ void __launch_bounds__(128) upload_rows_kernel(const float4* __restrict__ src_host, float4* __restrict__ dst,
                                                          int n_cand, int n_nodes, int row4, int a, int b, int cnt) {
    const size_t total = (size_t)n_cand * cnt * row4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent 16-byte host reads in flight per thread
    for (; idx + 3 * stride < total; idx += 4 * stride) {
        float4 v[4];
        size_t off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t i = idx + u * stride;
            const int e = i % row4;
            const size_t r = i / row4;
            const int k = r % cnt, c = r / cnt;
            off[u] = ((size_t)c * n_nodes + (a * k + b)) * row4 + e;
            v[u] = __ldcs(src_host + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[off[u]] = v[u];
    }
    for (; idx < total; idx += stride) {
        const int e = idx % row4;
        const size_t r = idx / row4;
        const int k = r % cnt, c = r / cnt;
        const size_t o = ((size_t)c * n_nodes + (a * k + b)) * row4 + e;
        dst[o] = __ldcs(src_host + o);
    }
}

// slot-major [n_slots*Bp][cols] (rows = slot, cand) -> candidate-major depth-first [B][n_nodes][cols]
__global__ void slot_to_df_kernel(const float* __restrict__ src, int Bp, int n_cand, int n_nodes, int cols,
                                  int src_ld, float* __restrict__ dst) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * n_nodes * cols) return;
    const int k = idx % cols;
    const int node = (idx / cols) % n_nodes;
    const int c = idx / ((size_t)cols * n_nodes);
    dst[idx] = src[((size_t)(node + 1) * Bp + c) * src_ld + k];
}

// Sequential rollout: time-major latent rows [n_frames][Bp][128] (slot t = latent of frame t) -> candidate-major
// zero-padded sequence dst[c][t][:] = t <= end_ind[c] ? src[t][c][:] : 0  (pad_sequence of cat(e_0, encodings)[:end+1],
// gcp/prediction/models/sequential.py:78-94, base_gcp.py:242).  Thread per float4.
__global__ void seq_gather_kernel(const float* __restrict__ src, const long long* __restrict__ end_ind, int Bp, int n_cand,
                                  int lcap, float* __restrict__ dst) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * lcap * 32) return;
    const int k = idx & 31;
    const int t = (idx >> 5) % lcap;
    const int c = idx / ((size_t)32 * lcap);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t <= (int)end_ind[c]) v = __ldg(reinterpret_cast<const float4*>(src) + ((size_t)t * Bp + c) * 32 + k);
    reinterpret_cast<float4*>(dst)[idx] = v;
}

// Sequential rollout: frame 0 of every candidate's image sequence is the start image itself
// (gcp/prediction/models/sequential.py:57).  images [B][n_frames][3072]; I_0 [B or 1][3072].
__global__ void copy_frame0_kernel(const float* __restrict__ I_0, int shared, int n_cand, int n_frames,
                                   float* __restrict__ images) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * 768) return;
    const int k = idx % 768, c = idx / 768;
    reinterpret_cast<float4*>(images)[(size_t)c * n_frames * 768 + k] =
        __ldg(reinterpret_cast<const float4*>(I_0) + (shared ? 0 : (size_t)c * 768) + k);
}

// AdaptiveBinding.prune_sequence (gcp/prediction/models/adaptive_binding/adaptive.py:62-77): node 0 is always kept,
// node n > 0 is dropped when sigmoid(distance[n-1]) > threshold, i.e. distance > logit(threshold).  One block per
// candidate; the kept depth-first indices are compacted in order (ballot + warp prefix), len = their count.
__global__ void __launch_bounds__(256) adaptive_prune_kernel(const float* __restrict__ dist, int dist_ld, int n_nodes,
                                                             float logit_thr, int* __restrict__ nodes, int* __restrict__ len,
                                                             long long* __restrict__ end_out) {
    __shared__ int wsum[8];
    const int c = blockIdx.x, n = threadIdx.x, warp = n >> 5, lane = n & 31;
    const bool keep = n < n_nodes && (n == 0 || !(dist[(size_t)c * dist_ld + n - 1] > logit_thr));
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    if (keep) nodes[(size_t)c * n_nodes + base + __popc(b & ((1u << lane) - 1u))] = n;
    if (n == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        len[c] = tot;
        if (end_out != nullptr) end_out[c] = tot - 1;
    }
}

// dst[c][t][:] = t < len[c] ? src[c][nodes[c][t]][:] : 0   (thread per float4; lcap rows per candidate)
__global__ void gather_nodes_kernel(const float* __restrict__ src, const int* __restrict__ nodes, const int* __restrict__ len,
                                    int n_cand, int n_nodes, int d4, float* __restrict__ dst) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_cand * n_nodes * d4) return;
    const int k = idx % d4;
    const int t = (idx / d4) % n_nodes;
    const int c = idx / ((size_t)d4 * n_nodes);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len[c]) v = __ldg(reinterpret_cast<const float4*>(src) + ((size_t)c * n_nodes + nodes[(size_t)c * n_nodes + t]) * d4 + k);
    reinterpret_cast<float4*>(dst)[idx] = v;
}

__global__ void len_to_end_kernel(const int* __restrict__ len, long long* __restrict__ end, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) end[i] = (long long)len[i] - 1;
}

__global__ void fill_i64_kernel(long long* p, long long v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Optimiser step (train.py:162): RAdam (blox/torch/radam.py:17-80) / torch.optim.Adam on one flat fp32 array, in place,
// with the gradient scaling of clip_grad_norm_ (blox/torch/training.py:154-159) folded in.  Elementwise, float4 accesses,
// 16 B read + 12 B written per parameter: HBM-bound.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sq_norm_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
    __shared__ double part[8];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(acc, t);
    }
}

struct OptimArgs {
    float* p;
    const float* g;
    float *m, *v;
    long long n;
    int radam;            // 0: Adam, 1: RAdam
    int rectified;        // RAdam: N_sma >= 5 (variance-rectified step), else the SGD-with-momentum step
    float beta1, beta2, eps;
    float omb1, omb2;     // 1 - beta, rounded from the double-precision difference as torch rounds the Python scalar
    float step_size;      // RAdam: step_size * lr (radam.py:62-78);  Adam: lr / (1 - beta1^t)
    float sqrt_bc2;       // Adam: sqrt(1 - beta2^t)
    float decay;          // RAdam: weight_decay * lr (p += -decay * p);  Adam: weight_decay (g += decay * p)
    const double* grad_sq_norm;
    float max_norm;
};
__device__ __forceinline__ void optim_one(const OptimArgs& a, float gs, float& p, float g, float& m, float& v) {
    g *= gs;
    if (a.radam) {
        v = v * a.beta2 + a.omb2 * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(1 - beta2, grad, grad)
        m = m * a.beta1 + a.omb1 * g;                     // exp_avg.mul_(beta1).add_(1 - beta1, grad)
        if (a.decay != 0.f) p += -a.decay * p;
        if (a.rectified) p += -a.step_size * (m / (sqrtf(v) + a.eps));
        else p += -a.step_size * m;
    } else {
        if (a.decay != 0.f) g += a.decay * p;
        m = m + a.omb1 * (g - m);                         // exp_avg.lerp_(grad, 1 - beta1)
        v = v * a.beta2 + a.omb2 * g * g;
        p += -a.step_size * (m / (sqrtf(v) / a.sqrt_bc2 + a.eps));
    }
}
__global__ void __launch_bounds__(256) optim_step_kernel(const OptimArgs a) {
    float gs = 1.0f;
    if (a.grad_sq_norm != nullptr && a.max_norm > 0.f) {
        const float coef = a.max_norm / ((float)sqrt(*a.grad_sq_norm) + 1e-6f);
        gs = coef < 1.0f ? coef : 1.0f;
    }
    const long long n4 = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
        optim_one(a, gs, p.x, g.x, m.x, v.x);
        optim_one(a, gs, p.y, g.y, m.y, v.y);
        optim_one(a, gs, p.z, g.z, m.z, v.z);
        optim_one(a, gs, p.w, g.w, m.w, v.w);
        reinterpret_cast<float4*>(a.p)[i] = p;
        reinterpret_cast<float4*>(a.m)[i] = m;
        reinterpret_cast<float4*>(a.v)[i] = v;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        optim_one(a, gs, a.p[i], a.g[i], a.m[i], a.v[i]);
}

}  // namespace gcp
