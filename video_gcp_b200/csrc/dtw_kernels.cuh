// DTW family for sm_100a: pairwise mean-squared-distance cost matrices, the float64 soft-DTW
// forward-backward of the adaptive binding (row sweep: the reference forbids horizontal moves, so row i only depends on
// row i-1) and the metric-time DTW (anti-diagonal wavefront + traceback + per-frame best match).
//   reference: gcp/prediction/models/adaptive_binding/probabilistic_dtw.py:11-121, adaptive.py:41-61,
//              gcp/evaluation/dtw_utils.py:77-130,201-241, gcp/evaluation/evaluation_matching.py:135-147,
//              blox/torch/ops.py:62-91
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace gcp {

// ---------------------------------------------------------------------------------------------
// cost matrix: out[b][i][j] = mean_d (x[b][i][d] - y[b][j][d])^2      (batch_cdist(..., 'mean'), blox/torch/ops.py:62-91)
// fp32 FMA-bound (2 instructions per pair element); 64 x 64 output tile per CTA, 4 x 4 per thread, K staged through
// shared memory in chunks of 32 with the partial sum of every chunk folded into the total (shorter rounding chains than
// the reference's |x|^2 + |y|^2 - 2xy expansion, which cancels catastrophically for close vectors).
// ---------------------------------------------------------------------------------------------
static const int CD_T = 64, CD_K = 32, CD_LD = 68;

__global__ void __launch_bounds__(256) cdist_mean_kernel(const float* __restrict__ x, const float* __restrict__ y, int n,
                                                          int m, int dim, float* __restrict__ out) {
    __shared__ __align__(16) float xs[CD_K][CD_LD];
    __shared__ __align__(16) float ys[CD_K][CD_LD];
    const int b = blockIdx.z, i0 = blockIdx.y * CD_T, j0 = blockIdx.x * CD_T;
    const float* xb = x + (size_t)b * n * dim;
    const float* yb = y + (size_t)b * m * dim;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lv = tid >> 2, lk = (tid & 3) * 8;          // loader: vector lv of the tile, 8 consecutive k
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int k0 = 0; k0 < dim; k0 += CD_K) {
        float4 xv[2], yv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + lk + 4 * h;
            xv[h] = (i0 + lv < n && k < dim) ? *reinterpret_cast<const float4*>(xb + (size_t)(i0 + lv) * dim + k) : make_float4(0, 0, 0, 0);
            yv[h] = (j0 + lv < m && k < dim) ? *reinterpret_cast<const float4*>(yb + (size_t)(j0 + lv) * dim + k) : make_float4(0, 0, 0, 0);
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            xs[lk + 4 * h + 0][lv] = xv[h].x; xs[lk + 4 * h + 1][lv] = xv[h].y; xs[lk + 4 * h + 2][lv] = xv[h].z; xs[lk + 4 * h + 3][lv] = xv[h].w;
            ys[lk + 4 * h + 0][lv] = yv[h].x; ys[lk + 4 * h + 1][lv] = yv[h].y; ys[lk + 4 * h + 2][lv] = yv[h].z; ys[lk + 4 * h + 3][lv] = yv[h].w;
        }
        __syncthreads();
        float part[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) part[a][c] = 0.f;
#pragma unroll
        for (int k = 0; k < CD_K; ++k) {
            const float4 xa = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
            const float4 ya = *reinterpret_cast<const float4*>(&ys[k][tx * 4]);
            const float xr[4] = {xa.x, xa.y, xa.z, xa.w}, yr[4] = {ya.x, ya.y, ya.z, ya.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float d = xr[a] - yr[c];
                    part[a][c] = fmaf(d, d, part[a][c]);
                }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] += part[a][c];
    }
    const float fdim = (float)dim;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx * 4 + c;
            if (j < m) out[((size_t)b * n + i) * m + j] = __fdiv_rn(acc[a][c], fdim);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// soft-DTW sweep.  One CTA per (sequence, direction); accum [2][B][r][c] float64 in UN-flipped coordinates.
//   C = -(double)(cost / temp)                                      (adaptive.py:51, probabilistic_dtw.py:91)
//   D[0][j] = C[0][begin] at j == begin, else -inf                  (probabilistic_dtw.py:32-33)
//   D[i][j] = C[i][j] + logsumexp(D[i-1][j], D[i-1][j-1])           (probabilistic_dtw.py:52-62)
// The reference's column index j-1 = -1 wraps to the last column as it was BEFORE the diagonal's writes; that value is
// -inf except for (i == 1, begin == c-1) and for c == 1 -- reproduced (oracle/dtw_oracle.py::gak_table).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double lse_pair(double a, double b) {
    // torch.logsumexp over two values: log(exp(a - m) + exp(b - m)) + m with m = max (0 when infinite); exp(0) == 1
    const double m = fmax(a, b), lo = fmin(a, b);
    if (m == -CUDART_INF) return -CUDART_INF;
    return log(1.0 + exp(lo - m)) + m;
}

__global__ void __launch_bounds__(1024) soft_dtw_sweep_kernel(const float* __restrict__ cost, float temp,
                                                               const long long* __restrict__ end_inds, int B, int r, int c,
                                                               double* __restrict__ accum) {
    extern __shared__ double sd_rows[];      // two rows of c doubles
    const int b = blockIdx.x % B, dir = blockIdx.x / B;
    const float* cb = cost + (size_t)b * r * c;
    double* ab = accum + ((size_t)dir * B + b) * r * c;
    const long long end = end_inds ? end_inds[b] : (long long)(c - 1);
    if (end < 0 || end >= c) {               // the reference raises an index error; leave an unmistakable table
        for (size_t k = threadIdx.x; k < (size_t)r * c; k += blockDim.x) ab[k] = CUDART_NAN;
        return;
    }
    const int begin = dir ? (int)(c - end - 1) : 0;
    // flipped coordinates (i, j) of the backward pass address (r-1-i, c-1-j) of the cost matrix and of the output table
    auto at = [&](int i, int j) -> size_t { return dir ? (size_t)(r - 1 - i) * c + (c - 1 - j) : (size_t)i * c + j; };
    auto C = [&](int i, int j) -> double { return -(double)__fdiv_rn(cb[at(i, j)], temp); };
    double* prev = sd_rows;
    double* cur = sd_rows + c;
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        const double v = (j == begin) ? C(0, j) : -CUDART_INF;
        prev[j] = v;
        ab[at(0, j)] = v;
    }
    __syncthreads();
    const int j1 = threadIdx.x;              // fast path (c <= blockDim): one column per thread, next row's cost prefetched
    double cn = (j1 < c && r > 1) ? C(1, j1) : 0.0;
    for (int i = 1; i < r; ++i) {
        const bool wrap = (c == 1) || (i == 1 && begin == c - 1);
        for (int j = threadIdx.x; j < c; j += blockDim.x) {
            double cij;
            if (j == j1) {
                cij = cn;
                if (i + 1 < r) cn = C(i + 1, j);
            } else {
                cij = C(i, j);
            }
            const double step = j > 0 ? prev[j - 1] : (wrap ? prev[c - 1] : -CUDART_INF);
            const double v = cij + lse_pair(prev[j], step);
            cur[j] = v;
            ab[at(i, j)] = v;
        }
        __syncthreads();
        double* t = prev;
        prev = cur;
        cur = t;
    }
}

// w = exp(forward + backward - C - z), z = forward[b][r-1][end]; -inf costs give 0       (probabilistic_dtw.py:108-113)
__global__ void soft_dtw_weights_kernel(const float* __restrict__ cost, float temp, const long long* __restrict__ end_inds,
                                        const double* __restrict__ accum, int B, int r, int c, float* __restrict__ w) {
    const size_t n = (size_t)B * r * c;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int b = (int)(k / ((size_t)r * c));
    long long end = end_inds ? end_inds[b] : (long long)(c - 1);
    if (end < 0 || end >= c) {
        w[k] = CUDART_NAN_F;
        return;
    }
    const double Cv = -(double)__fdiv_rn(cost[k], temp);
    const double z = accum[((size_t)b * r + (r - 1)) * c + end];
    double e = accum[k] + accum[n + k] - Cv;
    if (Cv == -CUDART_INF) e = -CUDART_INF;
    w[k] = (float)exp(e - z);
}

// max over (b, node) of sum_t w[b][node][t]: the reference's stability check `w.sum(2).max() ~ 1`  (probabilistic_dtw.py:115-117)
__global__ void soft_dtw_rowsum_max_kernel(const float* __restrict__ w, int rows, int c, float* __restrict__ out) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float s = 0.f;
    for (int j = lane_id(); j < c; j += 32) s += w[(size_t)row * c + j];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane_id() == 0) {
        if (s != s) s = CUDART_INF_F;        // NaN must fail the check
        atomicMax(reinterpret_cast<int*>(out), __float_as_int(fmaxf(s, 0.f)));
    }
}

// normalize(w, dim=1, eps) + depthfirst2breadthfirst            (blox/torch/dist.py:22-24; tree_utils.py:217-232)
// thread per (b, frame): sums the column over the nodes, then writes w / max(sum, eps) to the breadth-first row.
__global__ void binding_normalize_kernel(const float* __restrict__ w, int B, int r, int c, int depth, float eps,
                                         float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= c) return;
    const float* wb = w + (size_t)b * r * c;
    float* ob = out + (size_t)b * r * c;
    float s = 0.f;
    for (int i = 0; i < r; ++i) s += wb[(size_t)i * c + j];
    const float norm = fmaxf(s, eps);
    for (int i = 0; i < r; ++i) {
        int pos = i;
        if (depth > 0) {
            const int t = __ffs(i + 1) - 1;                      // level counted from the leaves
            pos = ((1 << (depth - 1 - t)) - 1) + ((i + 1) >> (t + 1));
        }
        ob[(size_t)pos * c + j] = __fdiv_rn(wb[(size_t)i * c + j], norm);
    }
}

// ---------------------------------------------------------------------------------------------
// metric-time DTW.  One CTA per sequence.  acc [B][r+1][c+1] float64 is the reference's padded table:
//   D[0][0] = 0, D[0][1:] = D[1:][0] = inf, D[i+1][j+1] = C[i][j] + min(D[i][j], D[i+1][j], D[i][j+1])
// (dtw_utils.py:85-94, cutils.pyx:21-28), swept by anti-diagonals with the last two diagonals in shared memory; then one
// thread walks the path back from (r-1, end) with the first-minimum rule of np.argmin over (diagonal, up, left)
// (dtw_utils.py:201-219,222-241) and records, per column, the path row with the smallest accumulated cost
// (evaluation_matching.py:142-146).  Paths are written right-aligned into [r+c-1] slots (start -> end order).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) dtw_wavefront_kernel(const T* __restrict__ cost, const long long* __restrict__ end_ind,
                                                              int r, int c, double* __restrict__ acc, double* __restrict__ dist,
                                                              int* __restrict__ path_p, int* __restrict__ path_q,
                                                              int* __restrict__ path_len, int* __restrict__ match_inds) {
    extern __shared__ double dg[];           // three diagonals indexed by row: [3][r]
    const int b = blockIdx.x;
    const T* cb = cost + (size_t)b * r * c;
    double* D = acc + (size_t)b * (r + 1) * (c + 1);
    const int ldD = c + 1;
    for (int k = threadIdx.x; k <= c; k += blockDim.x) D[k] = k ? CUDART_INF : 0.0;
    for (int k = threadIdx.x + 1; k <= r; k += blockDim.x) D[(size_t)k * ldD] = CUDART_INF;
    double *d0 = dg, *d1 = dg + r, *d2 = dg + 2 * r;   // d0 = current, d1 = previous, d2 = the one before
    for (int d = 0; d < r + c - 1; ++d) {
        const int lo = max(0, d - c + 1), hi = min(r - 1, d);
        for (int i = lo + threadIdx.x; i <= hi; i += blockDim.x) {
            const int j = d - i;
            const double diag = (i > 0 && j > 0) ? d2[i - 1] : ((i == 0 && j == 0) ? 0.0 : CUDART_INF);
            const double left = j > 0 ? d1[i] : CUDART_INF;        // (i, j-1)   = D[i+1][j]
            const double up = i > 0 ? d1[i - 1] : CUDART_INF;      // (i-1, j)   = D[i][j+1]
            const double v = (double)cb[(size_t)i * c + j] + fmin(diag, fmin(left, up));
            d0[i] = v;
            D[(size_t)(i + 1) * ldD + (j + 1)] = v;
        }
        __syncthreads();
        double* t = d2;
        d2 = d1;
        d1 = d0;
        d0 = t;
    }
    const int slots = r + c - 1;
    long long e = end_ind ? end_ind[b] : (long long)(c - 1);
    if (e < 0) e = 0;
    if (e > c - 1) e = c - 1;
    if (match_inds)
        for (int k = threadIdx.x; k < c; k += blockDim.x) match_inds[(size_t)b * c + k] = 0;   // np.argmin of an all-inf column
    __syncthreads();
    if (threadIdx.x == 0) {
        int i = r - 1, j = (int)e, n = 0;
        int* pp = path_p + (size_t)b * slots;
        int* pq = path_q + (size_t)b * slots;
        dist[b] = D[(size_t)r * ldD + (j + 1)] / (double)(r + j + 1);
        int col = -1;
        double best = CUDART_INF;
        while (true) {
            pp[slots - 1 - n] = i;
            pq[slots - 1 - n] = j;
            ++n;
            if (match_inds) {
                const double a = D[(size_t)(i + 1) * ldD + (j + 1)];
                if (j != col) {
                    col = j;
                    best = CUDART_INF;
                }
                if (a <= best && a < CUDART_INF) {       // walking towards smaller rows: <= keeps the first minimum
                    best = a;
                    match_inds[(size_t)b * c + j] = i;
                }
            }
            if (i == 0 && j == 0) break;
            const double x0 = D[(size_t)i * ldD + j], x1 = D[(size_t)i * ldD + (j + 1)], x2 = D[(size_t)(i + 1) * ldD + j];
            int tb = 0;
            double mn = x0;
            if (x1 < mn) { tb = 1; mn = x1; }
            if (x2 < mn) tb = 2;
            if (i == 0) tb = 2;                          // borders are +inf for every finite table; guards keep indices valid
            else if (j == 0) tb = 1;
            if (tb == 0) { --i; --j; }
            else if (tb == 1) --i;
            else --j;
        }
        path_len[b] = n;
        for (int k = 0; k < slots - n; ++k) pp[k] = pq[k] = 0;     // finished sequences keep re-emitting (0, 0)
    }
}

// out[k] = src[idx[k]] for rows of `row_f4` float4s (gen_images = estimates[inds], evaluation_matching.py:147)
__global__ void gather_rows_kernel(const float4* __restrict__ src, const int* __restrict__ idx, int n_out, int row_f4,
                                   float4* __restrict__ out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)n_out * row_f4) return;
    const int row = (int)(k / row_f4), q = (int)(k % row_f4);
    out[k] = src[(size_t)idx[row] * row_f4 + q];
}

}  // namespace gcp
