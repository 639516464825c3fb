"""CPU oracle of the DTW family: the soft-DTW forward-backward that binds tree nodes to frames when the
adaptive-binding model trains, and the metric-time DTW that matches predicted to ground-truth frames in evaluation.

TEST INFRASTRUCTURE ONLY (same rules as oracle/gcp_oracle.py: imported by tests/, never by the product package).

Parity status: PINNED against the reference itself.  `oracle/make_golden_dtw.py` runs the UNMODIFIED reference
`soft_dtw`, `AdaptiveBinding.get_w`'s post-processing, `basic_dtw`, `batched_dtw` and
`DTWEvalBinding.get_single_matches` on seeded inputs and stores inputs' seeds + outputs in tests/golden/dtw_*.npz;
tests/test_oracle_dtw.py checks every function below against them (float64 tables to 1e-12, integer paths exactly).

A numpy restatement written from the recurrences, NOT a transcription: the reference sweeps anti-diagonals with
masked scatters; because its soft-DTW forbids horizontal moves ('nohor'), row i only depends on row i-1, so the
restatement sweeps rows.  Paths cite /root/reference.
"""
import numpy as np
import torch

NEG_INF = -np.inf


def batch_cdist_mean(x1, x2):
    """batch_cdist(x1, x2, reduction='mean') (blox/torch/ops.py:62-91): squared L2 distance between every pair of
    (flattened) vectors via the quadratic expansion in fp32, clamped at 0, divided by the vector length.
    x1 [B,n,...], x2 [B,m,...] torch fp32 -> [B,n,m]."""
    x1 = x1.flatten(start_dim=2)
    x2 = x2.flatten(start_dim=2)
    n1 = x1.pow(2).sum(-1, keepdim=True)
    n2 = x2.pow(2).sum(-1, keepdim=True)
    res = torch.baddbmm(n2.transpose(-2, -1), x1, x2.transpose(-2, -1), alpha=-2).add_(n1)
    return res.clamp_min_(0) / x1.shape[2]


def _lse2(a, b):
    """torch.logsumexp over a pair (aten LogSumExp: subtract the max unless it is infinite)."""
    m = np.maximum(a, b)
    m0 = np.where(np.isinf(m), 0.0, m)
    with np.errstate(divide="ignore"):
        return np.log(np.exp(a - m0) + np.exp(b - m0)) + m0


def gak_table(C, begin):
    """fast_gak(C, 'nohor', begin_inds) for ONE sequence (probabilistic_dtw.py:11-72).  C [r,c] float64 log-costs.
    D[0,j] = C[0,begin] at j = begin, -inf elsewhere; D[i,j] = C[i,j] + logsumexp(D[i-1,j], D[i-1,j-1]).
    The reference indexes column j-1 = -1 for j = 0, which wraps to the LAST column read *before* this diagonal's
    writes: that is -inf except for (i = 1, begin = c-1), where it is the preset D[0,c-1], and for c = 1, where it is
    the same column.  Kept, so that the oracle equals the reference on every input."""
    r, c = C.shape
    assert r >= c                                                                  # probabilistic_dtw.py:35
    D = np.full((r, c), NEG_INF)
    D[0, begin] = C[0, begin]
    for i in range(1, r):
        prev = D[i - 1]
        step = np.empty(c)
        step[1:] = prev[:-1]
        step[0] = prev[c - 1] if (c == 1 or (i == 1 and begin == c - 1)) else NEG_INF
        D[i] = C[i] + _lse2(prev, step)
    return D


def soft_dtw(cost, end_inds=None):
    """soft_dtw (probabilistic_dtw.py:82-121): expected edge frequencies of the 'nohor' alignment.
    cost [B,r,c] (numpy/torch fp32, ALREADY divided by the temperature); end_inds [B] ints (last valid column).
    Returns (w float32 [B,r,c], forward, backward float64 tables)."""
    cost = cost.numpy() if isinstance(cost, torch.Tensor) else np.asarray(cost)
    C = (-cost).astype(np.float64)
    B, r, c = C.shape
    end_inds = np.full(B, c - 1, dtype=np.int64) if end_inds is None else np.asarray(end_inds, dtype=np.int64)
    fwd = np.stack([gak_table(C[b], 0) for b in range(B)])
    bwd = np.stack([gak_table(C[b, ::-1, ::-1], c - int(end_inds[b]) - 1)[::-1, ::-1] for b in range(B)])
    z = fwd[np.arange(B), r - 1, end_inds][:, None, None]
    with np.errstate(invalid="ignore"):
        e = fwd + bwd - C
    e[C == NEG_INF] = NEG_INF
    w = np.exp(e - z)
    return w.astype(np.float32), fwd, bwd


def df_to_bf_order(n_nodes):
    """Row permutation of depthfirst2breadthfirst (gcp/prediction/utils/tree_utils.py:217-232): breadth-first position
    -> depth-first index."""
    depth = int(np.log2(n_nodes + 1))
    idx = np.arange(n_nodes)
    layers = []
    for _ in range(depth):
        layers.append(idx[0::2])
        idx = idx[1::2]
    return np.concatenate(list(reversed(layers)))


def binding_weights(cost, temp, end_inds):
    """AdaptiveBinding.get_w after the cost matrix (adaptive.py:50-61): soft_dtw(cost / temp, end_ind), normalised
    over the nodes (blox/torch/dist.py:22-24, eps 1e-7), rows permuted depth-first -> breadth-first.
    cost torch fp32 [B,n_nodes,T]; temp float.  Returns torch fp32 [B,n_nodes,T]."""
    scaled = cost / torch.full((1,), float(temp))
    w = torch.from_numpy(soft_dtw(scaled, end_inds)[0])
    w = w / torch.clamp(w.sum(1, keepdim=True), 1e-7)
    return w[:, torch.from_numpy(df_to_bf_order(w.shape[1]))]


# ------------------------------------------------------------------------------------------------------------------
# metric-time DTW (gcp/evaluation/dtw_utils.py)
# ------------------------------------------------------------------------------------------------------------------
def min_cumsum(C):
    """Accumulated-cost table of basic_dtw / c_dtw / cutils.min_cumsum (dtw_utils.py:77-116, cutils.pyx:21-28):
    padded (r+1) x (c+1) float64 table, D[i+1,j+1] = C[i,j] + min(D[i,j], D[i+1,j], D[i,j+1])."""
    r, c = C.shape
    D = np.zeros((r + 1, c + 1))
    D[0, 1:] = np.inf
    D[1:, 0] = np.inf
    D[1:, 1:] = C
    for i in range(r):
        for j in range(c):
            D[i + 1, j + 1] += min(D[i, j], D[i + 1, j], D[i, j + 1])
    return D


def traceback(D, j_start=None):
    """_traceback (dtw_utils.py:201-219): from (r-1, c-1) [or column j_start] back to (0,0); argmin over
    (diagonal, up, left) with the first minimum winning.  Returns int arrays p (rows), q (columns), start -> end."""
    i = D.shape[0] - 2
    j = D.shape[1] - 2 if j_start is None else int(j_start)
    p, q = [i], [j]
    while i > 0 or j > 0:
        tb = int(np.argmin((D[i, j], D[i, j + 1], D[i + 1, j])))
        if tb == 0:
            i, j = i - 1, j - 1
        elif tb == 1:
            i -= 1
        else:
            j -= 1
        p.append(i)
        q.append(j)
    return np.array(p[::-1]), np.array(q[::-1])


def basic_dtw(C):
    """basic_dtw == c_dtw (dtw_utils.py:77-116): (distance / (r+c), accumulated table [r,c], (p, q))."""
    C = np.asarray(C)
    r, c = C.shape
    D = min_cumsum(C)
    return D[-1, -1] / (r + c), D[1:, 1:], traceback(D)


def batched_dtw(C, end_ind):
    """batched_dtw (dtw_utils.py:119-130) + _batched_traceback (dtw_utils.py:222-241) per sequence: the path of
    sequence b starts at column end_ind[b].  Returns distances [B], accumulated tables [B,r,c], list of (p, q), and the
    reference's `path_lengths` (the number of path cells, except that the longest sequence(s) of the batch keep 0
    because the reference's loop ends before it records them).  The returned distance is the reference's too:
    `_batched_traceback` walks `end_ind` down to 0 IN PLACE (`j = end_ind`, dtw_utils.py:224) before line 130 reads it,
    so what the reference returns is D[b, r, 1] / (r + 1), the cost of matching every row to column 0 -- kept."""
    C = np.asarray(C)
    B, r, c = C.shape
    end_ind = np.asarray(end_ind, dtype=np.int64)
    Ds = np.stack([min_cumsum(C[b]) for b in range(B)])
    paths = [traceback(Ds[b], end_ind[b]) for b in range(B)]
    n = np.array([len(p) for p, _ in paths])
    lengths = np.where(n < n.max(), n, 0)
    dist = Ds[:, -1, 1] / (r + 1)
    return dist, Ds[:, 1:, 1:], paths, lengths


def stack_batched_paths(paths):
    """The [steps, B] arrays `_batched_traceback` returns: every sequence's path right-aligned, shorter ones padded in
    front with (0, 0) (the finished sequences keep re-emitting their origin)."""
    n = max(len(p) for p, _ in paths)
    P = np.zeros((n, len(paths)), dtype=np.int64)
    Q = np.zeros((n, len(paths)), dtype=np.int64)
    for b, (p, q) in enumerate(paths):
        P[n - len(p):, b] = p
        Q[n - len(q):, b] = q
    return P, Q


def single_matches(estimates, targets):
    """DTWEvalBinding.get_single_matches (gcp/evaluation/evaluation_matching.py:135-147): cost = cdist(estimates,
    targets, 'mean'), DTW, then for every target frame the estimate on the path with the smallest ACCUMULATED cost.
    estimates [n,...], targets [m,...] torch fp32.  Returns (inds [m], (p, q), accumulated table, distance)."""
    matrix = batch_cdist_mean(estimates[None], targets[None])[0].numpy()
    d, acc, (p, q) = basic_dtw(matrix)
    match = np.full_like(acc, np.inf)
    match[p, q] = acc[p, q]
    return np.argmin(match, axis=0), (p, q), acc, d
