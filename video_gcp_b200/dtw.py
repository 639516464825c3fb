"""DTW family on the B200 behind the reference's own entry points.

Same names, argument meaning and return values as the reference functions they replace:
  * `batch_cdist(x1, x2, reduction='mean')`        blox/torch/ops.py:62-91
  * `soft_dtw(C, end_inds=None)`                   gcp/prediction/models/adaptive_binding/probabilistic_dtw.py:82-121
  * `get_w(cost_matrix, temp, end_ind)`            AdaptiveBinding.get_w after the cost matrix, adaptive.py:50-61
  * `c_dtw(C)` / `basic_dtw(C)`, `batched_dtw(C, end_ind)`        gcp/evaluation/dtw_utils.py:77-130
  * `DTWEvalBinding.get_single_matches(targets, estimates)`       gcp/evaluation/evaluation_matching.py:135-147
All arithmetic runs in libgcpb200.so (csrc/dtw_kernels.cuh) through `Engine`; there is no CPU fallback.  The numpy
entry points (`c_dtw`, `batched_dtw`) take and return numpy arrays like the reference; the torch ones keep tensors on
the device.
"""
import numpy as np
import torch

from . import _C
from .types import AttrDict

_ENGINES = {}


def get_engine(device=None):
    """A weight-less context per device for the DTW kernels (they need no model)."""
    from .engine import Engine
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise _C.GcpB200Error("the DTW kernels need a CUDA device (B200, sm_100a); there is no CPU path")
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _ENGINES:
        _ENGINES[key] = Engine(torch.device("cuda", key), max_candidates=128)
    return _ENGINES[key]


def _dev(t):
    return t.device if (isinstance(t, torch.Tensor) and t.is_cuda) else None


def batch_cdist(x1, x2, reduction='mean', engine=None):
    if reduction != 'mean':
        raise NotImplementedError("only reduction='mean' is on the DTW path (adaptive.py:44-47, evaluation_matching.py:138)")
    eng = engine or get_engine(_dev(x1))
    return eng.cdist_mean(x1, x2)


def cdist(x1, x2, reduction='mean', engine=None):
    return batch_cdist(x1[None], x2[None], reduction, engine)[0]


def soft_dtw(C, end_inds=None, temp=1.0, engine=None, check=True):
    """Expected edge frequencies [B,r,c] (float32, on the device).  `C` is the cost matrix; the reference's caller
    divides it by the temperature first (adaptive.py:51) -- pass `temp` to fuse that division into the kernel.
    Where the reference prints a warning and drops into pdb (row sums not within 1e-2 of 1), this raises."""
    eng = engine or get_engine(_dev(C))
    out = eng.soft_dtw(C, temp, end_inds)
    if check:
        m = float(out["rowsum_max"])
        if not (1 - 1e-2 < m < 1 + 1e-2):
            raise FloatingPointError("dtw is not stable with these cost values (max row sum %g)" % m)
    return out["w"]


def get_w(cost_matrix, temp, end_ind, engine=None):
    """AdaptiveBinding.get_w after the cost matrix (adaptive.py:50-61): soft-DTW posterior over (node, frame) edges,
    normalised over the nodes, nodes in breadth-first order.  [B,n_nodes,T] float32 on the device."""
    eng = engine or get_engine(_dev(cost_matrix))
    out = eng.soft_dtw(cost_matrix, float(temp), end_ind, want_bf=True)
    m = float(out["rowsum_max"])
    if not (1 - 1e-2 < m < 1 + 1e-2):
        raise FloatingPointError("dtw is not stable with these cost values (max row sum %g)" % m)
    return out["w_bf"]


def _paths(out, b):
    n = int(out["path_len"][b])
    return out["path_p"][b, -n:].cpu().numpy().astype(np.int64), out["path_q"][b, -n:].cpu().numpy().astype(np.int64)


def c_dtw(C, engine=None):
    """(distance / (r + c), accumulated cost matrix [r,c], (p, q)) like the reference's c_dtw / basic_dtw."""
    C = np.asarray(C)
    eng = engine or get_engine()
    t = torch.from_numpy(np.ascontiguousarray(C if C.dtype == np.float64 else C.astype(np.float32)))[None]
    out = eng.dtw(t, want_matches=False)
    return float(out["dist"][0]), out["acc"][0, 1:, 1:].cpu().numpy(), _paths(out, 0)


basic_dtw = c_dtw


def batched_dtw(C, end_ind, engine=None):
    """batched_dtw (dtw_utils.py:119-130).  Returns what the reference returns, including its two quirks: the distance is
    read after `_batched_traceback` has walked `end_ind` down to zero in place (so it is acc[:, -1, 0] / (r + 1), and the
    caller's `end_ind` array comes back zeroed), and `path_lengths` stays 0 for the longest sequence(s) of the batch."""
    C = np.asarray(C)
    eng = engine or get_engine()
    B, r, c = C.shape
    out = eng.dtw(torch.from_numpy(np.ascontiguousarray(C.astype(np.float64))), end_ind=np.asarray(end_ind))
    n = out["path_len"].cpu().numpy().astype(np.int64)
    steps = int(n.max())
    P = out["path_p"][:, -steps:].t().cpu().numpy().astype(np.int64)
    Q = out["path_q"][:, -steps:].t().cpu().numpy().astype(np.int64)
    acc = out["acc"][:, 1:, 1:].cpu().numpy()
    if isinstance(end_ind, np.ndarray):
        end_ind[...] = 0
    return acc[:, -1, 0] / (r + 1), acc, (P, Q), np.where(n < steps, n, 0)


class DTWEvalBinding:
    """Matches predicted frames to ground-truth frames by DTW (gcp/evaluation/evaluation_matching.py:124-147)."""

    def __init__(self, hp=None):
        self._hp = hp

    @staticmethod
    def get_single_matches(targets, estimates, engine=None):
        eng = engine or get_engine(_dev(estimates) or _dev(targets))
        estimates = estimates.to(eng.device)
        targets = targets.to(eng.device)
        matrix = eng.cdist_mean(estimates[None], targets[None])
        out = eng.dtw(matrix)
        gen_images = eng.gather_rows(estimates, out["match_inds"][0])
        path = _paths(out, 0)
        return gen_images, AttrDict(targets=targets, estimates=estimates, matching_path=path, gen_images=gen_images)
