"""Golden fixtures for the DTW family, produced by RUNNING THE UNMODIFIED REFERENCE:
    python -m oracle.make_golden_dtw        (build container only; /root/reference must exist)

TEST INFRASTRUCTURE.  Calls, unchanged: `soft_dtw` / `fast_gak` (gcp/prediction/models/adaptive_binding/
probabilistic_dtw.py), the post-processing of `AdaptiveBinding.get_w` (adaptive.py:41-61: batch_cdist 'mean', division
by the temperature, soft_dtw, normalize over nodes, depthfirst2breadthfirst), `basic_dtw` / `batched_dtw`
(gcp/evaluation/dtw_utils.py) and `DTWEvalBinding.get_single_matches` (gcp/evaluation/evaluation_matching.py:135-147).
Small inputs are stored next to the outputs; the full-size case stores its seed (`full_cost` below regenerates it).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox.torch.dist import normalize  # noqa: E402
from blox.torch.ops import batch_cdist  # noqa: E402
from gcp.evaluation.dtw_utils import basic_dtw, batched_dtw  # noqa: E402
from gcp.evaluation.evaluation_matching import DTWEvalBinding  # noqa: E402
from gcp.prediction.models.adaptive_binding.probabilistic_dtw import fast_gak, soft_dtw  # noqa: E402
from gcp.prediction.utils.tree_utils import depthfirst2breadthfirst  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FULL_SEED, FULL_END = 404, (199, 57)


def full_cost(seed=FULL_SEED, B=2, r=255, c=200):
    """Full-size cost matrix of the 25-room shape: 255 tree nodes x 200 frames, mean squared image distances."""
    return np.random.default_rng(seed).uniform(0.0, 1.3, size=(B, r, c)).astype(np.float32)


def main():
    rng = np.random.default_rng(7)
    out = {}
    # ---- soft-DTW, small ragged batch (depth-4 tree: 15 nodes, 8 frames), fp32 costs already divided by the temperature
    cost = rng.uniform(0, 2.0, size=(3, 15, 8)).astype(np.float32)
    ends = np.array([7, 3, 5])
    out["soft_cost"], out["soft_end"] = cost, ends
    out["soft_w"] = soft_dtw(torch.from_numpy(cost), torch.from_numpy(ends)).numpy()
    out["soft_w_noend"] = soft_dtw(torch.from_numpy(cost)).numpy()
    C = torch.from_numpy(-cost).double()
    out["soft_fwd"] = fast_gak(C, 'nohor', torch.zeros(3, dtype=torch.long)).numpy()
    out["soft_bwd_flipped"] = fast_gak(torch.flip(C, [-1, -2]), 'nohor', torch.from_numpy(8 - ends - 1)).numpy()
    # square case r == c: only the diagonal alignment exists
    sq = rng.uniform(0, 2.0, size=(2, 7, 7)).astype(np.float32)
    out["sq_cost"] = sq
    out["sq_w"] = soft_dtw(torch.from_numpy(sq)).numpy()
    # the reference's column wrap-around (j-1 = -1) is visible when the recursion begins in the last column
    wrap = rng.uniform(0, 2.0, size=(2, 5, 3))
    out["wrap_C"] = -wrap
    out["wrap_begin"] = np.array([2, 1])
    out["wrap_D"] = fast_gak(torch.from_numpy(-wrap), 'nohor', torch.tensor([2, 1])).numpy()
    one = rng.uniform(0, 2.0, size=(1, 4, 1))
    out["onecol_C"] = -one
    out["onecol_D"] = fast_gak(torch.from_numpy(-one), 'nohor', torch.tensor([0])).numpy()
    # ---- soft-DTW at the 25-room size (255 nodes x 200 frames); stored as marginals + every 16th row
    fc = full_cost()
    w = soft_dtw(torch.from_numpy(fc), torch.tensor(FULL_END)).numpy()
    out["full_seed"], out["full_end"] = np.int64(FULL_SEED), np.array(FULL_END)
    out["full_w_rows"] = w[:, ::16]
    out["full_w_sum_nodes"] = w.astype(np.float64).sum(1)
    out["full_w_sum_frames"] = w.astype(np.float64).sum(2)
    out["full_argmax_frame"] = w.argmax(2)
    # ---- AdaptiveBinding.get_w after the tree: images -> cost -> soft-DTW -> normalise -> breadth-first
    imgs = rng.uniform(-1, 1, size=(2, 15, 3, 8, 8)).astype(np.float32)
    traj = rng.uniform(-1, 1, size=(2, 8, 3, 8, 8)).astype(np.float32)
    gend = np.array([7, 4])
    temp = torch.ones(1) * 0.5
    cm = batch_cdist(torch.from_numpy(imgs), torch.from_numpy(traj), reduction='mean')
    gw = depthfirst2breadthfirst(normalize(soft_dtw(cm / temp, torch.from_numpy(gend)), 1))
    out.update(getw_imgs=imgs, getw_traj=traj, getw_end=gend, getw_temp=np.float32(0.5), getw_cost=cm.numpy(), getw_w=gw.numpy())
    # ---- metric-time DTW
    dc = rng.uniform(0, 1, size=(37, 29)).astype(np.float32)
    d, acc, path = basic_dtw(dc)
    out.update(dtw_cost=dc, dtw_dist=np.float64(d), dtw_acc=acc, dtw_p=np.asarray(path[0]), dtw_q=np.asarray(path[1]))
    tie = np.round(rng.uniform(0, 3, size=(9, 12))).astype(np.float32)             # many exact ties: first-minimum rule
    d, acc, path = basic_dtw(tie)
    out.update(tie_cost=tie, tie_dist=np.float64(d), tie_acc=acc, tie_p=np.asarray(path[0]), tie_q=np.asarray(path[1]))
    bc = rng.uniform(0, 1, size=(3, 12, 9)).astype(np.float32)
    bend = np.array([8, 4, 6])
    dist, acc, (P, Q), lengths = batched_dtw(bc.astype(np.float64), bend.copy())
    out.update(bat_cost=bc, bat_end=bend, bat_dist=dist, bat_acc=acc, bat_P=P, bat_Q=Q, bat_len=lengths)
    # ---- DTWEvalBinding.get_single_matches
    est = rng.uniform(-1, 1, size=(11, 3, 8, 8)).astype(np.float32)
    tgt = rng.uniform(-1, 1, size=(7, 3, 8, 8)).astype(np.float32)
    tgt[2] = est[5]
    gen, mo = DTWEvalBinding.get_single_matches(torch.from_numpy(tgt), torch.from_numpy(est))
    inds = np.array([int(np.argmin(((est - g.numpy()[None]) ** 2).reshape(11, -1).sum(1))) for g in gen])
    out.update(match_est=est, match_tgt=tgt, match_inds=inds, match_p=np.asarray(mo.matching_path[0]),
               match_q=np.asarray(mo.matching_path[1]))
    np.savez_compressed(os.path.join(GOLDEN, "dtw_family.npz"), **out)
    print("wrote dtw_family.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
