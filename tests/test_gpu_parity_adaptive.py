"""GPU parity tests of the adaptive-binding GCP-tree rollout (config 4); run with -m gpu on the B200 box.

Tolerances (bf16 tensor-core operands, fp32 accumulation): node latents as the balanced tree (2.5e-2 relative);
pixel-copy frames max-abs ADA_IMG_TOL (frames in [-1,1]); distance-predictor logits ADA_DIST_TOL relative; the keep
mask (integer work) is identical wherever |logit - logit(threshold)| exceeds the logit tolerance, and the compaction /
gather of kept nodes is bit-exact given the mask.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200 import hparams
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu

ADA_LAT_TOL, ADA_IMG_TOL, ADA_DIST_TOL = 2.5e-2, 5e-3, 6e-2


def rel(got, ref):
    got = got.detach().double().cpu() if isinstance(got, torch.Tensor) else torch.as_tensor(np.asarray(got)).double()
    ref = ref.detach().double().cpu() if isinstance(ref, torch.Tensor) else torch.as_tensor(np.asarray(ref)).double()
    assert not torch.isnan(got).any()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def maxabs(got, ref):
    return float((got.detach().double().cpu() - torch.as_tensor(np.asarray(ref)).double()).abs().max())


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ada_sd():
    hp = hparams.build_hparams(hparams.gcp_adaptive_25room_config(batch_size=1))
    return synthetic_state_dict(hp, 3)


@pytest.fixture(scope="module")
def ada_engine(dev, ada_sd):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=256, model="tree_adaptive")
    eng.load_weights(ada_sd)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def case(ada_engine, dev, ada_sd):
    B = 5
    inp = synthetic_rollout_inputs(B, seed=23, shared_images=False)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = O.adaptive_rollout(ada_sd, inp["I_0"], inp["I_g"], inp["z"])
    out = ada_engine.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev))
    torch.cuda.synchronize()
    return inp, ref, out


def _check_keep(out, dist_ref, keep_ref, b):
    """Kept-node list of candidate b equals the oracle's wherever the logits are not within tolerance of 0."""
    tol = ADA_DIST_TOL * float(np.abs(dist_ref).max())
    n = int(out["pruned_len"][b])
    got = np.zeros(255, dtype=bool)
    got[out["pruned_nodes"][b, :n].cpu().numpy()] = True
    nodes = out["pruned_nodes"][b, :n].cpu().numpy()
    assert (np.diff(nodes) > 0).all() and nodes[0] == 0              # in order, node 0 always kept
    sure = np.concatenate([[True], np.abs(dist_ref[b]) > tol])
    assert (got[sure] == keep_ref[b][sure]).all()
    # and the list is exactly what the DEVICE logits imply (integer compaction is exact)
    own = np.concatenate([[True], ~(out["distances"][b].cpu().numpy() > 0.0)])
    assert (got == own).all()


def test_adaptive_latents_frames_distances(case):
    _, ref, out = case
    assert rel(out["e_df"], ref["tree"]["e"]) < ADA_LAT_TOL
    assert maxabs(out["images_df"], ref["images_df"]) < ADA_IMG_TOL
    assert rel(out["distances"], ref["distances"]) < ADA_DIST_TOL
    for b in range(out["e_df"].shape[0]):
        _check_keep(out, ref["distances"].numpy(), ref["keep"].numpy(), b)


def test_adaptive_threshold_and_gather(ada_engine, case, dev):
    """Other thresholds move the cut; gather_nodes / cost_l2_nodes follow the node list bit-exactly."""
    inp, ref, out = case
    B = out["e_df"].shape[0]
    d = out["distances"].cpu().numpy()
    for thr in (0.3, 0.5, 0.62):
        o = ada_engine.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev),
                               prune_threshold=thr, fresh=True)
        lg = float(np.log(thr / (1 - thr)))
        for b in range(B):
            want = np.nonzero(np.concatenate([[True], ~(d[b] > np.float32(lg))]))[0]
            n = int(o["pruned_len"][b])
            assert n == len(want) and o["pruned_nodes"][b, :n].cpu().numpy().tolist() == want.tolist(), (thr, b)
    img = ada_engine.gather_nodes(out["images_df"], out["pruned_nodes"], out["pruned_len"]).cpu()
    goal = inp["I_g"][0]
    cost = ada_engine.cost_l2_nodes(out["images_df"], out["pruned_nodes"], out["pruned_len"], goal.to(dev), True, 2.0).cpu()
    for b in range(B):
        n = int(out["pruned_len"][b])
        nodes = out["pruned_nodes"][b, :n].cpu().long()
        want = out["images_df"][b].cpu()[nodes].reshape(n, -1)
        assert torch.equal(img[b, :n], want) and float(img[b, n:].abs().sum()) == 0.0
        c = ((want.reshape(n, 3, 32, 32) - goal) ** 2).sum((1, 2, 3)).sqrt()
        c[-1] *= 2.0
        assert abs(float(cost[b]) - float(c.sum())) < 1e-4 * float(c.sum())


def test_adaptive_golden_fixture_from_reference(ada_engine, dev, golden_dir):
    """Against outputs of the UNMODIFIED reference (tests/golden/adaptive_forward_B2.npz)."""
    g = np.load(os.path.join(golden_dir, "adaptive_forward_B2.npz"))
    inp = synthetic_rollout_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    out = ada_engine.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev))
    assert rel(out["e_df"], g["e_df"]) < ADA_LAT_TOL
    assert maxabs(out["images_df"][:, g["img_nodes"].tolist()], g["images_sel"]) < ADA_IMG_TOL
    assert maxabs(out["images_df"], g["images_f16"].astype(np.float32)) < ADA_IMG_TOL + 1e-3
    assert rel(out["distances"], g["distances"]) < ADA_DIST_TOL
    assert rel(out["seq_len_logits"], g["seq_len_logits"]) < 2.5e-2
    keep_ref = np.concatenate([np.ones((2, 1), bool), ~(g["distances"] > 0)], 1)
    for b in range(2):
        _check_keep(out, g["distances"], keep_ref, b)


def test_adaptive_tc_kernels_match_simt_verification_kernels(dev, ada_sd):
    from video_gcp_b200.engine import Engine
    inp = synthetic_rollout_inputs(4, seed=37, shared_images=True)
    outs = []
    from verify_lib import verify_engine
    for use_ref in (True, False):
        eng = verify_engine(dev, max_candidates=128, model="tree_adaptive") if use_ref else \
            Engine(dev, max_candidates=128, model="tree_adaptive")
        eng.load_weights(ada_sd)
        outs.append(eng.rollout(inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), images_shared=True, fresh=True))
        torch.cuda.synchronize()
        eng.close()
    a, b = outs
    assert rel(b["e_df"], a["e_df"]) < 1.5e-2
    assert maxabs(b["images_df"], a["images_df"].cpu().numpy()) < 3e-3


def test_adaptive_model_drop_in(dev, ada_sd):
    """Reference-facing API: TreeModel built from the adaptive config; outputs.pruned_prediction / distance_predictor."""
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.types import AttrDict
    model = TreeModel(hparams.gcp_adaptive_25room_config(batch_size=1), None, max_candidates=128)
    model.load_state_dict(ada_sd, strict=True)
    model.to(dev)
    model.device = dev
    model.eval()
    inp = synthetic_rollout_inputs(3, seed=41, shared_images=False)
    inputs = AttrDict(I_0=inp["I_0"].to(dev), I_g=inp["I_g"].to(dev), z=inp["z"].to(dev)[..., None, None],
                      start_ind=torch.zeros(3, dtype=torch.long, device=dev),
                      end_ind=torch.full((3,), 199, dtype=torch.long, device=dev))
    with model.val_mode():
        out = model(inputs)
    with torch.no_grad():
        ref = O.adaptive_rollout(ada_sd, inp["I_0"], inp["I_g"], inp["z"])
    assert rel(out.distance_predictor.distances, ref["distances"]) < ADA_DIST_TOL
    assert maxabs(out.tree.df.images, ref["images_df"]) < ADA_IMG_TOL
    tol = ADA_DIST_TOL * float(ref["distances"].abs().max())
    for b in range(3):
        got = out.pruned_prediction[b]
        if bool((ref["distances"][b].abs() > tol).all()):
            assert got.shape == ref["pruned_images"][b].shape
            assert maxabs(got, ref["pruned_images"][b]) < ADA_IMG_TOL
