"""GPU parity at the second tree shape: the 9-room planner model (experiments/control/9room/gcp_tree/mod_hyper.py:33-54 --
hierarchy_levels 7 = 127 nodes, max_seq_len 100, ONE TreeModule for all levels).  Tree depth, sequence length and tied /
untied layers are context parameters of the library (gcpb200_config.hierarchy_levels / max_seq_len / tied_layers); fixture
tests/golden/tree9room.npz was produced by the unmodified reference (oracle/make_golden_9room.py).  Tolerances as in
tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O

pytestmark = pytest.mark.gpu


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
def g9(golden_dir):
    return np.load(os.path.join(golden_dir, "tree9room.npz"))


@pytest.fixture(scope="module")
def sd9(g9):
    from video_gcp_b200 import hparams
    from video_gcp_b200.synthetic import synthetic_state_dict
    return synthetic_state_dict(hparams.build_hparams(hparams.gcp_tree_9room_config(batch_size=1)), int(g9["weight_seed"]))


def _inputs(B, seed):
    r = np.random.default_rng([int(seed), 999])
    I_0 = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    I_g = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    z = r.standard_normal(size=(B, 127, 256)).astype(np.float32)
    return torch.from_numpy(I_0), torch.from_numpy(I_g), torch.from_numpy(z)


@pytest.fixture(scope="module")
def eng9(dev, sd9):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=256, hierarchy_levels=7, max_seq_len=100, tied_layers=True)
    eng.load_weights(sd9)
    yield eng
    eng.close()


def test_9room_forward_against_reference_fixture(eng9, dev, g9):
    I_0, I_g, z = _inputs(2, int(g9["input_seed"]))
    out = eng9.rollout(I_0.to(dev), I_g.to(dev), z.to(dev), end_ind=torch.as_tensor(g9["end_ind"]).to(dev), want_prior=True)
    torch.cuda.synchronize()
    assert out["e_df"].shape == (2, 127, 128) and out["seq_len_logits"].shape == (2, 100) and out["actions"].shape == (2, 100, 2)
    assert rel(out["e_0"], g9["e0"]) < 1e-4
    assert rel(out["seq_len_logits"], g9["seq_len_logits"]) < 2.5e-2
    assert rel(out["e_df"], g9["e_df"]) < 2.5e-2
    assert rel(out["mu_df"], g9["mu_df"]) < 4e-2
    assert maxabs(out["images_df"][:, g9["img_nodes"].tolist()], g9["images_sel"]) < 5e-3
    assert maxabs(out["images_df"], g9["images_f16"].astype(np.float32)) < 6e-3
    assert rel(out["existence"], g9["existence"]) < 6e-2
    lmax = int(g9["end_ind"].max()) + 1
    assert rel(out["model_enc_seq"][:, :lmax], g9["model_enc_seq"]) < 2.5e-2
    assert rel(out["actions"][:, :lmax - 1], g9["actions"]) < 5e-2
    assert rel(out["regressed_state"][:, :lmax], g9["regressed_state"]) < 5e-2
    pruned = eng9.prune_gather(out["images_df"], out["end_ind"])
    L0 = int(g9["pruned_len"][0])
    assert maxabs(pruned[0, :L0].reshape(L0, 3, 32, 32), g9["pruned0"]) < 5e-3
    assert float(pruned[0, L0:].abs().max()) == 0.0


def test_9room_modes_agree_bitwise(eng9, dev):
    """Host-noise upload (level sets derived from the depth), planner mode (kept nodes only, fused L2 cost) and the full
    decode give the same bits at depth 7 too; 150 candidates (two 128-row tiles), every length class."""
    B = 150
    I_0, I_g, z = _inputs(B, 8)
    I0, Ig = I_0[:1].to(dev), I_g[:1].to(dev)
    end = torch.as_tensor(np.random.default_rng(3).integers(1, 100, size=B))
    end[:4] = torch.tensor([1, 2, 99, 63])
    kw = dict(end_ind=end.to(dev), images_shared=True, fresh=True)
    full = eng9.rollout(I0, Ig, z.to(dev), l2_goal=Ig[0], **kw)
    host = eng9.rollout(I0, Ig, z.pin_memory(), l2_goal=Ig[0], **kw)
    kept = eng9.rollout(I0, Ig, z.to(dev), decode_kept_only=True, l2_goal=Ig[0], **kw)
    for k in ("e_df", "images_df", "actions", "existence", "l2_cost"):
        assert torch.equal(full[k], host[k]), k
    assert torch.equal(kept["l2_cost"], full["l2_cost"]) and torch.equal(kept["e_df"], full["e_df"])
    assert torch.equal(eng9.prune_gather(kept["images_df"], kept["end_ind"]), eng9.prune_gather(full["images_df"], full["end_ind"]))
    c = eng9.cost_l2(full["images_df"], full["end_ind"], Ig[0], True, 1.0)
    assert float(((c - full["l2_cost"]).abs() / c.abs()).max()) < 2e-6


def test_9room_planner_against_oracle(dev, sd9, g9):
    """TreeModel built from the 9-room config behind the flat CEM planner: costs / elites / refit of one iteration on the
    fixture's 8 candidates against the reference's own numbers."""
    from functools import partial
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
    model = TreeModel(hparams.gcp_tree_9room_config(batch_size=1), None, max_candidates=128)
    model.load_state_dict(sd9, strict=True)
    model.device = dev
    model.eval()
    N = 8
    r = np.random.default_rng(int(g9["rng_seed"]))
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 0.3, size=(N, 127, 256)).astype(np.float32)
    pl = ImageCEMPlanner(dict(batch_size=N, n_iters=1, elite_frac=0.25, cost_fcn=L2ImageCost, dense_cost=True,
                              final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=7),
                              max_seq_len=100, action_dim=256, initial_std=0.3), GCPImageSimulator(model, append_latent=True))
    pl._sampler.init()
    model.inject_end_ind = torch.as_tensor(g9["cem_end_ind"])
    cost, idx, val, packed = pl.cem_iteration(state, goal, samples=torch.from_numpy(samples).to(dev))
    torch.cuda.synchronize()
    np.testing.assert_allclose(cost.cpu().numpy(), g9["l2_dense"], rtol=2e-3)
    assert idx.cpu().tolist() == g9["elite_idx"].tolist()
    d = pl._sampler.get_dists()
    assert np.abs(d.mean - g9["fit_mean"]).max() < 1e-6 and np.abs(d.std - g9["fit_std"]).max() < 1e-6
    # reference contract of the simulator at this shape
    ro = pl._simulator.rollout(state, goal, samples, 100)
    assert [p.shape[0] for p in ro.predictions] == g9["pred_len"].tolist()
    assert np.abs(ro.predictions[2][:, :3072] - g9["pred2"][:, :3072]).max() < 5e-3
    assert rel(ro.latents[2], g9["lat2"]) < 2.5e-2 and rel(ro.actions[2], g9["act2"]) < 5e-2
    model.engine.close()
