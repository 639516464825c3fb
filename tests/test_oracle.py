"""Pins oracle/gcp_oracle.py against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200.synthetic import synthetic_rollout_inputs


@pytest.fixture(scope="module")
def case_a(golden_dir, sd):
    g = np.load(os.path.join(golden_dir, "tree_forward_B2.npz"))
    inp = synthetic_rollout_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = O.rollout(sd, inp["I_0"], inp["I_g"], inp["z"], g["end_ind"])
    return g, out


def _close(a, b, atol, rtol=0.0):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, atol=atol, rtol=rtol)


def test_encoder_and_length(case_a):
    g, out = case_a
    _close(out["e0"], g["e0"], 1e-5)
    _close(out["eg"], g["eg"], 1e-5)
    _close(out["seq_len_logits"], g["seq_len_logits"], 2e-5)


def test_tree_latents(case_a):
    g, out = case_a
    _close(out["tree"]["e"], g["e_df"], 2e-5)
    _close(out["tree"]["mu"], g["mu_df"], 2e-5)
    _close(out["tree"]["log_sigma"], g["log_sigma_df"], 2e-5)
    _close(out["tree"]["hidden"][:, g["hid_nodes"].tolist()], g["hidden_sel"], 2e-5)


def test_decoder_images(case_a):
    g, out = case_a
    img = out["images_df"]
    _close(img[:, g["img_nodes"].tolist()], g["images_sel"], 2e-5)
    _close(img.double().sum((2, 3, 4)), g["images_sum"], 2e-2)
    _close(img.double().abs().sum((2, 3, 4)), g["images_abs"], 2e-2)
    _close(img, g["images_f16"].astype(np.float32), 1e-3)


def test_aux_heads(case_a):
    g, out = case_a
    _close(out["existence"], g["existence"], 2e-5)
    _close(out["model_enc_seq"], g["model_enc_seq"], 2e-5)
    _close(out["actions"], g["actions"], 2e-5)
    _close(out["regressed_state"], g["regressed_state"], 2e-5)
    assert [len(p) for p in out["pruned_images"]] == g["pruned_len"].tolist()
    _close(out["pruned_images"][0], g["pruned0"], 2e-5)


def test_balanced_pruning_all_lengths(golden_dir):
    g = np.load(os.path.join(golden_dir, "balanced_pruning.npz"))
    for e in range(1, 200):
        keep, t = O.balanced_keep_mask(e)
        assert (keep == g["keep"][e]).all(), e
        assert (t == g["timesteps"][e]).all(), e
        idx = O.prune_indices(e)
        assert len(idx) == e + 1
        assert (t[idx] == np.arange(e + 1)).all()      # kept nodes are frames 0..end_ind in order


def test_simulator_costs_elites_refit(golden_dir, sd):
    g = np.load(os.path.join(golden_dir, "cem_N12.npz"))
    N = 12
    r = np.random.default_rng(int(g["rng_seed"]))
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 0.3, size=(N, 255, 256))
    end = r.integers(2, 200, size=N)
    assert (end == g["end_ind"]).all() and np.array_equal(state, g["state"])
    with torch.no_grad():
        ro = O.simulator_rollout(sd, state, goal, samples, end)
    assert [p.shape[0] for p in ro["predictions"]] == g["pred_len"].tolist()
    _close(ro["predictions"][3], g["pred3"], 3e-5)
    _close(ro["actions"][3], g["act3"], 3e-5)
    _close(ro["states"][3], g["state3"], 3e-5)
    _close(ro["latents"][3], g["lat3"], 3e-5)
    np.testing.assert_allclose([p.astype(np.float64).sum() for p in ro["predictions"]], g["pred_sum"], atol=5e-2)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]]
    lats = [p[:, 3072:] for p in ro["predictions"]]
    l2 = O.l2_image_cost(imgs, goal, True, 1.0)
    np.testing.assert_allclose(l2, g["l2_dense"], rtol=1e-5)
    np.testing.assert_allclose(O.l2_image_cost(imgs, goal, False, 2.0), g["l2_last_w2"], rtol=1e-5)
    with torch.no_grad():
        learned = O.image_wrapped_learned_cost(sd, lats)
    np.testing.assert_allclose(learned, g["learned"], rtol=2e-5, atol=1e-4)
    el = O.elites(l2, N, 0.25)
    assert el.tolist() == g["elite_idx"].tolist()
    mean, std = O.refit(samples, el)
    np.testing.assert_allclose(mean, g["fit_mean"], atol=1e-12)
    np.testing.assert_allclose(std, g["fit_std"], atol=1e-12)


# ---- second tree shape: the 9-room planner model (7 levels, 100 frames, ONE TreeModule for all levels) ----------------------
def _inputs_9room(B, seed, n_nodes=127):
    r = np.random.default_rng([int(seed), 999])
    I_0 = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    I_g = r.uniform(-1, 1, size=(B, 3, 32, 32)).astype(np.float32)
    z = r.standard_normal(size=(B, n_nodes, 256)).astype(np.float32)
    return torch.from_numpy(I_0), torch.from_numpy(I_g), torch.from_numpy(z)


def test_9room_tied_depth7_against_reference(golden_dir):
    """oracle/make_golden_9room.py ran the unmodified reference with experiments/control/9room/gcp_tree/mod_hyper.py's model
    (hierarchy_levels 7, max_seq_len 100, untied_layers False): the oracle reads the depth off z and the tied state-dict
    layout off the keys."""
    from video_gcp_b200 import hparams
    from video_gcp_b200.synthetic import synthetic_state_dict
    g = np.load(os.path.join(golden_dir, "tree9room.npz"))
    hp9 = hparams.build_hparams(hparams.gcp_tree_9room_config(batch_size=1))
    sd9 = synthetic_state_dict(hp9, int(g["weight_seed"]))
    assert "tree_module.prior.input.conv.weight" in sd9 and not any(k.startswith("tree_module.tree_modules.") for k in sd9)
    I_0, I_g, z = _inputs_9room(2, int(g["input_seed"]))
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = O.rollout(sd9, I_0, I_g, z, g["end_ind"])
    _close(out["e0"], g["e0"], 1e-5)
    _close(out["seq_len_logits"], g["seq_len_logits"], 2e-5)
    assert out["seq_len_logits"].shape[1] == 100 and out["tree"]["e"].shape[1] == 127
    _close(out["tree"]["e"], g["e_df"], 2e-5)
    _close(out["tree"]["mu"], g["mu_df"], 2e-5)
    _close(out["images_df"][:, g["img_nodes"].tolist()], g["images_sel"], 2e-5)
    _close(out["images_df"], g["images_f16"].astype(np.float32), 1e-3)
    _close(out["existence"], g["existence"], 2e-5)
    _close(out["model_enc_seq"], g["model_enc_seq"], 2e-5)
    _close(out["actions"], g["actions"], 2e-5)
    _close(out["regressed_state"], g["regressed_state"], 2e-5)
    assert [len(p) for p in out["pruned_images"]] == g["pruned_len"].tolist()
    _close(out["pruned_images"][0], g["pruned0"], 2e-5)
    # simulator + cost + elites + refit
    N = 8
    r = np.random.default_rng(int(g["rng_seed"]))
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 0.3, size=(N, 127, 256))
    end = r.integers(2, 100, size=N)
    assert (end == g["cem_end_ind"]).all()
    with torch.no_grad():
        ro = O.simulator_rollout(sd9, state, goal, samples, end)
    assert [p.shape[0] for p in ro["predictions"]] == g["pred_len"].tolist()
    _close(ro["predictions"][2], g["pred2"], 3e-5)
    _close(ro["actions"][2], g["act2"], 3e-5)
    _close(ro["latents"][2], g["lat2"], 3e-5)
    l2 = O.l2_image_cost([p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]], goal, True, 1.0)
    np.testing.assert_allclose(l2, g["l2_dense"], rtol=1e-5)
    el = O.elites(l2, N, 0.25)
    assert el.tolist() == g["elite_idx"].tolist()
    mean, std = O.refit(samples, el)
    np.testing.assert_allclose(mean, g["fit_mean"], atol=1e-12)
    np.testing.assert_allclose(std, g["fit_std"], atol=1e-12)
