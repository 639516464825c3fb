"""Pins oracle/train_oracle.py (training-phase forward + loss, SURVEY 8(f)-2 / BASELINE config 1) against fixtures
produced by the unmodified reference (oracle/make_golden_train.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as TO
from oracle.gcp_oracle import df_index
from video_gcp_b200.synthetic import synthetic_train_batch


def _aux(g):
    return dict(inv_t0=g["inv_t0"], inv_t1=g["inv_t1"], cost_start=g["cost_start"], cost_end=g["cost_end"],
                cost_target=g["cost_target"])


def _bf_to_df():
    """breadth-first node order (level by level) -> depth-first index"""
    return np.array([df_index(l, j) for l in range(8) for j in range(2 ** l)])


@pytest.fixture(scope="module")
def case_a(golden_dir, sd):
    g = np.load(os.path.join(golden_dir, "train_forward_B2.npz"))
    batch = synthetic_train_batch(2, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = TO.forward_loss(sd, batch, _aux(g))
    return g, out


def _close(a, b, atol, rtol=0.0):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, atol=atol, rtol=rtol)


def test_batch_stat_encoders(case_a):
    g, out = case_a
    _close(out["e0"], g["e0"], 2e-5)
    _close(out["eg"], g["eg"], 2e-5)
    _close(out["s0"], g["skip0"], 1e-5)
    _close(out["s2"], g["skip2"], 2e-5)
    _close(out["enc_traj_seq"], g["enc_traj_seq"], 2e-5)
    _close(out["inf_enc_seq"], g["inf_enc_seq"], 5e-5)
    _close(out["seq_len_logits"], g["seq_len_logits"], 5e-5)


def test_matching_is_bit_exact(case_a):
    g, out = case_a
    np.testing.assert_array_equal(out["tstep"], g["match_timesteps"].astype(np.int64))
    bf2df = _bf_to_df()
    T = g["match_node"].shape[1]
    for b in range(2):
        L = int(g["end_ind"][b]) + 1
        np.testing.assert_array_equal(out["frame_node"][b, :L], bf2df[g["match_node"][b, :L]])
        assert (out["frame_node"][b, L:] == -1).all() and (g["match_node"][b, L:] == 0).all()
        assert T == 200


def test_posterior_tree(case_a):
    g, out = case_a
    t = out["tree"]
    _close(t["e"], g["e_df"], 5e-5)
    _close(t["p_mu"], g["p_mu"], 5e-5)
    _close(t["p_ls"], g["p_log_sigma"], 5e-5)
    _close(t["q_mu"], g["q_mu"], 5e-5)
    _close(t["q_ls"], g["q_log_sigma"], 5e-5)
    _close(out["kl_per_node"], g["kl_per_node"][:, np.argsort(_bf_to_df())], 0, 2e-3)


def test_batch_stat_decoder_and_nll(case_a):
    g, out = case_a
    nodes = g["img_nodes"].tolist()
    _close(out["images_df"][:, nodes], g["images_sel"], 2e-5)
    _close(out["distr_mu"][:, nodes[:2]], g["distr_mu_sel"], 2e-5)
    _close(out["distr_ls"][:, nodes[:2]], g["distr_ls_sel"], 1e-4)
    _close(out["images_df"].double().sum((2, 3, 4)), g["images_sum"], 2e-2)
    _close(out["nll_per_frame"], g["nll_per_frame"], 0.05, 1e-4)


def test_aux_heads(case_a):
    g, out = case_a
    _close(out["existence"], g["existence"], 5e-5)
    _close(out["model_enc_seq"], g["model_enc_seq"], 5e-5)
    _close(out["regressed_state"], g["regressed_state"], 5e-5)
    _close(out["inv_actions"], g["inv_actions"], 5e-5)
    _close(out["cost_pred"], g["cost_pred"], 5e-5)


def _check_losses(g, losses):
    ref = dict(zip([str(n) for n in g["loss_names"]], g["loss_values"]))
    assert set(ref) == set(TO.LOSS_NAMES)
    for k in TO.LOSS_NAMES:
        np.testing.assert_allclose(float(losses[k]), ref[k], rtol=2e-4, atol=1e-6, err_msg=k)


def test_losses_B2(case_a):
    g, out = case_a
    _check_losses(g, out["losses"])


def test_losses_config1_B16(golden_dir, sd):
    """BASELINE config 1: batch 16, T = 200, every loss term of the reference's forward + loss."""
    g = np.load(os.path.join(golden_dir, "train_losses_B16.npz"))
    batch = synthetic_train_batch(16, seed=int(g["batch_seed"]))
    np.testing.assert_array_equal(batch["end_ind"].numpy(), g["end_ind"])
    with torch.no_grad():
        out = TO.forward_loss(sd, batch, _aux(g))
    _check_losses(g, out["losses"])
    _close(out["kl_per_node"].sum(1), g["kl_per_seq"], 0, 1e-3)
    _close(out["nll_per_frame"].sum(1), g["nll_per_seq"], 0, 1e-4)


def test_losses_extreme_lengths_B3(golden_dir, sd):
    """Sequence lengths at the extremes (end_ind = 1: two frames, the shortest the inverse / cost models accept; 199: the
    whole buffer): every loss term, the per-frame NLL and the frame -> node matching of the reference."""
    g = np.load(os.path.join(golden_dir, "train_losses_edge_B3.npz"))
    batch = synthetic_train_batch(3, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    with torch.no_grad():
        out = TO.forward_loss(sd, batch, _aux(g))
    _check_losses(g, out["losses"])
    _close(out["kl_per_node"].sum(1), g["kl_per_seq"], 0, 1e-3)
    _close(out["nll_per_frame"], g["nll_per_frame"], 0, 1e-4)
    assert float(out["nll_per_frame"][0, 2:].abs().max()) == 0.0          # frames past end_ind = 1 carry no loss


def test_oracle_gradients_match_the_reference(sd, golden_dir):
    """Groundwork for the backward pass: autograd through the oracle (with the reference's three detach points) gives the
    gradients `losses.total.value.backward()` left in the UNMODIFIED reference (train_grads_B2.npz,
    oracle/make_golden_train_grad.py) for every one of its 540 parameters: L2 norm, sum and leading entries."""
    import os
    import numpy as np
    import torch
    from oracle import train_oracle as TO
    from video_gcp_b200.synthetic import synthetic_train_batch
    g = np.load(os.path.join(golden_dir, "train_grads_B2.npz"))
    batch = synthetic_train_batch(2, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    aux = dict(inv_t0=g["inv_t0"], inv_t1=g["inv_t1"], cost_start=g["cost_start"], cost_end=g["cost_end"],
               cost_target=g["cost_target"])
    torch.set_num_threads(os.cpu_count())
    losses, grads = TO.loss_gradients(sd, batch, aux)
    assert abs(float(losses["total"].detach()) - float(g["total"])) < 1e-4 * abs(float(g["total"]))
    names = [str(n) for n in g["names"]]
    assert len(names) == 540 and all(n in grads for n in names)
    # stage targets (gradients of the intermediates a staged device backward hands over): shapes, finiteness, and the one
    # closed form among them -- d total / d q_mu of KL + reconstruction is non-zero exactly on ... every node (KL touches all)
    st = grads["__stages__"]
    assert tuple(st["e_df"].shape) == (2, 255, 128) and tuple(st["distr_mu"].shape) == (2, 255, 5, 3, 32, 32)
    assert all(bool(torch.isfinite(t).all()) for t in st.values())
    bound = torch.as_tensor(TO.match_tables(g["end_ind"])[0])                      # [B,255] nodes bound to a frame
    per_node = st["distr_mu"].abs().sum((2, 3, 4, 5))
    assert bool((per_node[~bound] == 0).all()) and bool((per_node[bound] > 0).all())   # only bound nodes are reconstructed
    assert float(st["q_mu"].abs().sum((2,)).min()) > 0.0
    worst_norm = worst_head = 0.0
    for i, n in enumerate(names):
        gr = grads[n].double().reshape(-1)
        ref = float(g["norm"][i])
        worst_norm = max(worst_norm, abs(float(gr.norm()) - ref) / max(ref, 1e-12))
        h = np.zeros(8)
        h[:min(8, gr.numel())] = gr[:8].numpy()
        # leading entries against the scale of the whole gradient (single entries can be arbitrarily small)
        scale = max(float(gr.abs().max()), 1e-12)
        worst_head = max(worst_head, float(np.abs(h - g["head"][i]).max()) / scale)
    print("gradient norms max rel err %.2e, leading entries max err / |grad|_inf %.2e" % (worst_norm, worst_head))
    assert worst_norm < 1e-3 and worst_head < 1e-3
