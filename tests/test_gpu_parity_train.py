"""GPU parity for the training-phase forward + loss (BASELINE config 1): gcpb200_forward_loss through
the C ABI against oracle/train_oracle.py (same seeded weights, batch, posterior noise and auxiliary pair indices) and
against the fixtures recorded from the unmodified reference (tests/golden/train_*.npz).  Run with -m gpu on the B200.

Tolerances (bf16 tensor-core operands with fp32 accumulation in the tree + dense decoder layers, fp32 SIMT
batch-statistic encoders; observed / allowed are printed by the tests and recorded in profiles/):
    batch-stat encoders e_0, e_g, enc_traj_seq   fp32, 1e-4 relative
    inference encoder, length logits             2.5e-2 relative
    node latents                                 2.9e-2 / 5e-2 relative
    prior/posterior mu, log_sigma                5.0e-2 / 8e-2 relative
    decoded frames                               1.2e-2 / 3e-2 max-abs (frames in [-1,1]; batch-stat BN over 510 node images)
    aux heads                                    4.1e-2 / 8e-2 relative
  These are wider than the planner-side rollout's (2.5e-2 / 4e-2 / 5e-3) for a reason the CPU can reproduce: with the
  random-init weights the posterior reaches sigma = 17.8, and z = mu + sigma * eps multiplies log-sigma rounding by
  sigma * |eps|.  train_oracle.bf16_operands() (the fp32 oracle with only the GEMM operands rounded to bf16) shows the
  same error per level as the device (level 0: 2.79e-2 vs 2.81e-2), and test_posterior_tree also checks the device
  against that envelope level by level.
    losses                                       2e-2 relative each (KL 5e-2: ratio of two bf16-rounded Gaussians)
Integer work (matched time steps) is bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as TO
from video_gcp_b200 import _C
from video_gcp_b200.synthetic import synthetic_train_batch

pytestmark = pytest.mark.gpu

ALL = ("nll_per_frame", "kl_per_seq", "e_0", "e_g", "enc_traj_seq", "inf_enc_seq", "seq_len_logits", "e_df", "p_mu",
       "p_log_sigma", "q_mu", "q_log_sigma", "match_timesteps", "images_df", "existence", "model_enc_seq",
       "regressed_state", "inv_actions", "cost_pred")


def rel(got, ref):
    got = got.detach().double().cpu()
    ref = torch.as_tensor(np.asarray(ref)).double()
    assert not torch.isnan(got).any()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def engine(dev, sd):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=128, attach_cost_mdl=True)
    eng.load_weights(sd)
    yield eng
    eng.close()


def _aux(g):
    return dict(inv_t0=g["inv_t0"], inv_t1=g["inv_t1"], cost_start=g["cost_start"], cost_end=g["cost_end"],
                cost_target=g["cost_target"])


def _run(engine, dev, batch, aux, want=ALL, cost_target="given"):
    d = lambda t: t.to(dev)
    out = engine.forward_loss(d(batch["traj_seq"]), d(batch["pad_mask"]), d(batch["end_ind"]), d(batch["states"]),
                              d(batch["actions"]), d(batch["eps"]), aux["inv_t0"], aux["inv_t1"], aux["cost_start"],
                              aux["cost_end"], cost_target=aux["cost_target"] if cost_target == "given" else None,
                              I_0=d(batch["I_0"]), I_g=d(batch["I_g"]), want=want)
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in out.items()}


@pytest.fixture(scope="module")
def case_a(engine, dev, sd, golden_dir):
    g = np.load(os.path.join(golden_dir, "train_forward_B2.npz"))
    batch = synthetic_train_batch(2, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = TO.forward_loss(sd, batch, _aux(g))
    out = _run(engine, dev, batch, _aux(g))
    return g, batch, ref, out


def test_batch_stat_encoders(case_a):
    g, _, ref, out = case_a
    errs = dict(e_0=rel(out["e_0"], ref["e0"]), e_g=rel(out["e_g"], ref["eg"]),
                enc_traj_seq=rel(out["enc_traj_seq"], ref["enc_traj_seq"]),
                inf_enc_seq=rel(out["inf_enc_seq"], ref["inf_enc_seq"]),
                seq_len_logits=rel(out["seq_len_logits"], ref["seq_len_logits"]))
    print("train encoders", errs)
    assert errs["e_0"] < 1e-4 and errs["e_g"] < 1e-4 and errs["enc_traj_seq"] < 1e-4
    assert errs["inf_enc_seq"] < 2.5e-2 and errs["seq_len_logits"] < 2.5e-2
    assert rel(out["e_0"], g["e0"]) < 1e-4 and rel(out["enc_traj_seq"], g["enc_traj_seq"]) < 1e-4      # the reference itself


def test_matching_bit_exact(case_a):
    g, _, ref, out = case_a
    np.testing.assert_array_equal(out["match_timesteps"].cpu().numpy().astype(np.int64), ref["tstep"])
    np.testing.assert_array_equal(out["match_timesteps"].cpu().numpy().astype(np.int64), g["match_timesteps"].astype(np.int64))


def test_posterior_tree(case_a):
    g, _, ref, out = case_a
    t = ref["tree"]
    errs = dict(e=rel(out["e_df"], t["e"]), p_mu=rel(out["p_mu"], t["p_mu"]), p_ls=rel(out["p_log_sigma"], t["p_ls"]),
                q_mu=rel(out["q_mu"], t["q_mu"]), q_ls=rel(out["q_log_sigma"], t["q_ls"]))
    print("train tree", errs)
    from oracle.gcp_oracle import df_index
    for l in range(8):
        idx = [df_index(l, j) for j in range(2 ** l)]
        print("  level %d: e %.2e  p_mu %.2e  p_ls %.2e  q_mu %.2e  q_ls %.2e" % (
            l, rel(out["e_df"][:, idx], t["e"][:, idx]), rel(out["p_mu"][:, idx], t["p_mu"][:, idx]),
            rel(out["p_log_sigma"][:, idx], t["p_ls"][:, idx]), rel(out["q_mu"][:, idx], t["q_mu"][:, idx]),
            rel(out["q_log_sigma"][:, idx], t["q_ls"][:, idx])))
    assert errs["e"] < 5e-2 and max(errs.values()) < 8e-2
    assert rel(out["e_df"], g["e_df"]) < 5e-2 and rel(out["q_mu"], g["q_mu"]) < 8e-2


def test_tree_error_is_operand_rounding(case_a, sd):
    """The device-vs-fp32 difference of every level must stay within 2x the difference the fp32 oracle shows against
    ITSELF when only its GEMM operands are rounded to bf16 (plus a small floor) -- i.e. nothing but operand rounding."""
    g, batch, ref, out = case_a
    from oracle.gcp_oracle import df_index
    t = ref["tree"]
    with torch.no_grad(), TO.bf16_operands():
        em = TO.tree_inference(sd, ref["e0"], ref["eg"], ref["inf_enc_seq"], ref["tstep"], batch["eps"])
    names = (("e", "e_df"), ("p_mu", "p_mu"), ("p_ls", "p_log_sigma"), ("q_mu", "q_mu"), ("q_ls", "q_log_sigma"))
    for l in range(8):
        idx = [df_index(l, j) for j in range(2 ** l)]
        row = []
        for k, ko in names:
            env = rel(em[k][:, idx], t[k][:, idx])
            dv = rel(out[ko][:, idx], t[k][:, idx])
            row.append("%s %.2e/%.2e" % (k, dv, env))
            assert dv < 2 * env + 5e-3, (l, k, dv, env)
        print("  level %d device/envelope: %s" % (l, "  ".join(row)))


def test_decoder_and_nll(case_a):
    g, _, ref, out = case_a
    i_err = float((out["images_df"].cpu() - ref["images_df"]).abs().max())
    nodes = g["img_nodes"].tolist()
    g_err = float((out["images_df"].cpu()[:, nodes] - torch.as_tensor(g["images_sel"])).abs().max())
    nll_err = rel(out["nll_per_frame"], ref["nll_per_frame"])
    print("train decoder: images max-abs %.3e (reference fixture %.3e), nll_per_frame rel %.3e" % (i_err, g_err, nll_err))
    assert i_err < 3e-2 and g_err < 3e-2
    assert nll_err < 2e-2
    # frames past end_ind carry zero weight
    pad = (torch.arange(200)[None] <= torch.as_tensor(g["end_ind"])[:, None])
    assert float(out["nll_per_frame"].cpu()[~pad].abs().max()) == 0.0


def test_aux_heads(case_a):
    g, _, ref, out = case_a
    Lmax = ref["model_enc_seq"].shape[1]
    errs = dict(existence=rel(out["existence"], ref["existence"]),
                model_enc_seq=rel(out["model_enc_seq"][:, :Lmax], ref["model_enc_seq"]),
                regressed_state=rel(out["regressed_state"][:, :Lmax], ref["regressed_state"]),
                inv_actions=rel(out["inv_actions"], ref["inv_actions"]),
                cost_pred=rel(out["cost_pred"].reshape(-1), ref["cost_pred"].reshape(-1)))
    print("train heads", errs)
    assert max(errs.values()) < 8e-2


def _check_losses(losses, ref_by_name, label):
    got = dict(zip(_C.LOSS_NAMES, losses.cpu().tolist()))
    rows = []
    for k in _C.LOSS_NAMES:
        r = float(ref_by_name[k])
        tol = 5e-2 if k == "kl" else 2e-2
        err = abs(got[k] - r) / max(abs(r), 1e-6) if r != 0 else abs(got[k])
        rows.append((k, got[k], r, err))
        assert err < tol, (label, k, got[k], r)
    print(label, " ".join("%s=%.5g(ref %.5g, %.1e)" % r for r in rows))


def test_losses_B2(case_a):
    g, _, ref, out = case_a
    _check_losses(out["losses"], {k: float(v) for k, v in ref["losses"].items()}, "losses B2 vs oracle:")
    _check_losses(out["losses"], dict(zip([str(n) for n in g["loss_names"]], g["loss_values"])), "losses B2 vs reference:")


def test_losses_config1_B16(engine, dev, sd, golden_dir):
    """BASELINE config 1: batch 16, T = 200, every loss term against the reference's own values."""
    g = np.load(os.path.join(golden_dir, "train_losses_B16.npz"))
    batch = synthetic_train_batch(16, seed=int(g["batch_seed"]))
    out = _run(engine, dev, batch, _aux(g), want=("nll_per_frame", "kl_per_seq"))
    _check_losses(out["losses"], dict(zip([str(n) for n in g["loss_names"]], g["loss_values"])), "losses B16 vs reference:")
    assert rel(out["kl_per_seq"], g["kl_per_seq"]) < 5e-2
    assert rel(out["nll_per_frame"].sum(1), g["nll_per_seq"]) < 2e-2


def test_losses_extreme_lengths_B3(engine, dev, golden_dir):
    """end_ind = 1 (two frames), 199 (the whole buffer) and 100 in one batch, against the reference's own values."""
    g = np.load(os.path.join(golden_dir, "train_losses_edge_B3.npz"))
    batch = synthetic_train_batch(3, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    out = _run(engine, dev, batch, _aux(g), want=("nll_per_frame", "kl_per_seq"))
    _check_losses(out["losses"], dict(zip([str(n) for n in g["loss_names"]], g["loss_values"])), "losses edge B3 vs reference:")
    assert rel(out["kl_per_seq"], g["kl_per_seq"]) < 5e-2
    assert rel(out["nll_per_frame"], g["nll_per_frame"]) < 2e-2
    assert float(out["nll_per_frame"][0, 2:].abs().max()) == 0.0


def test_device_cost_target_matches_reference(engine, dev, golden_dir):
    """cost_target = NULL: EuclideanPathLength of the ground-truth frames is computed on the device (cost_fcn.py:39-59);
    the cost-estimation loss must equal the one obtained with the reference's recorded target."""
    g = np.load(os.path.join(golden_dir, "train_forward_B2.npz"))
    batch = synthetic_train_batch(2, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    a = _run(engine, dev, batch, _aux(g), want=("cost_pred",))
    b = _run(engine, dev, batch, _aux(g), want=("cost_pred",), cost_target="device")
    i = _C.LOSS_NAMES.index("cost_estimation")
    assert abs(float(a["losses"][i]) - float(b["losses"][i])) <= 1e-4 * abs(float(a["losses"][i]))


def test_repeatable(engine, dev, golden_dir):
    """Same inputs twice -> the same losses (no atomics-order dependence beyond fp32 rounding of the final reductions)."""
    g = np.load(os.path.join(golden_dir, "train_forward_B2.npz"))
    batch = synthetic_train_batch(2, seed=int(g["batch_seed"]), end_ind=g["end_ind"])
    a = _run(engine, dev, batch, _aux(g), want=("kl_per_seq",))
    b = _run(engine, dev, batch, _aux(g), want=("kl_per_seq",))
    np.testing.assert_allclose(a["losses"].cpu().numpy(), b["losses"].cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_rejects_bad_arguments(engine, dev):
    batch = synthetic_train_batch(2, seed=3)
    z = np.zeros(2, np.int64)
    with pytest.raises(AssertionError):
        engine.forward_loss(batch["traj_seq"][:, :100].to(dev), batch["pad_mask"].to(dev), batch["end_ind"].to(dev),
                            batch["states"].to(dev), batch["actions"].to(dev), batch["eps"].to(dev), z, z, z, z)


def test_model_drop_in_train_forward_and_loss(dev, sd, golden_dir):
    """The module-level API train.py calls (train.py:155-157 and its validation pass :204-206): `model(inputs)` outside
    val_mode, `model.loss`, `model.get_total_loss`.  With np.random seeded like the reference run the model draws the
    same auxiliary pairs; with the fixture's posterior noise every loss must equal the unmodified reference's."""
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.types import AttrDict
    g = np.load(os.path.join(golden_dir, "train_losses_B16.npz"))
    batch = synthetic_train_batch(16, seed=int(g["batch_seed"]))
    model = TreeModel(hparams.gcp_tree_25room_config(batch_size=16, attach_cost_mdl=True), None, max_candidates=128)
    model.load_state_dict(sd, strict=True)
    model.device = dev
    model.train()
    d = lambda t: t.to(dev)
    inputs = AttrDict(traj_seq=d(batch["traj_seq"]), traj_seq_images=d(batch["traj_seq"]), pad_mask=d(batch["pad_mask"]),
                      end_ind=d(batch["end_ind"]), traj_seq_states=d(batch["states"]), actions=d(batch["actions"]),
                      I_0=d(batch["I_0"]), I_g=d(batch["I_g"]), eps=d(batch["eps"]))
    np.random.seed(int(g["np_seed"]))
    out = model(inputs)
    losses = model.loss(inputs, out)
    losses.total = model.get_total_loss(inputs, losses)
    torch.cuda.synchronize()
    for k in ("inv_t0", "inv_t1", "cost_start", "cost_end"):
        np.testing.assert_array_equal(out.aux_indices[k], g[k])
    ref = dict(zip([str(n) for n in g["loss_names"]], g["loss_values"]))
    assert set(losses.keys()) == set(ref.keys())
    vec = torch.stack([losses[k].value for k in _C.LOSS_NAMES])
    _check_losses(vec, ref, "drop-in losses B16 vs reference:")
    assert rel(losses.kl.breakdown, g["kl_per_seq"]) < 5e-2
    lmax = int(batch["end_ind"].max()) + 1
    assert tuple(out.regressed_state.shape) == (16, lmax, 2) and tuple(inputs.model_enc_seq.shape) == (16, lmax, 128)
    assert tuple(out.tree.bf.e_g_prime.shape) == (16, 255, 128, 1, 1) and tuple(out.tree.df.images.shape) == (16, 255, 3, 32, 32)
    # without injected noise the model samples the posterior noise itself (device Philox): finite, different losses
    del inputs["eps"]
    kl1 = float(losses.kl.value)
    out2 = model(inputs)
    l2 = model.loss(inputs, out2)
    assert torch.isfinite(torch.stack([l2[k].value for k in _C.LOSS_NAMES[:8]])).all()
    assert float(l2.kl.value) != kl1 and float(losses.kl.value) == kl1
    model.engine.close()
