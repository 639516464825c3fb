"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI.

Tolerances (bf16 tensor-core operands, fp32 accumulation, 8 chained TreeLSTM levels) were set from
tests/gpu_report.py measurements -- observed / allowed:
    node latents          max|d|/max|ref| 7.5e-3 / 2.5e-2   (relative, north_star "latents relative")
    prior mu, log_sigma   1.3e-2 / 4e-2
    decoded frames        max-abs 8.2e-4 / 5e-3             (north_star "frames max-abs"; frames in [-1,1])
    actions / states      1.5e-2 / 5e-2,  existence logits 2.2e-2 / 6e-2
    L2 image cost         2.5e-6 / 1e-4 relative; learned cost 2.2e-3 / 1e-2
Integer work (pruning maps, gathers, top-k indices) is bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200.synthetic import synthetic_rollout_inputs

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
def engine(dev, sd):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=256, attach_cost_mdl=True)
    eng.load_weights(sd)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def case(engine, dev, sd):
    B = 6
    inp = synthetic_rollout_inputs(B, seed=21, shared_images=False)
    inp["end_ind"][:3] = torch.tensor([2, 199, 25])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = O.rollout(sd, inp["I_0"], inp["I_g"], inp["z"], inp["end_ind"].numpy())
    out = engine.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev),
                         want_prior=True)
    torch.cuda.synchronize()
    return inp, ref, out


def test_encoder_and_length_logits(case):
    _, ref, out = case
    assert rel(out["e_0"], ref["e0"]) < 1e-5 and rel(out["e_g"], ref["eg"]) < 1e-5     # fp32 SIMT encoder
    assert rel(out["seq_len_logits"], ref["seq_len_logits"]) < 2.5e-2


def test_tree_latents_and_prior(case):
    _, ref, out = case
    assert rel(out["e_df"], ref["tree"]["e"]) < 2.5e-2
    for lvl in range(8):
        idx = [O.df_index(lvl, j) for j in range(2 ** lvl)]
        assert rel(out["e_df"][:, idx], ref["tree"]["e"][:, idx]) < 2.5e-2, lvl
    assert rel(out["mu_df"], ref["tree"]["mu"]) < 4e-2
    assert rel(out["log_sigma_df"], ref["tree"]["log_sigma"]) < 4e-2


def test_decoded_frames(case):
    _, ref, out = case
    assert maxabs(out["images_df"], ref["images_df"]) < 5e-3
    assert float(out["images_df"].abs().max()) <= 1.0


def test_heads_and_pruned_sequences(case):
    inp, ref, out = case
    lmax = ref["model_enc_seq"].shape[1]
    assert (out["end_ind"].cpu() == inp["end_ind"]).all()
    assert rel(out["existence"], ref["existence"]) < 6e-2
    assert rel(out["model_enc_seq"][:, :lmax], ref["model_enc_seq"]) < 2.5e-2
    assert rel(out["actions"][:, :lmax - 1], ref["actions"]) < 5e-2
    assert rel(out["regressed_state"][:, :lmax], ref["regressed_state"]) < 5e-2
    # zero padding past each candidate's length is exact
    for b, e in enumerate(inp["end_ind"].tolist()):
        assert float(out["model_enc_seq"][b, e + 1:].abs().max() if e < 199 else 0.0) == 0.0


def test_prune_gather_bit_exact(engine, case, dev):
    """Integer path: frames 0..end_ind gathered from the depth-first array exactly as the oracle picks them."""
    inp, ref, out = case
    src = ref["images_df"].to(dev).contiguous()
    got = engine.prune_gather(src, inp["end_ind"].to(dev)).cpu()
    for b, ix in enumerate(ref["prune_idx"]):
        want = ref["images_df"][b, torch.as_tensor(ix)].reshape(len(ix), -1)
        assert torch.equal(got[b, :len(ix)], want)
        assert float(got[b, len(ix):].abs().sum()) == 0.0


def test_prune_every_length_bit_exact(engine, dev):
    """All end_ind 1..199: the device pruning map equals the reference's (golden fixture via the oracle)."""
    ends = torch.arange(1, 200)
    src = torch.arange(255, dtype=torch.float32).repeat(len(ends), 1)[..., None].repeat(1, 1, 4).contiguous()
    got = engine.prune_gather(src.to(dev), ends.to(dev)).cpu()[..., 0].long()
    for i, e in enumerate(ends.tolist()):
        assert got[i, :e + 1].tolist() == O.prune_indices(e).tolist(), e


def test_costs_elites_refit(engine, case, dev, sd):
    inp, ref, out = case
    B = inp["z"].shape[0]
    goal = inp["I_g"][0]
    goal_hwc = ((goal.permute(1, 2, 0)[None] + 1) / 2).numpy()
    imgs = [p.numpy() for p in ref["pruned_images"]]
    ends = inp["end_ind"].to(dev)
    ref_img_dev = ref["images_df"].to(dev).contiguous()
    for dense, w in ((True, 1.0), (False, 2.0), (True, 3.0)):
        want = O.l2_image_cost(imgs, goal_hwc, dense, w)
        got = engine.cost_l2(ref_img_dev, ends, goal.to(dev), dense, w)
        assert rel(got, want) < 1e-5                                   # same inputs: fp32 reduction order only
        got_own = engine.cost_l2(out["images_df"], ends, goal.to(dev), dense, w)
        assert rel(got_own, want) < 1e-4
    lat = [p.numpy() for p in ref["pruned_latents"]]
    with torch.no_grad():
        want = O.image_wrapped_learned_cost(sd, lat)
    got = engine.cost_learned(ref["tree"]["e"].to(dev).contiguous(), ends, ref["pruned_latents"][-1].to(dev))
    assert rel(got, want) < 1e-2
    # elites: identical index sets wherever the cost gaps exceed the tolerance
    c_ref = O.l2_image_cost(imgs, goal_hwc, True, 1.0)
    idx, val = engine.topk(engine.cost_l2(out["images_df"], ends, goal.to(dev), True, 1.0), 3)
    # no silent skip (round-1 review): assert the cost error, then every candidate outside the 2 * err band around the oracle's
    # k-th cost must be in / out of the device's elite set (tests/test_gpu_planner.py does the same at 1024 candidates)
    c_dev = engine.cost_l2(out["images_df"], ends, goal.to(dev), True, 1.0).cpu().numpy().astype(np.float64)
    err = np.abs(c_dev - c_ref).max()
    assert err <= 1e-4 * np.abs(c_ref).max()
    order = np.argsort(c_ref, kind="stable")
    got = set(idx.tolist())
    assert set(np.nonzero(c_ref < c_ref[order[2]] - 2 * err)[0].tolist()) <= got
    assert not (set(np.nonzero(c_ref > c_ref[order[3]] + 2 * err)[0].tolist()) & got)
    assert c_ref[order[3]] - c_ref[order[2]] > 2 * err and idx.tolist() == O.elites(c_ref, B, 0.5).tolist()
    mean, std = engine.refit(inp["z"].to(dev), idx)
    m_ref, s_ref = O.refit(inp["z"].double().numpy(), idx.cpu().numpy())
    assert maxabs(mean, m_ref) < 1e-6 and maxabs(std, s_ref) < 1e-6


def test_topk_exact_with_ties(engine, dev):
    r = np.random.default_rng(5)
    for n, k in ((1, 1), (7, 3), (1024, 102), (5000, 500), (65536, 6553)):
        c = r.integers(0, 50, size=n).astype(np.float32)          # many ties
        idx, val = engine.topk(torch.as_tensor(c).to(dev), k)
        want = np.argsort(c, kind="stable")[:k]
        assert idx.cpu().numpy().tolist() == want.tolist(), n
        assert np.array_equal(val.cpu().numpy(), c[want])


def test_golden_fixture_from_reference(engine, dev, golden_dir):
    """Against outputs of the UNMODIFIED reference (tests/golden/tree_forward_B2.npz)."""
    g = np.load(os.path.join(golden_dir, "tree_forward_B2.npz"))
    inp = synthetic_rollout_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    out = engine.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev),
                         end_ind=torch.as_tensor(g["end_ind"]).to(dev))
    assert rel(out["e_df"], g["e_df"]) < 2.5e-2
    assert maxabs(out["images_df"][:, g["img_nodes"].tolist()], g["images_sel"]) < 5e-3
    assert maxabs(out["images_df"], g["images_f16"].astype(np.float32)) < 6e-3
    assert rel(out["actions"][:, :199], g["actions"]) < 5e-2
    assert rel(out["regressed_state"][:, :200], g["regressed_state"]) < 5e-2
    assert rel(out["existence"], g["existence"]) < 6e-2


def test_tc_kernels_match_simt_verification_kernels(dev, sd):
    """tcgen05 product kernels vs the SIMT verification kernels on identical packed bf16 operands."""
    from video_gcp_b200.engine import Engine
    inp = synthetic_rollout_inputs(5, seed=33, shared_images=True)
    outs = []
    from verify_lib import verify_engine
    for use_ref in (True, False):
        # SIMT cross-check kernels live in the separately built verification library, not in the shipped one
        eng = verify_engine(dev, max_candidates=128, attach_cost_mdl=True) if use_ref else \
            Engine(dev, max_candidates=128, attach_cost_mdl=True)
        eng.load_weights(sd)
        outs.append(eng.rollout(inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev),
                                end_ind=inp["end_ind"].to(dev), images_shared=True))
        torch.cuda.synchronize()
        eng.close()
    a, b = outs
    assert rel(b["e_df"], a["e_df"]) < 1.5e-2          # bf16 re-rounding of intermediates amplifies fp32 order noise
    assert maxabs(b["images_df"], a["images_df"].cpu().numpy()) < 3e-3
    assert rel(b["seq_len_logits"], a["seq_len_logits"]) < 1e-4


def test_candidate_independence_and_shared_images(engine, dev):
    """Size-independent property at BASELINE size: a candidate's rollout does not depend on its batch, and the
    shared-image fast path equals the per-candidate path."""
    inp = synthetic_rollout_inputs(256, seed=44, shared_images=True)
    z, ei = inp["z"].to(dev), inp["end_ind"].to(dev)
    full = engine.rollout(inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), z, end_ind=ei, images_shared=True)
    part = engine.rollout(inp["I_0"][:40].to(dev), inp["I_g"][:40].to(dev), z[:40].contiguous(), end_ind=ei[:40].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(full["e_df"][:40], part["e_df"])
    assert torch.equal(full["images_df"][:40], part["images_df"])
    assert torch.equal(full["actions"][:40], part["actions"])


def test_noise_sampler_statistics_and_regeneration(engine, dev):
    z = engine.sample_noise(64, std_scalar=0.3, seed=9, first_candidate_id=1000)
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.std()) - 0.3) < 2e-3
    ids = torch.tensor([1063, 1000, 1005], dtype=torch.int32, device=dev)
    again = engine.sample_noise_ids(ids, std_scalar=0.3, seed=9)
    assert torch.equal(again[0], z[63]) and torch.equal(again[1], z[0]) and torch.equal(again[2], z[5])
    mean = torch.full((255, 256), 2.0, device=dev)
    std = torch.full((255, 256), 0.0, device=dev)
    z2 = engine.sample_noise(3, mean=mean, std=std, seed=1, clip=1.5)
    assert float(z2.min()) == 1.5 and float(z2.max()) == 1.5


def test_model_and_simulator_drop_in(dev, sd):
    """Reference-facing API: model(inputs) under val_mode, dense_rec.get_sample_with_len, simulator.rollout."""
    from video_gcp_b200 import hparams
    from video_gcp_b200.model import TreeModel
    from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost
    model = TreeModel(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True), None, max_candidates=128)
    model.load_state_dict(sd, strict=True)
    model.to(dev)
    model.device = dev
    model.eval()
    r = np.random.default_rng(7)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    N = 5
    samples = r.normal(0, 0.3, size=(N, 255, 256))
    end = r.integers(2, 200, size=N)
    model.inject_end_ind = torch.as_tensor(end)
    sim = GCPImageSimulator(model, append_latent=True)
    ro = sim.rollout(state, goal, samples, 200)
    with torch.no_grad():
        want = O.simulator_rollout(sd, state, goal, samples, end)
    for key in ("predictions", "actions", "states", "latents"):
        assert [a.shape for a in ro[key]] == [a.shape for a in want[key]], key
    assert max(np.abs(a[:, :3072] - b[:, :3072]).max() for a, b in zip(ro.predictions, want["predictions"])) < 5e-3
    assert rel(np.concatenate(ro.latents), np.concatenate(want["latents"])) < 2.5e-2
    cost = L2ImageCost(True, 1.0, engine=model.engine)(ro.predictions, goal)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in want["predictions"]]
    assert rel(cost, O.l2_image_cost(imgs, goal, True, 1.0)) < 1e-4


def test_host_noise_upload_matches_device_noise(engine, dev):
    """Pinned host noise uploaded by the library (copy stream, per-level events) gives bit-identical results."""
    inp = synthetic_rollout_inputs(40, seed=77, shared_images=True)
    I0, Ig, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev)
    a = engine.rollout(I0, Ig, inp["z"].to(dev), end_ind=ei, images_shared=True, fresh=True)
    for _ in range(3):       # repeated calls reuse the staging buffer: ordering against earlier readers
        b = engine.rollout(I0, Ig, inp["z"].pin_memory(), end_ind=ei, images_shared=True, fresh=True)
    torch.cuda.synchronize()
    assert torch.equal(b["z"].cpu(), inp["z"])
    for k in ("e_df", "images_df", "actions", "existence"):
        assert torch.equal(a[k], b[k]), k
