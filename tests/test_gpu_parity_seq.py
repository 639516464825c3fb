"""GPU parity tests of the sequential GCP rollout (config 3); run with -m gpu on the B200 box.  Everything goes
through the C ABI (gcpb200_seq_rollout / gcpb200_cost_l2_seq).

Tolerances (bf16 tensor-core operands, fp32 accumulation, fp32 cell state, 199 chained LSTM steps; measured with
tests/gpu_report_seq.py on B200 -- observed / allowed):
    predicted latents     max|d|/max|ref| 2.3e-2 / 5e-2   (5e-3 after step 0, 2.6e-2 after step 198: bf16 operand
                          rounding accumulating along the recurrence; the SIMT verification kernels show the same)
    prior mu, log_sigma   6.5e-2 / 1e-1  (rms 1.8e-2)
    decoded frames        max-abs 1.4e-3 / 5e-3           (frames in [-1,1]; frame 0 = I_0 is bit-exact)
    actions / states      4.9e-2 / 1e-1  (rms 1.7e-2)
Integer / copy work (frame 0, zero padding, sequence cut) is bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200 import hparams
from video_gcp_b200.synthetic import synthetic_seq_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu

SEQ_LAT_TOL, SEQ_PRIOR_TOL, SEQ_IMG_TOL, SEQ_HEAD_TOL = 5e-2, 1e-1, 5e-3, 1e-1


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
def seq_sd():
    hp = hparams.build_hparams(hparams.gcp_sequential_25room_config(batch_size=1))
    return synthetic_state_dict(hp, 2)


@pytest.fixture(scope="module")
def seq_engine(dev, seq_sd):
    from video_gcp_b200.engine import Engine
    eng = Engine(dev, max_candidates=256, model="sequential")
    eng.load_weights(seq_sd)
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def case(seq_engine, dev, seq_sd):
    B = 5
    inp = synthetic_seq_inputs(B, seed=31, shared_images=False)
    given = torch.tensor([199, 199, 60, 2, 199])
    pred = torch.tensor([2, 199, 25, 100, 57])
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = O.seq_rollout(seq_sd, inp["I_0"], inp["I_g"], inp["z"], given.numpy())
    out = seq_engine.seq_rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=pred.to(dev),
                                 given_end_ind=given.to(dev), want_prior=True)
    torch.cuda.synchronize()
    return inp, ref, out, given, pred


def test_seq_encoder_and_length(case):
    _, ref, out, _, pred = case
    assert rel(out["e_0"], ref["e0"]) < 1e-5 and rel(out["e_g"], ref["eg"]) < 1e-5
    assert rel(out["seq_len_logits"], ref["seq_len_logits"]) < 2.5e-2
    assert (out["end_ind"].cpu() == pred).all()


def test_seq_latents_and_prior(case):
    _, ref, out, _, _ = case
    assert rel(out["encodings"], ref["encodings"]) < SEQ_LAT_TOL
    for t0 in (0, 50, 100, 150):           # no drift hiding behind the global maximum
        sl = slice(t0, t0 + 49)
        assert rel(out["encodings"][:, sl], ref["encodings"][:, sl]) < SEQ_LAT_TOL, t0
    assert rel(out["mu"], ref["mu"]) < SEQ_PRIOR_TOL
    assert rel(out["log_sigma"], ref["log_sigma"]) < SEQ_PRIOR_TOL


def test_seq_decoded_frames(case):
    inp, ref, out, _, _ = case
    assert torch.equal(out["images"][:, 0].cpu(), inp["I_0"])           # frame 0 is the start image itself
    assert maxabs(out["images"], ref["images"]) < SEQ_IMG_TOL
    assert float(out["images"][:, 1:].abs().max()) <= 1.0


def test_seq_aux_heads_and_padding(case):
    _, ref, out, given, _ = case
    lmax = ref["model_enc_seq"].shape[1]
    assert rel(out["model_enc_seq"][:, :lmax], ref["model_enc_seq"]) < SEQ_LAT_TOL
    assert rel(out["actions"][:, :lmax - 1], ref["actions"]) < SEQ_HEAD_TOL
    assert rel(out["regressed_state"][:, :lmax], ref["regressed_state"]) < SEQ_HEAD_TOL
    for b, e in enumerate(given.tolist()):
        if e < 199:
            assert float(out["model_enc_seq"][b, e + 1:].abs().max()) == 0.0
        assert torch.equal(out["model_enc_seq"][b, 0], out["e_0"][b])     # first row is e_0 (sequential.py:90-92)
        assert torch.equal(out["model_enc_seq"][b, 1:e + 1], out["encodings"][b, :e])


def test_seq_golden_fixture_from_reference(seq_engine, dev, golden_dir):
    """Against outputs of the UNMODIFIED reference SequentialModel (tests/golden/seq_forward_B2.npz)."""
    g = np.load(os.path.join(golden_dir, "seq_forward_B2.npz"))
    inp = synthetic_seq_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    out = seq_engine.seq_rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev),
                                 end_ind=torch.as_tensor(g["end_pred"]).to(dev), want_prior=True)
    assert rel(out["encodings"], g["encodings"]) < SEQ_LAT_TOL
    assert rel(out["mu"], g["mu"]) < SEQ_PRIOR_TOL
    assert maxabs(out["images"][:, g["img_t"].tolist()], g["images_sel"]) < SEQ_IMG_TOL
    assert maxabs(out["images"], g["images_f16"].astype(np.float32)) < SEQ_IMG_TOL + 1e-3
    assert rel(out["actions"][:, :199], g["actions"]) < SEQ_HEAD_TOL
    assert rel(out["regressed_state"], g["regressed_state"]) < SEQ_HEAD_TOL
    assert rel(out["model_enc_seq"], g["model_enc_seq"]) < SEQ_LAT_TOL


def test_seq_tc_kernels_match_simt_verification_kernels(dev, seq_sd):
    from video_gcp_b200.engine import Engine
    inp = synthetic_seq_inputs(3, seed=35, shared_images=True)
    outs = []
    from verify_lib import verify_engine
    for use_ref in (True, False):
        eng = verify_engine(dev, max_candidates=128, model="sequential") if use_ref else \
            Engine(dev, max_candidates=128, model="sequential")
        eng.load_weights(seq_sd)
        outs.append(eng.seq_rollout(inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev),
                                    end_ind=inp["end_ind"].to(dev), images_shared=True, fresh=True))
        torch.cuda.synchronize()
        eng.close()
    a, b = outs
    assert rel(b["encodings"], a["encodings"]) < 2.5e-2
    assert maxabs(b["images"], a["images"].cpu().numpy()) < 3e-3


def test_seq_candidate_independence(seq_engine, dev):
    """Size-independent property: a candidate's rollout does not depend on its batch; shared-image fast path equals
    the per-candidate path."""
    inp = synthetic_seq_inputs(256, seed=46, shared_images=True)
    z, ei = inp["z"].to(dev), inp["end_ind"].to(dev)
    full = seq_engine.seq_rollout(inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), z, end_ind=ei, images_shared=True, fresh=True)
    part = seq_engine.seq_rollout(inp["I_0"][:40].to(dev), inp["I_g"][:40].to(dev), z[:40].contiguous(),
                                  end_ind=ei[:40].contiguous(), fresh=True)
    torch.cuda.synchronize()
    assert torch.equal(full["encodings"][:40], part["encodings"])
    assert torch.equal(full["images"][:40], part["images"])
    assert torch.equal(full["actions"][:40], part["actions"])


def test_seq_graph_replay_is_bit_identical(seq_engine, dev):
    """The CUDA-graph replay of the recurrence (second and later calls on the same buffers) equals the first,
    directly launched call bit for bit, also after the noise buffer's contents change."""
    inp = synthetic_seq_inputs(130, seed=52, shared_images=True)
    I0, Ig = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev)
    z = inp["z"].to(dev)
    outs = []
    for i in range(3):
        o = seq_engine.seq_rollout(I0, Ig, z, end_ind=inp["end_ind"].to(dev), images_shared=True, want_prior=True)
        outs.append({k: o[k].clone() for k in ("encodings", "mu", "images", "actions")})
    z2 = torch.roll(inp["z"], 1, 0).to(dev)
    z.copy_(z2)
    o = seq_engine.seq_rollout(I0, Ig, z, end_ind=inp["end_ind"].to(dev), images_shared=True, want_prior=True)
    torch.cuda.synchronize()
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]) and torch.equal(outs[0][k], outs[2][k]), k
    assert torch.equal(o["encodings"][1:], outs[0]["encodings"][:-1])      # rolled noise -> rolled rollouts


def test_seq_model_and_simulator_drop_in(dev, seq_sd, golden_dir):
    """Reference-facing API: SequentialModel under val_mode through GCPImageSimulator.rollout + L2 cost, against the
    oracle and against the golden fixture produced by the reference simulator."""
    from video_gcp_b200.model import SequentialModel
    from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost
    g = np.load(os.path.join(golden_dir, "seq_sim_N6.npz"))
    model = SequentialModel(hparams.gcp_sequential_25room_config(batch_size=1), None, max_candidates=128)
    model.load_state_dict(seq_sd, strict=True)
    model.to(dev)
    model.device = dev
    model.eval()
    N = 6
    r = np.random.default_rng(int(g["rng_seed"]))
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 1.0, size=(N, 199, 256))
    end = r.integers(2, 200, size=N)
    model.inject_end_ind = torch.as_tensor(end)
    sim = GCPImageSimulator(model, append_latent=True)
    dro = sim.rollout_device(state, goal, samples, 200)
    ro = dro.to_host(True)
    with torch.no_grad():
        want = O.seq_simulator_rollout(seq_sd, state, goal, samples, end)
    for key in ("predictions", "actions", "states", "latents"):
        assert [a.shape for a in ro[key]] == [a.shape for a in want[key]], key
    assert [p.shape[0] for p in ro.predictions] == g["pred_len"].tolist()
    assert max(np.abs(a[:, :3072] - b[:, :3072]).max() for a, b in zip(ro.predictions, want["predictions"])) < SEQ_IMG_TOL
    assert rel(np.concatenate(ro.latents), np.concatenate(want["latents"])) < SEQ_LAT_TOL
    assert rel(np.concatenate(ro.actions), np.concatenate(want["actions"])) < SEQ_HEAD_TOL
    assert np.abs(ro.predictions[2][:, :3072] - g["pred2"].astype(np.float32)[:, :3072]).max() < SEQ_IMG_TOL + 1e-3
    cost = L2ImageCost(True, 1.0).device_cost(dro).cpu().numpy()
    assert rel(cost, g["l2_dense"]) < 1e-4
    idx, _ = model.engine.topk(torch.as_tensor(cost).to(dev), 2)
    assert idx.tolist() == np.argsort(g["l2_dense"], kind="stable")[:2].tolist()
