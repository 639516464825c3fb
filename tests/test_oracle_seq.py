"""Pins the sequential-GCP part of oracle/gcp_oracle.py against fixtures produced by the unmodified reference
SequentialModel (oracle/make_golden_seq.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200 import hparams, spec
from video_gcp_b200.synthetic import synthetic_seq_inputs, synthetic_state_dict


@pytest.fixture(scope="module")
def seq_sd():
    hp = hparams.build_hparams(hparams.gcp_sequential_25room_config(batch_size=1))
    return synthetic_state_dict(hp, 2)


def _close(a, b, atol, rtol=0.0):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, atol=atol, rtol=rtol)


def test_sequential_manifest(golden_dir):
    hp = hparams.build_hparams(hparams.gcp_sequential_25room_config(batch_size=1))
    mine = {k: list(v) for k, v in spec.full_manifest(hp).items()}
    with open(os.path.join(golden_dir, "state_dict_manifest.json")) as f:
        ref = json.load(f)["sequential"]
    assert set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)


def test_sequential_forward(golden_dir, seq_sd):
    g = np.load(os.path.join(golden_dir, "seq_forward_B2.npz"))
    inp = synthetic_seq_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = O.seq_rollout(seq_sd, inp["I_0"], inp["I_g"], inp["z"], np.full(2, 199))
    _close(out["e0"], g["e0"], 1e-5)
    _close(out["seq_len_logits"], g["seq_len_logits"], 2e-5)
    _close(out["encodings"], g["encodings"], 5e-5)
    _close(out["mu"], g["mu"], 5e-5)
    _close(out["log_sigma"], g["log_sigma"], 5e-5)
    _close(out["images"][:, g["img_t"].tolist()], g["images_sel"], 5e-5)
    _close(out["images"], g["images_f16"].astype(np.float32), 1e-3)
    _close(out["images"].double().sum((2, 3, 4)), g["images_sum"], 5e-2)
    assert torch.equal(out["images"][:, 0], inp["I_0"])          # frame 0 is the start image itself
    _close(out["model_enc_seq"], g["model_enc_seq"], 5e-5)
    _close(out["actions"], g["actions"], 5e-5)
    _close(out["regressed_state"], g["regressed_state"], 5e-5)


def test_sequential_simulator_and_cost(golden_dir, seq_sd):
    g = np.load(os.path.join(golden_dir, "seq_sim_N6.npz"))
    N = 6
    r = np.random.default_rng(int(g["rng_seed"]))
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    samples = r.normal(0, 1.0, size=(N, 199, 256))
    end = r.integers(2, 200, size=N)
    assert (end == g["end_ind"]).all() and np.array_equal(goal, g["goal"])
    with torch.no_grad():
        ro = O.seq_simulator_rollout(seq_sd, state, goal, samples, end)
    assert [p.shape[0] for p in ro["predictions"]] == g["pred_len"].tolist()
    _close(ro["predictions"][2], g["pred2"].astype(np.float32), 2e-3)
    _close(ro["latents"][2], g["lat2"], 5e-5)
    _close(ro["actions"][2], g["act2"], 5e-5)
    _close(ro["states"][2], g["state2"], 5e-5)
    np.testing.assert_allclose([p.astype(np.float64).sum() for p in ro["predictions"]], g["pred_sum"], atol=5e-2)
    imgs = [p[:, :3072].reshape(-1, 3, 32, 32) for p in ro["predictions"]]
    np.testing.assert_allclose(O.l2_image_cost(imgs, goal, True, 1.0), g["l2_dense"], rtol=1e-5)
