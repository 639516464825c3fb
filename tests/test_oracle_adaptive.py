"""Pins the adaptive-binding part of oracle/gcp_oracle.py against fixtures produced by the unmodified reference
(oracle/make_golden_adaptive.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from video_gcp_b200 import hparams, spec
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict


@pytest.fixture(scope="module")
def ada_sd():
    hp = hparams.build_hparams(hparams.gcp_adaptive_25room_config(batch_size=1))
    return synthetic_state_dict(hp, 3)


def _close(a, b, atol):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, atol=atol, rtol=0)


def test_adaptive_manifest(golden_dir):
    hp = hparams.build_hparams(hparams.gcp_adaptive_25room_config(batch_size=1))
    mine = {k: list(v) for k, v in spec.full_manifest(hp).items()}
    with open(os.path.join(golden_dir, "state_dict_manifest.json")) as f:
        ref = json.load(f)["adaptive"]
    assert set(mine) == set(ref)
    assert all(mine[k] == ref[k] for k in ref)


def test_adaptive_forward(golden_dir, ada_sd):
    g = np.load(os.path.join(golden_dir, "adaptive_forward_B2.npz"))
    inp = synthetic_rollout_inputs(2, seed=int(g["input_seed"]), shared_images=False)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        out = O.adaptive_rollout(ada_sd, inp["I_0"], inp["I_g"], inp["z"])
    _close(out["seq_len_logits"], g["seq_len_logits"], 2e-5)
    _close(out["tree"]["e"], g["e_df"], 2e-5)
    _close(out["images_df"][:, g["img_nodes"].tolist()], g["images_sel"], 2e-5)
    _close(out["images_df"], g["images_f16"].astype(np.float32), 1e-3)
    _close(out["images_df"].double().sum((2, 3, 4)), g["images_sum"], 2e-2)
    _close(out["distances"], g["distances"], 5e-5)
    assert [int(k.sum()) for k in out["keep"]] == g["pruned_len"].tolist()
    assert 0 < int((~out["keep"]).sum())                       # the fixture really prunes something
    _close(out["pruned_images"][0], g["pruned0"].astype(np.float32), 1e-3)
    np.testing.assert_allclose([float(p.double().sum()) for p in out["pruned_images"]], g["pruned_sum"], atol=2e-2)
