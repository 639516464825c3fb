"""The hierarchical-optimiser oracle (oracle/hier_oracle.py) against the fixture produced by the unmodified reference
planner (oracle/make_golden_hier.py -> tests/golden/hier_plan.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from oracle import hier_oracle as H
from video_gcp_b200 import hparams
from video_gcp_b200.synthetic import synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def replay():
    g = np.load(os.path.join(GOLDEN, "hier_plan.npz"))
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, int(g["weight_seed"]))
    calls = []

    def rollout_fn(samples):
        end = H.injected_end_ind(len(calls), samples.shape[0])
        calls.append(end)
        with torch.no_grad():
            return O.simulator_rollout(sd, g["state"], g["goal"], samples, end)["predictions"]

    np.random.seed(int(g["np_seed"]))
    del H.ARGMIN_LOG[:]
    torch.set_num_threads(os.cpu_count())
    log = H.plan(sd, rollout_fn, g["goal"])
    return g, log, calls, list(H.ARGMIN_LOG)


def test_samples_follow_the_reference_stream(replay):
    g, log, calls, _ = replay
    zs = log["samples"] + [log["final_samples"]]
    assert len(zs) == int(g["n_calls"])
    assert [len(c) for c in calls] == g["call_sizes"].tolist()
    assert np.array_equal(np.concatenate(calls), g["end_inds"])
    for i, z in enumerate(zs):
        np.testing.assert_allclose(z[:, ::8, ::32], g["samples_sub_%d" % i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(z.sum((1, 2)), g["samples_sum_%d" % i], rtol=1e-9, atol=1e-6)


def test_choices_and_costs(replay):
    g, log, _, argmins = replay
    assert [k for k, _ in argmins] == g["argmin_choice"].tolist()
    assert [len(c) for _, c in argmins] == g["argmin_sizes"].tolist()
    mine = np.concatenate([c for _, c in argmins])
    ref = g["argmin_costs"]
    assert np.array_equal(np.isnan(mine), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(mine[ok], ref[ok], rtol=2e-4, atol=2e-4)
    assert bool(log["complete"]) == bool(g["fully_optimized"])


def test_plans_and_final_rollout(replay):
    g, log, _, _ = replay
    for i in range(3):
        ref = g["plan_%d" % i]
        assert log["plans"][i].shape == ref.shape
        np.testing.assert_allclose(log["plans"][i], ref, rtol=0, atol=2e-5)
        np.testing.assert_allclose(log["costs"][i], g["plan_cost_%d" % i], rtol=2e-4, atol=2e-4)
    fin = log["final_rollouts"][0]
    assert fin.shape == g["final_pred_f16"].shape
    np.testing.assert_allclose(fin, g["final_pred_f16"].astype(np.float32), rtol=0, atol=4e-3)
    assert abs(float(fin.astype(np.float64).sum()) - float(g["final_pred_sum"])) < 1e-2 * max(1.0, abs(float(g["final_pred_sum"])) * 1e-3)
    np.testing.assert_allclose(fin[:, -128:], g["final_latents"], rtol=0, atol=2e-5)


def test_closed_loop_step_oracle():
    """O.infer_action against the reference's encoder + inv_mdl.run_single outputs stored in the fixture."""
    g = np.load(os.path.join(GOLDEN, "hier_plan.npz"))
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, int(g["weight_seed"]))
    for i in range(g["cl_images"].shape[0]):
        with torch.no_grad():
            act, enc = O.infer_action(sd, g["cl_images"][i], g["final_latents"][i + 1])
        np.testing.assert_allclose(enc, g["cl_enc"][i], rtol=0, atol=2e-5)
        np.testing.assert_allclose(act, g["cl_actions"][i], rtol=0, atol=2e-5)
