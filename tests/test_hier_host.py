"""Host logic of the device hierarchical optimiser (video_gcp_b200/planning/tree_optimizer.py, sampler.py) on CPU:
the index-based optimiser is driven with the reference-style list-of-numpy rollouts (from the CPU oracle) and a cost
function whose `pairs_device` is the oracle's cost MLP, and must reproduce the decision trace, costs and plans of the
fixture made by the unmodified reference planner (tests/golden/hier_plan.npz)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import gcp_oracle as O
from oracle import hier_oracle as H
from video_gcp_b200 import hparams
from video_gcp_b200.planning import sampler as S
from video_gcp_b200.planning import tree_optimizer as T
from video_gcp_b200.synthetic import synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleCost:
    """LearnedCostEstimate stand-in: same `pairs_device` contract, arithmetic by the CPU oracle."""
    input_dim = 128

    def __init__(self, sd):
        self.sd = sd
        self.engine = types.SimpleNamespace(device=torch.device("cpu"))

    def pairs_device(self, lat, idx1, idx2, seg_off=None):
        i1 = torch.as_tensor(np.asarray(idx1), dtype=torch.long)
        i2 = torch.as_tensor(np.asarray(idx2), dtype=torch.long)
        with torch.no_grad():
            c = O.mlp(self.sd, "cost_mdl.cost_pred", torch.cat([lat[i1], lat[i2]], 1), conv=False)[:, 0]
        if seg_off is None:
            return c
        off = np.asarray(seg_off)
        return torch.stack([c[off[i]:off[i + 1]].sum() for i in range(len(off) - 1)])


@pytest.fixture(scope="module")
def replay():
    g = np.load(os.path.join(GOLDEN, "hier_plan.npz"))
    hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd = synthetic_state_dict(hp, int(g["weight_seed"]))
    cost = OracleCost(sd)
    smp = S.ImageHierarchicalTreeCEMSampler(float("inf"), 200, 256, 0.3, n_level_hierarchy=8,
                                            sampling_rates_per_layer=[10, 10], subgoal_cost_fcn=cost, ll_cost_fcn=cost,
                                            n_ll_samples=5)
    trace = []
    T.CHOICE_HOOK = lambda costs, k: (trace.append((k, costs.copy())), k)[1]
    torch.set_num_threads(os.cpu_count())
    np.random.seed(int(g["np_seed"]))
    plans, costs, zs = [], [], []
    try:
        smp.init()
        for it in range(3):
            z = smp.sample(10)
            zs.append(z)
            end = H.injected_end_ind(it, z.shape[0])
            with torch.no_grad():
                ro = O.simulator_rollout(sd, g["state"], g["goal"], z, end)["predictions"]
            p, c = smp.optimize(ro, g["goal"])
            plans.append(p[0])
            costs.append(np.asarray(c, dtype=np.float64).reshape(-1))
            best = smp.sample(10)
        zs.append(best)
    finally:
        T.CHOICE_HOOK = None
    return g, smp, trace, plans, costs, zs


def test_proposals_match_reference_stream(replay):
    g, smp, _, _, _, zs = replay
    for i, z in enumerate(zs):
        assert z.shape[0] == int(g["call_sizes"][i])
        np.testing.assert_allclose(z[:, ::8, ::32], g["samples_sub_%d" % i], rtol=0, atol=1e-6)
    assert smp.fully_optimized == bool(g["fully_optimized"])


def test_decision_trace(replay):
    g, _, trace, _, _, _ = replay
    assert [k for k, _ in trace] == g["argmin_choice"].tolist()
    assert [len(c) for _, c in trace] == g["argmin_sizes"].tolist()
    mine = np.concatenate([c for _, c in trace])
    ref = g["argmin_costs"]
    assert np.array_equal(np.isnan(mine), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(mine[ok], ref[ok], rtol=2e-4, atol=2e-4)


def test_plans(replay):
    g, _, _, plans, costs, _ = replay
    for i in range(3):
        assert plans[i].shape == g["plan_%d" % i].shape
        np.testing.assert_allclose(plans[i], g["plan_%d" % i], rtol=0, atol=2e-5)
        np.testing.assert_allclose(costs[i], g["plan_cost_%d" % i], rtol=2e-4, atol=2e-4)
