"""Flat CEM planner (reference: gcp/planning/cem/cem_planner.py:15-135), device-resident.

Per iteration: sample candidates -> batched tree rollout -> per-candidate cost -> k lowest -> refit.
Only the elites' rollouts ever reach the host.  With torch.distributed initialised, candidates are
sharded over ranks: each rank rolls out its slice, ONE all-gather of the [N/R] fp32 costs gives every
rank the full cost vector, every rank runs the same top-k, and the elite noise is regenerated locally
from the shared counter-based RNG (no sample payload crosses NVLink).
"""
import copy

import numpy as np
import torch
import torch.distributed as dist

from ..types import AttrDict, ParamDict
from .cost_fcn import L2ImageCost, LearnedCostEstimate
from .sampler import FlatCEMSampler, ImageHierarchicalTreeCEMSampler


class CEMPlanner:
    def __init__(self, hp, simulator):
        self._hp = self._default_hparams().overwrite(hp)
        self._simulator = simulator
        self._cost_fcn = self._build_cost()
        self._sampler = self._build_sampler()
        self._sampler.attach(simulator._model.engine, seed=self._hp.seed)
        self._logs = []

    def _default_hparams(self):
        return ParamDict(
            horizon=None, action_dim=None, n_iters=1, batch_size=64, max_rollout_bs=100, elite_frac=0.1,
            cost_fcn=L2ImageCost, dense_cost=False, final_step_cost_weight=1.0,
            sampler=FlatCEMSampler, sampler_clip_val=float("Inf"), initial_std=3e-1,
            verbose=False, dump_planning_data=False, use_delta_state_actions=False, use_inferred_actions=True,
            max_seq_len=None, seed=0,
        )

    def _build_cost(self):
        return self._hp.cost_fcn(self._hp.dense_cost, self._hp.final_step_cost_weight)

    def _build_sampler(self):
        return self._hp.sampler(self._hp.sampler_clip_val, self._hp.max_seq_len, self._hp.action_dim, self._hp.initial_std)

    @property
    def append_latent(self):
        return getattr(self._sampler, "append_latent", False)

    # ---- sharding -----------------------------------------------------------------------------
    @staticmethod
    def _world():
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _rollout_costs(self, state, goal_state, z):
        """Rolls out device samples z in chunks of max_rollout_bs; returns (costs, chunks)."""
        bs = max(int(self._hp.max_rollout_bs), 1)
        costs, chunks = [], []
        for s in range(0, z.shape[0], bs):
            ro = self._simulator.rollout_device(state, goal_state, z[s:s + bs], self._hp.max_seq_len)
            costs.append(self._cost_fcn.device_cost(ro))
            chunks.append((s, ro))
        return torch.cat(costs), chunks

    def cem_iteration(self, state, goal_state):
        """One sharded CEM iteration on the device.  Returns (all costs [N], elite ids [k], elite costs)."""
        rank, world = self._world()
        N = int(self._hp.batch_size)
        assert N % world == 0, "batch_size must divide over ranks"
        n_loc = N // world
        first = rank * n_loc
        z = self._sampler.sample_device(n_loc, first_id=first)
        cost_loc, chunks = self._rollout_costs(state, goal_state, z)
        if world > 1:
            cost = torch.empty(N, device=cost_loc.device, dtype=torch.float32)
            dist.all_gather_into_tensor(cost, cost_loc.contiguous())
        else:
            cost = cost_loc
        k = max(int(N * self._hp.elite_frac), 1)
        idx, val = self._simulator._model.engine.topk(cost, k)
        if world > 1:
            z_elite = self._sampler.regenerate(idx)
            self._sampler.fit_device(z_elite, torch.arange(k, device=idx.device, dtype=torch.int32))
        else:
            z_elite = z[idx.long()]
            self._sampler.fit_device(z, idx)
        return cost, idx, val, z_elite, chunks

    def __call__(self, state, goal_state):
        self._sampler.init()
        logs = []
        z_elite = val = None
        for _ in range(self._hp.n_iters):
            cost, idx, val, z_elite, chunks = self.cem_iteration(state, goal_state)
            logs.append(AttrDict(elite_scores=val.cpu().numpy(), goal_state=goal_state))
        # final rollout of the elites with the best samples (cem_planner.py:81-96)
        ro = self._simulator.rollout_device(state, goal_state, z_elite, self._hp.max_seq_len)
        final = ro.to_host(self._simulator._append_latent)
        self._sampler.sync_host()
        logs.append(AttrDict(elite_rollouts=copy.deepcopy(self._maybe_split_image(final.predictions)),
                             elite_scores=val.cpu().numpy(), dists=self._sampler.get_dists(), goal_state=goal_state,
                             elite_states=copy.deepcopy(final.states)))
        self._logs.append(logs)
        best_actions = self._get_action_plan(final, z_elite)
        return final.predictions[0], best_actions[0], final.latents[0], float(val[0])

    def _maybe_split_image(self, rollout):
        if hasattr(self._cost_fcn, "_split_state_rollout") and self._simulator._append_latent:
            return self._cost_fcn._split_state_rollout(rollout).image_rollout
        return rollout

    def _get_action_plan(self, final_rollouts, best_samples):
        if self._hp.use_delta_state_actions:
            return [b[1:] - b[:-1] for b in final_rollouts.states]
        elif self._hp.use_inferred_actions:
            return final_rollouts.actions
        return best_samples.cpu().numpy()

    def log_verbose(self, logger, step, phase, i_tr, dump_dir):
        self._logs = []


class ImageCEMPlanner(CEMPlanner):
    pass


class HierarchicalCEMPlanner(CEMPlanner):
    """CEM planner for hierarchical optimisation (cem_planner.py:166-218 over CEMPlanner.__call__, :55-96): every
    iteration rolls out the current proposals, optimises one more layer of the latent tree (best-of-N by learned
    cost), and the final iteration's single optimised latent tree is rolled out as the plan.  Rollouts and all cost
    inputs stay on the device; per iteration only the [n] costs and the chosen plan frames reach the host."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self._hp.sampling_rates_per_layer is not None:
            assert self._hp.n_iters == len(self._hp.sampling_rates_per_layer) + 1

    def _default_hparams(self):
        d = super()._default_hparams()
        d.update(ParamDict(horizon=None, cost_fcn=LearnedCostEstimate, cost_config={}, LL_cost_fcn=None,
                           sampler=ImageHierarchicalTreeCEMSampler, n_level_hierarchy=None, sampling_rates_per_layer=None,
                           n_ll_samples=5, sampler_rng="numpy"))
        return d

    def _build_cost(self):
        # the learned cost network is the model's own cost_mdl.cost_pred (TestTimeCostModel loads those weights from
        # the same checkpoint, cost_mdl.py:126-136)
        cost_fcn = self._hp.cost_fcn(self._hp.cost_config, model=self._simulator._model)
        if self._hp.LL_cost_fcn is not None:
            raise NotImplementedError("a separate LL_cost_fcn is not on the 25-room planner path")
        self._ll_cost_fcn = cost_fcn
        return cost_fcn

    def _build_sampler(self):
        return self._hp.sampler(self._hp.sampler_clip_val, self._hp.max_seq_len, self._hp.action_dim, self._hp.initial_std,
                                n_level_hierarchy=self._hp.n_level_hierarchy,
                                sampling_rates_per_layer=self._hp.sampling_rates_per_layer,
                                subgoal_cost_fcn=self._cost_fcn, ll_cost_fcn=self._ll_cost_fcn,
                                n_ll_samples=self._hp.n_ll_samples, rng=self._hp.sampler_rng)

    def _rollout_device(self, state, goal_state, samples):
        """CEMPlanner._rollout (:115-122): chunks of max_rollout_bs, the remainder beyond the last full chunk is
        dropped; one device call when everything fits the engine."""
        bs = int(self._hp.max_rollout_bs)
        n = samples.shape[0]
        n_used = n if n <= bs else (n // bs) * bs
        return self._simulator.rollout_device(state, goal_state, samples[:n_used], self._hp.max_seq_len)

    def __call__(self, state, goal_state):
        self._sampler.init()
        logs = []
        best_samples = best_scores = None
        for it in range(self._hp.n_iters):
            samples = self._sampler.sample(self._hp.batch_size)
            ro = self._rollout_device(state, goal_state, samples)
            best_rollouts, best_scores = self._sampler.optimize(ro, goal_state)
            # cem_planner.py:214: the next proposals are drawn right after the optimisation step.  Only the last draw
            # is used (it is the fully optimised tree); the numpy stream needs the others to stay reference-identical
            if self._hp.sampler_rng == "numpy" or it == self._hp.n_iters - 1:
                best_samples = self._sampler.sample(self._hp.batch_size)
            logs.append(AttrDict(elite_rollouts=copy.deepcopy(best_rollouts), elite_scores=best_scores,
                                 dists=self._sampler.get_dists(), goal_state=goal_state))
        final = self._rollout_device(state, goal_state, best_samples).to_host(self._simulator._append_latent)
        logs.append(AttrDict(elite_rollouts=copy.deepcopy(self._maybe_split_image(final.predictions)),
                             elite_scores=best_scores, dists=self._sampler.get_dists(), goal_state=goal_state,
                             elite_states=copy.deepcopy(final.states)))
        self._logs.append(logs)
        best_actions = self._get_action_plan(final, best_samples)
        return final.predictions[0], best_actions[0], final.latents[0], best_scores[0]


class HierarchicalImageCEMPlanner(HierarchicalCEMPlanner, ImageCEMPlanner):
    pass
