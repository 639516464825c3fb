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
from .cost_fcn import L2ImageCost
from .sampler import FlatCEMSampler


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
