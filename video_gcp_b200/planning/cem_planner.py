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

from .. import dist_utils
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
        # max_rollout_bs: the reference default is 100 (cem_planner.py:31); the device pads a rollout to 128-row tiles,
        # so chunks that are multiples of 128 waste nothing -- one chunk of up to 1024 candidates by default
        return ParamDict(
            horizon=None, action_dim=None, n_iters=1, batch_size=64, max_rollout_bs=1024, elite_frac=0.1,
            cost_fcn=L2ImageCost, dense_cost=False, final_step_cost_weight=1.0,
            sampler=FlatCEMSampler, sampler_clip_val=float("Inf"), initial_std=3e-1,
            verbose=False, dump_planning_data=False, use_delta_state_actions=False, use_inferred_actions=True,
            max_seq_len=None, seed=0,
            # True: roll out in planner mode -- only the nodes balanced pruning keeps are decoded (the planner never reads
            # the others) and an L2 image cost is reduced inside the decoder, so CEM iterations write no images at all;
            # costs / elites / plans are identical to False, which decodes all 255 nodes of every candidate as the
            # reference does (cem_simulator.py:29-61 then discards them)
            prune_before_decode=True,
            # True: the sampled rollout lengths of a call are handed to its candidates in descending order (they are i.i.d.
            # draws from one distribution, independent of the noise, so nothing changes statistically); candidate tiles then
            # share their length class and the pruned tree recursion skips most of the deep levels.  None = prune_before_decode
            sort_lengths=None,
        )

    def _build_cost(self):
        return self._hp.cost_fcn(self._hp.dense_cost, self._hp.final_step_cost_weight)

    def _planner_mode(self, images, l2_out=None):
        """rollout_device kwargs for a rollout whose frames are / are not needed beyond the cost.  Either way an L2 image
        cost is reduced inside the decoder (same bits in both modes); prune_before_decode decides whether the decoder
        visits the kept nodes only (and skips the image writes the caller does not need) or all 255 nodes."""
        spec = self._cost_fcn.fused_spec() if hasattr(self._cost_fcn, "fused_spec") else None
        sort = bool(self._hp.prune_before_decode if self._hp.sort_lengths is None else self._hp.sort_lengths)
        if not self._hp.prune_before_decode:
            return dict(planner_mode=dict(kept_only=False, images=True, l2=spec, l2_out=l2_out, sort_lengths=sort))
        latent_cost = hasattr(self._cost_fcn, "pairs_device")       # learned latent-space cost: reads no image at all
        # CEM iterations of an image-space cost read nothing but the cost: the existence / inverse-model / state heads are
        # left to the final rollout of the elites (which returns actions and states); a latent-space cost needs the pruned
        # latent sequence the heads' path produces
        return dict(planner_mode=dict(kept_only=True, images=images or (spec is None and not latent_cost), l2=spec, l2_out=l2_out,
                                      heads=images or latent_cost or spec is None, sort_lengths=sort))

    def _build_sampler(self):
        return self._hp.sampler(self._hp.sampler_clip_val, self._hp.max_seq_len, self._hp.action_dim, self._hp.initial_std)

    @property
    def append_latent(self):
        return getattr(self._sampler, "append_latent", False)

    @property
    def engine(self):
        return self._simulator._model.engine

    def _chunks(self, n):
        """Chunk boundaries of CEMPlanner._rollout (cem_planner.py:115-122): max(n // max_rollout_bs, 1) chunks of
        max_rollout_bs -- the remainder beyond the last full chunk is dropped, one short chunk when n < max_rollout_bs."""
        bs = max(int(self._hp.max_rollout_bs), 1)
        return [(i * bs, min((i + 1) * bs, n)) for i in range(max(n // bs, 1))]

    def _rollout_costs(self, state, goal_state, z):
        """Rolls out samples z (device tensor, or pinned host tensor) chunk by chunk; returns the [n_used] cost vector.
        Every chunk's costs are written into their own slice of one preallocated vector: the rollout / cost buffers of
        the engine are persistent and reused by the next chunk, so nothing else of a chunk is kept."""
        chunks = self._chunks(z.shape[0])
        cost = torch.empty(chunks[-1][1], device=self.engine.device, dtype=torch.float32)
        if len(chunks) > 1 and not getattr(self._cost_fcn, "chunkable", True):
            raise NotImplementedError("%s needs all candidates in one rollout (raise max_rollout_bs)" % type(self._cost_fcn).__name__)
        z_dev = []
        for s, e in chunks:
            ro = self._simulator.rollout_device(state, goal_state, z[s:e], self._hp.max_seq_len,
                                                **self._planner_mode(images=False, l2_out=cost[s:e]))
            self._cost_fcn.device_cost(ro, out=cost[s:e])
            if not z.is_cuda:
                z_dev.append(ro.z if len(chunks) == 1 else ro.z.clone())    # device copy of host-resident samples
        return cost, (z if z.is_cuda else (z_dev[0] if len(z_dev) == 1 else torch.cat(z_dev)))

    def cem_iteration(self, state, goal_state, samples=None):
        """One CEM iteration on the device (cem_planner.py:58-69): sample -> rollout -> cost -> elites -> refit.
        With torch.distributed initialised the N candidates are sharded over the ranks (global candidate ids
        [rank * N/R, (rank+1) * N/R)); one all-gather of the costs, the same top-k on every rank, elites regenerated from
        the shared counter-based noise stream.  `samples` (optional, single rank only): this iteration's candidates,
        [N,255,256] on the device or in pinned host memory (the reference simulator's contract), instead of a draw.
        Returns (costs [N_used], elite ids [k] int32, elite costs [k], elite samples [k,255,256])."""
        rank, world = dist_utils.world()
        N = int(self._hp.batch_size)
        first, last = dist_utils.shard_range(N, rank, world)
        if samples is not None:
            if world > 1:
                raise ValueError("injected samples are not sharded; use the sampler's counter-based stream")
            z = samples
        else:
            z = self._sampler.sample_device(last - first, first_id=first)
        cost_loc, z = self._rollout_costs(state, goal_state, z)
        if world > 1 and cost_loc.shape[0] != last - first:
            raise ValueError("sharded CEM needs batch_size / world_size to be a multiple of max_rollout_bs (or smaller)")
        cost = dist_utils.gather_costs(cost_loc)
        k = max(int(N * self._hp.elite_frac), 1)
        idx, val = self.engine.topk(cost, k)
        if world > 1:
            z_elite = self._sampler.regenerate(idx)
            self._sampler.fit_device(z_elite, torch.arange(k, device=idx.device, dtype=torch.int32))
        else:
            z_elite = None
            self._sampler.fit_device(z, idx)
        return cost, idx, val, (z, z_elite)

    def _elite_samples(self, packed, idx):
        z, z_elite = packed
        return z_elite if z_elite is not None else z[idx.long()]

    def _rollout_host(self, state, goal_state, z, only_best=False):
        """Final rollout of the best samples (cem_planner.py:81-82), joined over chunks, as host lists.  only_best: all
        elites are rolled out, but only elite 0 -- the plan -- is copied to the host."""
        out = None
        for s, e in self._chunks(z.shape[0]):
            ro = self._simulator.rollout_device(state, goal_state, z[s:e], self._hp.max_seq_len, **self._planner_mode(images=True))
            if only_best and s > 0:
                continue
            part = ro.to_host(self._simulator._append_latent, idx=[0] if only_best else None)
            if out is None:
                out = part
            else:
                for key in out:
                    out[key] = out[key] + part[key]
        return out

    def __call__(self, state, goal_state):
        self._sampler.init()
        logs = []
        best = val = None
        # the reference logs every elite rollout of every call (cem_planner.py:71-89); that is 100+ MB of frames per plan,
        # so they are only brought to the host when something will read them (log_verbose with verbose / dump_planning_data)
        full_logs = bool(self._hp.verbose or self._hp.dump_planning_data)
        for _ in range(self._hp.n_iters):
            cost, idx, val, packed = self.cem_iteration(state, goal_state)
            best = self._elite_samples(packed, idx)
            logs.append(AttrDict(elite_scores=val if not full_logs else val.cpu().numpy(), goal_state=goal_state))
        # final rollout of the elites with the best samples (cem_planner.py:81-96)
        final = self._rollout_host(state, goal_state, best, only_best=not full_logs)
        scores = val.cpu().numpy()
        if full_logs:
            self._sampler.sync_host()
            logs.append(AttrDict(elite_rollouts=self._maybe_split_image(final.predictions), elite_scores=scores,
                                 dists=self._sampler.get_dists(), goal_state=goal_state, elite_states=final.states))
            self._logs.append(logs)
        best_actions = self._get_action_plan(final, best)
        return final.predictions[0], best_actions[0], final.latents[0], float(scores[0])

    def _maybe_split_image(self, rollout):
        if hasattr(self._cost_fcn, "_split_state_rollout") and self._simulator._append_latent:
            return self._cost_fcn._split_state_rollout(rollout).image_rollout
        return rollout

    def _get_action_plan(self, final_rollouts, best_samples):
        if self._hp.use_delta_state_actions:
            return [b[1:] - b[:-1] for b in final_rollouts.states]
        elif self._hp.use_inferred_actions:
            return final_rollouts.actions
        return best_samples[:1].cpu().numpy()

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
