"""CEM samplers (reference: gcp/planning/cem/sampler.py:7-80).

Same interface (init / sample / fit / get_dists); state lives on the device when an engine is attached:
`sample_device(n, first_id)` draws Philox noise keyed by the GLOBAL candidate id, so any rank of a sharded
CEM call can regenerate any candidate (used for the elite refit without moving samples between GPUs).
"""
import numpy as np
import torch

from ..types import AttrDict
from .tree_optimizer import ImageHierarchicalTreeLatentOptimizer


class CEMSampler:
    def __init__(self, clip_val, n_steps, action_dim, initial_std):
        self._clip_val = clip_val
        self._n_steps = n_steps
        self._action_dim = action_dim
        self._initial_std = initial_std
        self.engine = None
        self.seed = 0
        self._draws = 0          # device draws so far: never reset, so no two draws of a sampler share a noise stream
        self._last_seed = 0
        self._last_dist = (None, None)
        self.init()

    def attach(self, engine, seed=0):
        self.engine = engine
        self.seed = int(seed)
        return self


class FlatCEMSampler(CEMSampler):
    """Samples flat arrays from per-element Gaussians."""

    def init(self):
        self.mean = np.zeros((self._n_steps, self._action_dim))
        self.std = self._initial_std * np.ones((self._n_steps, self._action_dim))
        self._mean_d = self._std_d = None

    # ---- reference contract (host numpy) ----
    def sample(self, n_samples):
        raw = np.random.normal(loc=self.mean, scale=self.std, size=(n_samples, self._n_steps, self._action_dim))
        return np.clip(raw, -self._clip_val, self._clip_val)

    def fit(self, data, scores):
        if isinstance(data, torch.Tensor):
            raise TypeError("use fit_device for device tensors")
        self.mean = np.mean(data, axis=0)
        self.std = np.std(data, axis=0)
        self._mean_d = self._std_d = None

    def get_dists(self):
        self.sync_host()
        return AttrDict(mean=self.mean, std=self.std)

    # ---- device path ----
    def _next_seed(self):
        """Philox key of the next device draw.  The reference draws from the advancing np.random stream, so every CEM
        iteration AND every replan sees fresh samples (sampler.py:40-42); here the draw counter plays that role: it
        advances with every sample_device() call and init() does not reset it.  Ranks of a sharded planner make the
        same sequence of calls, so they agree on the key without communicating."""
        seed = (self.seed * 1000003 + self._draws) & 0xFFFFFFFFFFFFFFFF
        self._draws += 1
        self._last_seed = seed
        self._last_dist = (getattr(self, "_mean_d", None), getattr(self, "_std_d", None))    # distribution of this draw
        return seed

    def sample_device(self, n_samples, first_id=0, out=None):
        assert self.engine is not None, "attach(engine) first"
        return self.engine.sample_noise(n_samples, self._mean_d, self._std_d, float(self._initial_std), self._next_seed(),
                                        first_id, self._clip_val, out=out)

    def regenerate(self, ids):
        """Noise of the given global candidate ids (int32 device tensor) of the LATEST draw (same key, same
        distribution): bit-identical to the rows sample_device produced for those ids on whichever rank owned them."""
        mean, std = self._last_dist
        return self.engine.sample_noise_ids(ids.int(), mean, std, float(self._initial_std), self._last_seed, self._clip_val)

    def fit_device(self, z, elite_idx):
        """Refit from rows `elite_idx` (int32 device tensor) of device samples z."""
        self._mean_d, self._std_d = self.engine.refit(z, elite_idx)

    def sync_host(self):
        """Brings the host copies of mean / std up to date with the device refit (lazily: get_dists() calls it)."""
        if self._mean_d is not None and getattr(self, "_synced", None) is not self._mean_d:
            self.mean = self._mean_d.double().cpu().numpy()
            self.std = self._std_d.double().cpu().numpy()
            self._synced = self._mean_d


class SimpleTreeCEMSampler(FlatCEMSampler):
    """Flat CEM over all 2^n - 1 tree nodes at once (sampler.py:74-80)."""

    def __init__(self, *args, n_level_hierarchy, **kwargs):
        self._n_layer_hierarchy = n_level_hierarchy
        super().__init__(*args)
        self._n_steps = 2 ** n_level_hierarchy - 1
        self.init()


class ImageHierarchicalTreeCEMSampler(SimpleTreeCEMSampler):
    """Tree-GCP sampler that optimises the layers of the hierarchy sequentially, starting from the top, for image
    prediction GCPs (sampler.py:83-143).  `rng`: "numpy" reproduces the reference's np.random stream, "device" draws
    from the engine's Philox stream and keeps the proposals in HBM."""

    def __init__(self, *args, sampling_rates_per_layer, subgoal_cost_fcn, ll_cost_fcn, n_ll_samples, rng="numpy", **kwargs):
        self._sampling_rates_per_layer = sampling_rates_per_layer
        self._subgoal_cost_fcn = subgoal_cost_fcn
        self._ll_cost_fcn = ll_cost_fcn
        self._n_ll_samples = n_ll_samples
        self._rng = rng
        self._optimizer = None
        super().__init__(*args, **kwargs)
        assert self._n_layer_hierarchy >= len(sampling_rates_per_layer)     # not enough layers in tree

    def init(self):
        if getattr(self, "_n_layer_hierarchy", None) is None or getattr(self, "_ll_cost_fcn", None) is None:
            return      # base-class constructor calls init() before the tree parameters exist
        self._optimizer = ImageHierarchicalTreeLatentOptimizer(
            self._action_dim, list(self._sampling_rates_per_layer), self._n_layer_hierarchy, self._subgoal_cost_fcn,
            self._ll_cost_fcn, self._n_ll_samples, engine=self.engine, rng=self._rng, seed=self._next_seed())

    def sample(self, n_samples):
        z = self._optimizer.sample()
        if isinstance(z, torch.Tensor):
            return z.clamp(-self._clip_val, self._clip_val).contiguous()
        return np.clip(z, -self._clip_val, self._clip_val)

    def optimize(self, rollouts, goal):
        best_rollout, best_cost = self._optimizer.optimize(rollouts, goal)
        goal = np.asarray(goal, dtype=np.float32)
        if (best_rollout[-1] != goal[0].transpose(2, 0, 1)).any():    # can happen if too few frames on right tree side
            best_rollout = np.concatenate((best_rollout, goal.transpose(0, 3, 1, 2)))
        if not hasattr(best_cost, "__len__"):
            best_cost = [best_cost]
        return [best_rollout], best_cost

    def fit(*args, **kwargs):
        """Does not support refitting distributions (sampler.py:113-115)."""

    def get_dists(self):
        return AttrDict(mean=0., std=1.)

    def sync_host(self):
        pass

    @property
    def append_latent(self):
        return True     # latent rollouts are needed to compute subgoal costs

    @property
    def fully_optimized(self):
        return self._optimizer.fully_optimized


HierarchicalTreeCEMSampler = ImageHierarchicalTreeCEMSampler    # only the image variant is on the 25-room path
