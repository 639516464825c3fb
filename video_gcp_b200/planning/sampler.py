"""CEM samplers (reference: gcp/planning/cem/sampler.py:7-80).

Same interface (init / sample / fit / get_dists); state lives on the device when an engine is attached:
`sample_device(n, first_id)` draws Philox noise keyed by the GLOBAL candidate id, so any rank of a sharded
CEM call can regenerate any candidate (used for the elite refit without moving samples between GPUs).
"""
import numpy as np
import torch

from ..types import AttrDict


class CEMSampler:
    def __init__(self, clip_val, n_steps, action_dim, initial_std):
        self._clip_val = clip_val
        self._n_steps = n_steps
        self._action_dim = action_dim
        self._initial_std = initial_std
        self.engine = None
        self.seed = 0
        self._iter = 0
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
        self._iter = 0

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
        return AttrDict(mean=self.mean, std=self.std)

    # ---- device path ----
    def _iter_seed(self):
        return (self.seed * 1000003 + self._iter) & 0xFFFFFFFFFFFFFFFF

    def sample_device(self, n_samples, first_id=0, out=None):
        assert self.engine is not None, "attach(engine) first"
        return self.engine.sample_noise(n_samples, self._mean_d, self._std_d, float(self._initial_std), self._iter_seed(),
                                        first_id, self._clip_val, out=out)

    def regenerate(self, ids):
        """Noise of the given global candidate ids (int32 cuda tensor), this iteration's distribution."""
        return self.engine.sample_noise_ids(ids.int(), self._mean_d, self._std_d, float(self._initial_std),
                                            self._iter_seed(), self._clip_val)

    def fit_device(self, z, elite_idx):
        """Refit from rows `elite_idx` (int32 cuda) of device samples z."""
        self._mean_d, self._std_d = self.engine.refit(z, elite_idx)
        self._iter += 1

    def sync_host(self):
        if self._mean_d is not None:
            self.mean = self._mean_d.double().cpu().numpy()
            self.std = self._std_d.double().cpu().numpy()


class SimpleTreeCEMSampler(FlatCEMSampler):
    """Flat CEM over all 2^n - 1 tree nodes at once (sampler.py:74-80)."""

    def __init__(self, *args, n_level_hierarchy, **kwargs):
        self._n_layer_hierarchy = n_level_hierarchy
        super().__init__(*args)
        self._n_steps = 2 ** n_level_hierarchy - 1
        self.init()
