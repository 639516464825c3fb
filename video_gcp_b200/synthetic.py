"""Deterministic synthetic weights / inputs of the 25-room shapes.

There is no checkpoint or dataset available offline, so tests, `smoke()` and `bench.py` use seeded
random-init weights of the reference architecture.  Values come from numpy's PCG64 keyed by
(seed, crc32(key)), so the same state dict can be rebuilt bit-identically on any box and loaded into
both the reference model and this package's model.  Scales keep activations O(1) through the 8 TreeLSTM
levels and the decoder (fan-in scaled uniform; BatchNorm running statistics and affine terms are
randomised too, so eval-mode BN folding is actually exercised).
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch

from . import spec


def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode())])


def _make(key, shape, kind, seed):
    r = _rng(seed, key)
    u = lambda lo, hi: r.uniform(lo, hi, size=shape).astype(np.float32)
    if kind == spec.W:
        if len(shape) == 4 and shape[2:] == (3, 3):      # 3x3 conv on a 1x1 map: only centre tap is live
            fan_in = shape[1]
        elif len(shape) == 3:                             # conv1d (training-only nets)
            fan_in = shape[1] * shape[2]
        elif len(shape) == 4 and "decoder.net.net.net.conv" in key:   # ConvTranspose2d [in,out,k,k]
            fan_in = shape[0]
        else:
            fan_in = int(np.prod(shape[1:]))
        a = np.sqrt(3.0 / fan_in)
        return u(-a, a)
    if kind in (spec.B_, spec.BN_B, spec.GN_B):
        return u(-0.1, 0.1)
    if kind in (spec.BN_W, spec.GN_W):
        return u(0.8, 1.2)
    if kind == spec.BN_RM:
        return u(-0.2, 0.2)
    if kind == spec.BN_RV:
        return u(0.5, 1.5)
    if kind == spec.BN_NBT:
        return np.zeros(shape, dtype=np.int64)
    if kind == spec.LSTM_W:
        a = 1.0 / np.sqrt(shape[1])
        return u(-a, a)
    if kind == spec.LSTM_B:
        b = u(-0.1, 0.1)
        n = shape[0]
        b[n // 4:n // 2] += 1.0                           # forget-gate bias 1 (recurrent_modules.py:156-162)
        return b
    if kind == spec.ZEROS:
        return np.zeros(shape, dtype=np.float32)
    if kind == spec.ONES:
        return np.ones(shape, dtype=np.float32)
    raise ValueError(kind)


def synthetic_state_dict(hp, seed=0, expand_aliases=True):
    """State dict (reference key names) of seeded synthetic weights; aliases share storage."""
    canon = spec.canonical_entries(hp)
    sd = OrderedDict()
    for key, (shape, kind) in canon.items():
        sd[key] = torch.from_numpy(_make(key, tuple(shape), kind, seed))
    if expand_aliases:
        for alias, target in spec.aliases(hp):
            for key in list(canon.keys()):
                if key.startswith(target + "."):
                    sd[alias + key[len(target):]] = sd[key]
    return sd


def synthetic_rollout_inputs(n_candidates, seed=0, z_std=1.0, shared_images=True, end_ind_range=(2, 200)):
    """Start/goal images in [-1,1] (NCHW), noise z [B,255,256] and injected end_ind, seeded.
    `shared_images`: all candidates see the same start/goal pair, as in a CEM call."""
    r = np.random.default_rng([int(seed), 12345])
    nimg = 1 if shared_images else n_candidates
    I_0 = r.uniform(-1, 1, size=(nimg, 3, 32, 32)).astype(np.float32)
    I_g = r.uniform(-1, 1, size=(nimg, 3, 32, 32)).astype(np.float32)
    if shared_images:
        I_0 = np.repeat(I_0, n_candidates, 0)
        I_g = np.repeat(I_g, n_candidates, 0)
    z = (r.standard_normal(size=(n_candidates, 255, 256)) * z_std).astype(np.float32)
    end_ind = r.integers(end_ind_range[0], end_ind_range[1], size=(n_candidates,)).astype(np.int64)
    return dict(I_0=torch.from_numpy(I_0), I_g=torch.from_numpy(I_g), z=torch.from_numpy(z),
                end_ind=torch.from_numpy(end_ind))


def synthetic_seq_inputs(n_candidates, seed=0, z_std=1.0, shared_images=True, n_steps=199):
    """Inputs of the sequential GCP rollout: start/goal images in [-1,1] and noise z [B,199,256] (one latent
    per predicted frame), seeded."""
    r = np.random.default_rng([int(seed), 54321])
    nimg = 1 if shared_images else n_candidates
    I_0 = r.uniform(-1, 1, size=(nimg, 3, 32, 32)).astype(np.float32)
    I_g = r.uniform(-1, 1, size=(nimg, 3, 32, 32)).astype(np.float32)
    if shared_images:
        I_0 = np.repeat(I_0, n_candidates, 0)
        I_g = np.repeat(I_g, n_candidates, 0)
    z = (r.standard_normal(size=(n_candidates, n_steps, 256)) * z_std).astype(np.float32)
    end_ind = r.integers(2, n_steps + 1, size=(n_candidates,)).astype(np.int64)
    return dict(I_0=torch.from_numpy(I_0), I_g=torch.from_numpy(I_g), z=torch.from_numpy(z),
                end_ind=torch.from_numpy(end_ind))


def synthetic_train_batch(batch_size, seed=0, end_ind=None, max_seq_len=200):
    """A training batch of the 25-room dataset_spec shape (BASELINE config 1): traj_seq uniform(-1,1)
    [B,T,3,32,32] zeroed past end_ind, end_ind ~ randint(10,T), pad_mask = (t <= end_ind), states N(0,1) [B,T,2],
    actions N(0,1) [B,T-1,2], I_0 = traj[:,0], I_g = traj[b,end_ind_b]; plus the posterior noise eps [B,255,256]
    (depth-first node order) and the numpy seed used for the auxiliary heads' pair sampling."""
    r = np.random.default_rng([int(seed), 777])
    T = max_seq_len
    traj = r.uniform(-1, 1, size=(batch_size, T, 3, 32, 32)).astype(np.float32)
    if end_ind is None:
        end_ind = r.integers(10, T, size=(batch_size,))
    end_ind = np.asarray(end_ind, dtype=np.int64)
    pad = (np.arange(T)[None] <= end_ind[:, None]).astype(np.float32)
    traj *= pad[:, :, None, None, None]
    states = r.standard_normal(size=(batch_size, T, 2)).astype(np.float32)
    actions = r.standard_normal(size=(batch_size, T - 1, 2)).astype(np.float32)
    eps = r.standard_normal(size=(batch_size, 255, 256)).astype(np.float32)
    t = torch.from_numpy
    traj_t = t(traj)
    return dict(traj_seq=traj_t, pad_mask=t(pad), end_ind=t(end_ind), states=t(states), actions=t(actions),
                I_0=traj_t[:, 0].clone(), I_g=traj_t[torch.arange(batch_size), t(end_ind)].clone(), eps=t(eps),
                np_seed=int(seed) + 1000)
