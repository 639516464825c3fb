"""Import shims that let the UNMODIFIED reference tree (/root/reference) run on torch 2.x / py3.12.

TEST INFRASTRUCTURE ONLY.  This file is used by `oracle/make_golden.py` (run in the build container,
where /root/reference exists) to generate the golden fixtures under `tests/golden/`.  It never runs on
the GPU box and nothing in the product package imports it.

What it installs (SURVEY.md section 8c lists why each is needed):
  * fake modules: tensorflow.contrib.training.HParams, funcsigs, imp, matplotlib, tensorboardX,
    skimage, dload, h5py, moviepy, imageio, gym  (all trivial stubs; none is on the rollout path)
  * np.int / np.float aliases (gcp/prediction/utils/tree_utils.py:225)
  * ONE semantic patch: BalancedBinding.comp_timestep -> truncating integer division, restoring the
    torch-1.3 behaviour of `(t_l + t_r) / 2` on int64 tensors
    (gcp/prediction/models/tree/frame_binding.py:52-54; requirements.txt pins torch==1.3.0).
"""
import importlib.machinery
import importlib.util
import inspect
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("GCP_REFERENCE_ROOT", "/root/reference")


class HParams:
    """Minimal stand-in for tensorflow.contrib.training.HParams."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_hparam_types", {})
        for k, v in kwargs.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._hparam_types and getattr(self, name, None) is not None:
            raise ValueError("Hyperparameter name is reserved: %s" % name)
        self._hparam_types[name] = type(value)
        object.__setattr__(self, name, value)

    def set_hparam(self, name, value):
        if name not in self._hparam_types:
            raise KeyError(name)
        object.__setattr__(self, name, value)

    def values(self):
        return {k: getattr(self, k) for k in self._hparam_types if hasattr(self, k)}

    def __contains__(self, name):
        return name in self._hparam_types

    def get(self, name, default=None):
        return getattr(self, name, default) if name in self._hparam_types else default


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Object that swallows any attribute access / call (for plotting & logging stubs)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _lenient(name):
    """module-level __getattr__ for stub modules: anything goes except dunder probes."""
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything()


def _load_source(name, path):
    loader = importlib.machinery.SourceFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    loader.exec_module(module)
    return module


_installed = False


def install():
    """Install the shims and put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not found at %s (it only exists in the build container)"
                           % REFERENCE_ROOT)
    import torch  # noqa: F401  (must be fully imported before the stub modules exist)
    import torchvision  # noqa: F401
    os.environ.setdefault("GCP_DATA_DIR", "/tmp/gcp_data")
    os.environ.setdefault("GCP_EXP_DIR", "/tmp/gcp_exp")

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "bool"):
        np.bool = bool

    tf = _mod("tensorflow")
    contrib = _mod("tensorflow.contrib")
    training = _mod("tensorflow.contrib.training", HParams=HParams)
    tf.contrib = contrib
    contrib.training = training

    _mod("funcsigs", signature=inspect.signature, Parameter=inspect.Parameter)
    _mod("imp", load_source=_load_source)

    mpl = _mod("matplotlib", use=lambda *a, **k: None)
    mpl.cm = _mod("matplotlib.cm", get_cmap=_Anything())
    mpl.pyplot = _mod("matplotlib.pyplot")
    mpl.pyplot.__getattr__ = _lenient
    mpl.patches = _mod("matplotlib.patches")
    mpl.patches.__getattr__ = _lenient
    mpl.backends = _mod("matplotlib.backends")
    mpl.backends.backend_agg = _mod("matplotlib.backends.backend_agg", FigureCanvasAgg=_Anything)
    mpl.figure = _mod("matplotlib.figure", Figure=_Anything)

    _mod("tensorboardX", SummaryWriter=_Anything)
    sk = _mod("skimage")
    sk.io = _mod("skimage.io", imsave=lambda *a, **k: None)
    sk.transform = _mod("skimage.transform", resize=_Anything())
    sk.measure = _mod("skimage.measure", compare_ssim=_Anything(), compare_psnr=_Anything())
    _mod("dload")
    _mod("h5py", File=_Anything)
    mp = _mod("moviepy")
    mp.editor = _mod("moviepy.editor")
    mp.editor.__getattr__ = _lenient
    mp.audio = _mod("moviepy.audio")
    mp.audio.AudioClip = _mod("moviepy.audio.AudioClip", AudioArrayClip=_Anything)
    _mod("imageio", mimsave=lambda *a, **k: None, imwrite=lambda *a, **k: None)
    gym = _mod("gym")
    gym.__getattr__ = _lenient

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # the one semantic patch: integer midpoint with C-style truncation (torch 1.3 semantics)
    import torch
    from gcp.prediction.models.tree import frame_binding

    def comp_timestep(t_l, t_r, *unused_args):
        s = t_l + t_r
        if s.is_floating_point():
            return s / 2
        return torch.div(s, 2, rounding_mode="trunc")

    frame_binding.BalancedBinding.comp_timestep = staticmethod(comp_timestep)
    _installed = True
