"""Golden trajectories of the reference's optimiser, produced by RUNNING THE UNMODIFIED REFERENCE classes:
    python -m oracle.make_golden_optim        (build container only)

TEST INFRASTRUCTURE.  The two calls of `ClipGradOptimizer.step` (blox/torch/training.py:146-161) on `RAdam`
(blox/torch/radam.py) and `torch.optim.Adam`, the optimisers gcp_builder.py:174-186 builds, on three seeded parameter tensors (sizes 1, 37, 4099: a scalar, an
odd length, a length that is not a multiple of 4) for 12 steps -- RAdam's first five steps take the un-rectified branch
-- with and without gradient clipping / weight decay.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox.torch.radam import RAdam  # noqa: E402

from oracle.make_golden import GOLDEN  # noqa: E402

from oracle.make_golden_optim_data import CASES, STEPS, data  # noqa: E402

CLASSES = dict(radam=RAdam, adam=torch.optim.Adam)


def main():
    params, grads = data()
    out = {}
    for name, c in CASES.items():
        ps = [torch.nn.Parameter(torch.tensor(p)) for p in params]
        # ClipGradOptimizer.step (blox/torch/training.py:154-159) = clip_grad_norm_ over all parameters, then the wrapped
        # optimiser's step.  Its `np.concatenate([group['params'] ...])` is an old-numpy idiom (object array of
        # Parameters) that modern numpy refuses, so the two calls it makes are issued directly on the unmodified classes.
        opt = CLASSES[c["kind"]](ps, lr=c["lr"], betas=c["betas"], weight_decay=c["weight_decay"])
        for t in range(STEPS):
            for p, g in zip(ps, grads[t]):
                p.grad = torch.tensor(g)
            if c["clip"] is not None:
                torch.nn.utils.clip_grad_norm_(ps, c["clip"])
            opt.step()
        for i, p in enumerate(ps):
            out["%s_p%d" % (name, i)] = p.detach().numpy().copy()
            st = opt.state[p]
            out["%s_m%d" % (name, i)] = st["exp_avg"].numpy().copy()
            out["%s_v%d" % (name, i)] = st["exp_avg_sq"].numpy().copy()
        print(name, [float(np.abs(out["%s_p%d" % (name, i)] - params[i]).max()) for i in range(len(params))])
    np.savez_compressed(os.path.join(GOLDEN, "optim_steps.npz"), **out)


if __name__ == "__main__":
    main()
