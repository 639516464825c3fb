"""Seeded inputs and case table shared by oracle/make_golden_optim.py (which needs the reference) and the tests (which
must not): parameter / gradient arrays and the optimiser settings of every golden trajectory.  TEST INFRASTRUCTURE."""
import numpy as np

SIZES, STEPS = (1, 37, 4099), 12
CASES = dict(
    radam=dict(kind="radam", lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, clip=None),
    radam_clip_wd=dict(kind="radam", lr=2e-3, betas=(0.9, 0.999), weight_decay=0.01, clip=1.0),
    adam=dict(kind="adam", lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, clip=None),
    adam_clip_wd=dict(kind="adam", lr=2e-3, betas=(0.5, 0.999), weight_decay=0.01, clip=0.5),
)


def data(seed=0):
    r = np.random.default_rng(seed)
    params = [r.normal(0, 1, n).astype(np.float32) for n in SIZES]
    grads = [[(r.normal(0, 1, n) * (0.1 + t)).astype(np.float32) for n in SIZES] for t in range(STEPS)]
    return params, grads
