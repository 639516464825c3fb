"""CPU: the optimiser oracle (oracle/optim_oracle.py) against trajectories of the unmodified reference optimisers
(tests/golden/optim_steps.npz, oracle/make_golden_optim.py): 12 steps of RAdam / Adam with and without clipping and decay."""
import os

import numpy as np
import pytest

from oracle import optim_oracle as OO
from oracle.make_golden_optim_data import CASES, data


@pytest.mark.parametrize("name", sorted(CASES))
def test_trajectory_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "optim_steps.npz"))
    c = CASES[name]
    params, grads = data()
    ps, ms, vs = OO.run(c["kind"], params, grads, c["lr"], c["betas"], 1e-8, c["weight_decay"], c["clip"])
    for i in range(len(params)):
        for tag, got in (("p", ps[i]), ("m", ms[i]), ("v", vs[i])):
            ref = g["%s_%s%d" % (name, tag, i)]
            err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12)
            assert err < 2e-6, (name, tag, i, err)
    # the update is not a no-op
    assert max(np.abs(ps[i] - params[i]).max() for i in range(len(params))) > 1e-4 or name == "radam_clip_wd"
