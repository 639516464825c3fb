"""The reference's own CPU implementation of one CEM step, for `bench.py --impl reference` and its `cpu_baseline` leg.

MEASUREMENT INFRASTRUCTURE (never imported by the product package).  Runs the UNMODIFIED reference staged under
baseline/_ref (oracle/stage_reference.py) -- its TreeModel, GCPImageSimulator.rollout, L2ImageCost and FlatCEMSampler.fit,
i.e. the stock code path of gcp/planning/cem/cem_planner.py:55-69 -- on torch 2.x through the import shims of
oracle/refshim.py, with the same seeded synthetic weights as the product arm.  `available()` is False when the staged tree
is missing (then bench.py falls back to the oracle port, kind "port").
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(STAGED, "gcp")) and os.path.isdir(os.path.join(STAGED, "blox"))


class ReferenceCEM:
    def __init__(self, state_dict, threads=None):
        os.environ["GCP_REFERENCE_ROOT"] = STAGED
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle import refshim
        refshim.REFERENCE_ROOT = STAGED
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints at import; bench.py prints ONE line
            refshim.install()
            import torch
            from blox import AttrDict
            from experiments.prediction.base_configs import gcp_tree as base_conf
            from gcp.planning.cem import cost_fcn
            from gcp.planning.cem.cem_simulator import GCPImageSimulator
            from gcp.planning.cem.sampler import SimpleTreeCEMSampler
            from gcp.prediction.models.tree.tree import TreeModel
        torch.set_num_threads(threads or os.cpu_count())
        h = AttrDict(base_conf.model_config)          # experiments/control/25room/gcp_tree/mod_hyper.py:33-55
        h.update({
            'state_dim': 2, 'ngf': 16, 'max_seq_len': 200, 'hierarchy_levels': 8, 'nz_mid_lstm': 512,
            'n_lstm_layers': 3, 'nz_mid': 128, 'nz_enc': 128, 'nz_vae': 256, 'regress_length': True,
            'attach_state_regressor': True, 'attach_inv_mdl': True,
            'inv_mdl_params': AttrDict(n_actions=2, use_convs=False, build_encoder=False),
            'untied_layers': True, 'decoder_distribution': 'discrete_logistic_mixture', 'batch_size': 1,
        })
        h.pop("add_weighted_pixel_copy")
        with contextlib.redirect_stdout(io.StringIO()):
            model = TreeModel(h, None)
        model.device = torch.device("cpu")
        model._hp.device = model.device
        model.eval()
        model.load_state_dict({k: v for k, v in state_dict.items() if not k.startswith("cost_mdl.")}, strict=True)
        self.torch = torch
        self.sim = GCPImageSimulator(model, append_latent=True)      # L2ImageCost always splits the latent off (cost_fcn.py:33-39)
        self.cost = cost_fcn.L2ImageCost(True, 1.0)
        self.sampler = SimpleTreeCEMSampler(float("inf"), 200, 256, 0.3, n_level_hierarchy=8)

    def step(self, state, goal, n, elite_frac, seed):
        """One CEM iteration on n candidates: sample -> simulator.rollout -> cost -> argsort -> fit.  Returns seconds
        (the draw of the samples included, as in the product arm)."""
        np.random.seed(seed)
        self.torch.manual_seed(seed)
        self.sampler.init()
        t0 = time.perf_counter()
        samples = self.sampler.sample(n)
        with self.torch.no_grad():
            ro = self.sim.rollout(state, goal, samples, 200)
        scores = self.cost(ro.predictions, goal)
        elite = scores.argsort()[:max(int(n * elite_frac), 1)]
        self.sampler.fit(samples[elite], scores[elite])
        return time.perf_counter() - t0
