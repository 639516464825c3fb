"""Golden fixture of the hierarchical planner, made by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_hier
TEST INFRASTRUCTURE.  Output: tests/golden/hier_plan.npz.

What runs: HierarchicalImageCEMPlanner (gcp/planning/cem/cem_planner.py:166-218) with
ImageHierarchicalTreeCEMSampler (cem/sampler.py:130-143), ImageHierarchicalTreeLatentOptimizer
(gcp/planning/tree_optimizer.py:164-190) and ImageLearnedCostEstimate (cem/cost_fcn.py:79-105) over the reference
TreeModel + GCPImageSimulator, with the 25-room control settings (experiments/control/25room/gcp_tree/mod_hyper.py:
n_iters 3, batch_size 10, sampling_rates_per_layer [10, 10], n_ll_samples 5).

Determinism: synthetic weights (seed 1) incl. cost_mdl.*; np.random.seed(SEED) right before the planner call (the
optimiser draws with np.random.normal); the sampled rollout length is replaced by `hier_oracle.injected_end_ind`
(a fixed function of the rollout-call index), which deliberately contains short rollouts so that the optimiser's
dummy-sequence / too-short branches run.

One shim beyond oracle/refshim.py, needed only because numpy >= 1.24 refuses ragged arrays:
np.array_split(list_of_rollouts, 1) in tree_optimizer.py:128 is given the old object-array behaviour ([the list]).
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
import torch  # noqa: E402
from blox import AttrDict as RefAttrDict  # noqa: E402
from gcp.planning import tree_optimizer as ref_opt  # noqa: E402
from gcp.planning.cem.cem_planner import HierarchicalImageCEMPlanner  # noqa: E402
from gcp.planning.cem.cem_simulator import GCPImageSimulator as RefSimulator  # noqa: E402
from gcp.planning.cem.cost_fcn import ImageLearnedCostEstimate  # noqa: E402
from gcp.planning.cem.sampler import ImageHierarchicalTreeCEMSampler  # noqa: E402

from oracle.hier_oracle import injected_end_ind  # noqa: E402
from oracle.make_golden import build_ref_model, GOLDEN, WEIGHT_SEED  # noqa: E402
from video_gcp_b200 import hparams as my_hparams  # noqa: E402
from video_gcp_b200.synthetic import synthetic_state_dict  # noqa: E402

SEED = 2024


class _NumpyProxy:
    """`np` as tree_optimizer.py sees it: logs argmin calls, tolerates ragged array_split."""

    def __init__(self, log):
        self._log = log

    def __getattr__(self, name):
        return getattr(np, name)

    def argmin(self, a, *args, **kw):
        k = np.argmin(a, *args, **kw)
        self._log.append((int(k), np.asarray(a, dtype=np.float64).reshape(-1).copy()))
        return k

    def array_split(self, ary, n):
        if isinstance(ary, list):
            bounds = np.linspace(0, len(ary), n + 1).astype(int)
            assert len(ary) % n == 0
            return [ary[bounds[i]:bounds[i + 1]] for i in range(n)]
        return np.array_split(ary, n)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    hp = my_hparams.build_hparams(my_hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
    sd_full = synthetic_state_dict(hp, WEIGHT_SEED)
    ref = build_ref_model()
    ref.load_state_dict({k: v for k, v in sd_full.items() if not k.startswith("cost_mdl.")}, strict=True)

    calls = []
    orig = ref.get_end_ind

    def patched(inputs, outputs):
        orig(inputs, outputs)
        e = torch.as_tensor(injected_end_ind(len(calls), inputs.I_0.shape[0]))
        calls.append(e.numpy().copy())
        outputs.end_ind = e
        return e

    ref.get_end_ind = patched

    tmp = tempfile.mkdtemp()
    torch.save({'epoch': 0, 'global_step': 0, 'state_dict': sd_full, 'optimizer': {}}, os.path.join(tmp, "weights_ep0.pth"))
    sim = RefSimulator(ref, append_latent=True)
    recorded = []
    sim_rollout = sim.rollout

    def rec_rollout(state, goal_state, samples, rollout_len, prune=False):
        out = sim_rollout(state, goal_state, samples, rollout_len, prune)
        recorded.append((np.array(samples), out))
        return out

    sim.rollout = rec_rollout
    cem_params = RefAttrDict(
        prune_final=True, horizon=200, action_dim=256, verbose=False, n_iters=3, batch_size=10, n_level_hierarchy=8,
        sampler=ImageHierarchicalTreeCEMSampler, sampling_rates_per_layer=[10, 10], cost_fcn=ImageLearnedCostEstimate,
        cost_config=RefAttrDict(checkpt_path=tmp), max_seq_len=200)
    with contextlib.redirect_stdout(io.StringIO()):
        planner = HierarchicalImageCEMPlanner(cem_params, sim)

    r = np.random.default_rng(11)
    state = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    goal = r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
    argmins = []
    ref_opt.np = _NumpyProxy(argmins)
    np.random.seed(SEED)
    try:
        with torch.no_grad():
            pred0, act0, lat0, score0 = planner(state, goal)
    finally:
        ref_opt.np = np
    logs = planner._logs[-1]
    out = dict(weight_seed=WEIGHT_SEED, np_seed=SEED, state=state, goal=goal,
               n_calls=len(recorded), end_inds=np.concatenate(calls), call_sizes=np.array([len(c) for c in calls]),
               n_argmin=len(argmins), argmin_choice=np.array([k for k, _ in argmins]),
               argmin_sizes=np.array([len(c) for _, c in argmins]),
               argmin_costs=np.concatenate([c for _, c in argmins]),
               final_pred_f16=pred0.astype(np.float16), final_pred_sum=np.float64(pred0.astype(np.float64).sum()),
               final_actions=act0, final_latents=lat0, final_score=np.asarray(score0, dtype=np.float64).reshape(-1),
               fully_optimized=planner._sampler.fully_optimized)
    # closed-loop execution step on the plan (ImageCEMPolicy._infer_action, planner_policy.py:215-221), reference code
    # path: encoder on the current image, inverse model against the next latent of the plan
    cur = r.uniform(0, 255, size=(3, 1, 32, 32, 3)).astype(np.float32)
    cl_act, cl_enc = [], []
    with torch.no_grad():
        for i in range(3):
            img = torch.tensor(cur[i], dtype=torch.float32)
            enc = ref.encoder(RefSimulator._env2planner(img))[0][:, :, 0, 0]
            a = ref.inv_mdl.run_single(enc, torch.tensor(lat0[i + 1][None]))[0]
            cl_act.append(a.numpy())
            cl_enc.append(enc[0].numpy())
    out.update(cl_images=cur, cl_actions=np.stack(cl_act), cl_enc=np.stack(cl_enc))
    for i, (samples, ro) in enumerate(recorded):
        out["samples_sub_%d" % i] = samples[:, ::8, ::32].astype(np.float32)
        out["samples_sum_%d" % i] = samples.sum((1, 2))
        out["pred_len_%d" % i] = np.array([p.shape[0] for p in ro.predictions])
        out["pred_sum_%d" % i] = np.array([p.astype(np.float64).sum() for p in ro.predictions])
    for i in range(3):
        out["plan_%d" % i] = np.asarray(logs[i].elite_rollouts[0], dtype=np.float32)
        out["plan_cost_%d" % i] = np.asarray(logs[i].elite_scores, dtype=np.float64).reshape(-1)
    np.savez_compressed(os.path.join(GOLDEN, "hier_plan.npz"), **out)
    print("hier plan: calls", [len(c) for c in calls], "argmin", [k for k, _ in argmins],
          "plan lens", [out["plan_%d" % i].shape[0] for i in range(3)], "final len", pred0.shape, "score", score0)


if __name__ == "__main__":
    main()
