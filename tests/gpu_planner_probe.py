"""Host-side probe of the planner step (not a test): where does a synchronised CEM iteration spend its time?"""
import os, sys, time
from functools import partial
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.model import TreeModel
from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
from video_gcp_b200.synthetic import synthetic_state_dict

dev = torch.device("cuda:0")
N = 1024
model = TreeModel(hparams.gcp_tree_25room_config(batch_size=1), None, max_candidates=N)
model.load_state_dict(synthetic_state_dict(model._hp, 1), strict=True)
model.device = dev
model.eval()
sim = GCPImageSimulator(model, append_latent=False)
r = np.random.default_rng(0)
state = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()
goal = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()


def planner(pruned):
    p = ImageCEMPlanner(dict(batch_size=N, n_iters=1, elite_frac=0.1, cost_fcn=L2ImageCost, dense_cost=True,
                             final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=8),
                             max_seq_len=200, action_dim=256, initial_std=0.3, max_rollout_bs=N, seed=7,
                             prune_before_decode=pruned), sim)
    p._sampler.init()
    return p


def wall(fn, n=6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    return "host %.2f ms, host+gpu %.2f ms" % (np.median([a for a, _ in ts]), np.median([b for _, b in ts]))


if os.environ.get("PROBE_NCU") == "1":       # three planner-mode iterations and out: the launch list for ncu
    p = planner(True)
    for _ in range(3):
        p.cem_iteration(state, goal)
    torch.cuda.synchronize()
    sys.exit(0)

for pruned in (False, True):
    p = planner(pruned)
    print("pruned", pruned, "cem_iteration:", wall(lambda: p.cem_iteration(state, goal)))

    def e2e():
        c, i, v, _ = p.cem_iteration(state, goal)
        return c.cpu(), i.cpu()
    print("pruned", pruned, "cem_iteration + .cpu():", wall(e2e))
    st = torch.cuda.memory_stats()
    print("   cudaMalloc calls so far:", st.get("num_device_alloc"), "retries", st.get("num_alloc_retries"))

p = planner(False)
z = p._sampler.sample_device(N)
print("rollout_device only:", wall(lambda: sim.rollout_device(state, goal, z, 200)))
print("rollout_device planner_mode full+l2:", wall(lambda: sim.rollout_device(state, goal, z, 200, planner_mode=dict(kept_only=False, images=True, l2=(True, 1.0)))))
print("sample_device:", wall(lambda: p._sampler.sample_device(N)))
c = torch.randn(N, device=dev)
print("topk:", wall(lambda: model.engine.topk(c, 102)))
idx, _ = model.engine.topk(c, 102)
print("refit:", wall(lambda: model.engine.refit(z, idx)))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    cc, ii, vv, _ = p.cem_iteration(state, goal)
    cc.cpu()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
