"""Where does a whole ImageCEMPlanner.__call__ spend its host time (measurement aid, not a test)?  Six plans, each stage
synchronised and timed on the host, plus the caching allocators' counters (a cudaMalloc / cudaHostAlloc inside a plan shows
up as a jump)."""
import os, sys, time
from functools import partial
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.model import TreeModel
from video_gcp_b200.planning import GCPImageSimulator, ImageCEMPlanner, L2ImageCost, SimpleTreeCEMSampler
from video_gcp_b200.planning import cem_simulator
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
p = ImageCEMPlanner(dict(batch_size=N, n_iters=2, elite_frac=0.1, cost_fcn=L2ImageCost, dense_cost=True,
                         final_step_cost_weight=1.0, sampler=partial(SimpleTreeCEMSampler, n_level_hierarchy=8),
                         max_seq_len=200, action_dim=256, initial_std=0.3, max_rollout_bs=N, seed=11,
                         prune_before_decode=True), sim)
marks = []


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        marks.append((label, (time.perf_counter() - t0) * 1e3))
        return out
    setattr(obj, name, inner)


wrap(p, "cem_iteration", "iter")
wrap(p, "_elite_samples", "elite_z")
wrap(sim, "rollout_device", "  rollout_device")
wrap(cem_simulator.DeviceRollouts, "to_host", "  to_host")
wrap(p, "_rollout_host", "final(total)")
for i in range(6):
    marks.clear()
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    p(state, goal)
    dt = (time.perf_counter() - t0) * 1e3
    st1 = torch.cuda.memory_stats()
    hs = getattr(torch.cuda, "host_memory_stats", lambda: {})()
    print("plan %d: %.2f ms | %s | cudaMalloc +%d, host allocs %s" % (
        i, dt, ", ".join("%s %.2f" % m for m in marks), st1["num_device_alloc"] - st0["num_device_alloc"],
        hs.get("num_host_alloc", hs.get("host_alloc.all.allocated", "?"))))
