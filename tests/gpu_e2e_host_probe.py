"""Measurement aid: host-side time line of one bench-style e2e step (where the GPU idles between synchronised steps)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.model import TreeModel
from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost, SimpleTreeCEMSampler
from video_gcp_b200.synthetic import synthetic_state_dict

B = 1024
dev = torch.device("cuda:0")
model = TreeModel(hparams.gcp_tree_25room_config(batch_size=1), None, max_candidates=B)
model.load_state_dict(synthetic_state_dict(model._hp, 1), strict=True); model.device = dev; model.eval()
eng = model.engine
sim = GCPImageSimulator(model, append_latent=False)
cost_fcn = L2ImageCost(True, 1.0)
sampler = SimpleTreeCEMSampler(float("inf"), 200, 256, 0.3, n_level_hierarchy=8).attach(eng, seed=7)
r = np.random.default_rng(0)
state_t = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()
goal_t = torch.as_tensor(r.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)).pin_memory()
z_host = sampler.sample_device(B).cpu().pin_memory()
marks = {}
orig = eng.lib.gcpb200_rollout
def wrapped(*a):
    marks["c_in"] = time.perf_counter(); rc = orig(*a); marks["c_out"] = time.perf_counter(); return rc
eng.lib.gcpb200_rollout = wrapped

def step():
    t0 = time.perf_counter()
    ro = sim.rollout_device(state_t, goal_t, z_host, 200)
    t1 = time.perf_counter()
    cost = cost_fcn.device_cost(ro); idx, val = eng.topk(cost, 102); mean, std = eng.refit(ro.z, idx)
    t2 = time.perf_counter()
    c = cost.cpu(); i = idx.cpu()
    t3 = time.perf_counter()
    return t0, t1, t2, t3

for _ in range(3): step()
torch.cuda.synchronize()
acc = np.zeros(6)
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    t0, t1, t2, t3 = step()
    acc += np.array([marks["c_in"] - t0, marks["c_out"] - marks["c_in"], t1 - marks["c_out"], t2 - t1, t3 - t2, t3 - t0])
e1.record(); torch.cuda.synchronize()
names = ["python before the C call", "gcpb200_rollout (enqueue)", "python after the C call", "cost/topk/refit enqueue", "D2H reads (wait for the GPU)", "step wall"]
for k, v in zip(names, acc / n * 1e3):
    print("%-32s %.3f ms" % (k, v))
print("device time per step (events): %.3f ms" % (e0.elapsed_time(e1) / n))
