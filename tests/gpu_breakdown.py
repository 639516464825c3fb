"""Where does a CEM step spend its time?  (measurement aid; run on the B200 box)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict

def ev_time(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, r

B = 1024
dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
eng = Engine(dev, max_candidates=B); eng.load_weights(synthetic_state_dict(hp, 1))
inp = synthetic_rollout_inputs(B, seed=5, shared_images=True)
I0, Ig, z, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
zh = inp["z"].pin_memory()
print("pinned:", zh.is_pinned())
ms, wall, _ = ev_time(lambda: zh.to(dev, non_blocking=True)); print("H2D z 267MB: %.2f ms (wall %.2f) -> %.1f GB/s" % (ms, wall, 0.267 / ms * 1e3))
ms, wall, _ = ev_time(lambda: torch.empty(B, 255, 3, 32, 32, device=dev)); print("torch.empty images: %.3f ms wall %.3f" % (ms, wall))
ms, wall, out = ev_time(lambda: eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)); print("rollout (injected len): %.2f ms (wall %.2f)" % (ms, wall))
ms, wall, out = ev_time(lambda: eng.rollout(I0, Ig, z, end_ind=None, seed=3, images_shared=True)); print("rollout (sampled len): %.2f ms (wall %.2f)" % (ms, wall))
ms, wall, out2 = ev_time(lambda: eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True, want_images=False)); print("rollout without decoder: %.2f ms" % ms)
ms, wall, cost = ev_time(lambda: eng.cost_l2(out["images_df"], out["end_ind"], Ig[0], True, 1.0)); print("cost_l2: %.3f ms  (%.1f GB/s algorithmic)" % (ms, float((out["end_ind"] + 1).sum()) * 12288 / ms / 1e6))
ms, wall, tk = ev_time(lambda: eng.topk(cost, 102)); print("topk: %.3f ms" % ms)
ms, wall, _ = ev_time(lambda: eng.refit(z, tk[0])); print("refit: %.3f ms" % ms)
ms, wall, _ = ev_time(lambda: eng.prune_gather(out["images_df"], out["end_ind"])); print("prune_gather images: %.3f ms" % ms)
ms, wall, _ = ev_time(lambda: eng.sample_noise(B, std_scalar=0.3, seed=1)); print("sample_noise: %.3f ms (%.1f GB/s)" % (ms, 0.267 / ms * 1e3))
ms, wall, _ = ev_time(lambda: int(out["end_ind"].max())); print("end_ind.max sync: %.3f ms wall %.3f" % (ms, wall))
eng.profile_enable(True)
for _ in range(2): eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
torch.cuda.synchronize(); print({k: (round(v / 2, 3) if isinstance(v, float) else v) for k, v in eng.profile_read().items()})

# ---- bench-like step through the simulator API
from video_gcp_b200.model import TreeModel
from video_gcp_b200.planning import GCPImageSimulator, L2ImageCost
eng.close(); del out, out2; torch.cuda.empty_cache()
model = TreeModel(hparams.gcp_tree_25room_config(batch_size=1), None, max_candidates=B)
model.load_state_dict(synthetic_state_dict(model._hp, 1)); model.device = dev; model.eval()
sim = GCPImageSimulator(model, append_latent=False); cf = L2ImageCost(True, 1.0)
state = np.random.rand(1, 32, 32, 3).astype(np.float32); goal = np.random.rand(1, 32, 32, 3).astype(np.float32)
def step():
    ro = sim.rollout_device(state, goal, z, 200)
    c = cf.device_cost(ro); i, v = model.engine.topk(c, 102); return model.engine.refit(z, i)
ms, wall, _ = ev_time(step, 5); print("simulator step: %.2f ms (wall %.2f)" % (ms, wall))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
