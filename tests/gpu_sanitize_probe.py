"""Small end-to-end pass for compute-sanitizer (measurement aid):
    compute-sanitizer --tool memcheck python tests/gpu_sanitize_probe.py
One 4-candidate rollout (device and pinned-host noise), cost / top-k / refit, and one B=2 training-phase forward + loss."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict, synthetic_train_batch

dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))
eng = Engine(dev, max_candidates=128, attach_cost_mdl=True)
eng.load_weights(synthetic_state_dict(hp, 1))
inp = synthetic_rollout_inputs(4, seed=11, shared_images=False)
for z in (inp["z"].to(dev), inp["z"].pin_memory()):
    out = eng.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), z, end_ind=inp["end_ind"].to(dev))
    cost = eng.cost_l2(out["images_df"], out["end_ind"], inp["I_g"][0].to(dev), True, 1.0)
    idx, _ = eng.topk(cost, 2)
    eng.refit(out["z"], idx)
torch.cuda.synchronize()
# round 2: planner mode (kept nodes only + fused L2 cost, with and without image writes), full decode with the fused cost,
# radix-select top-k with NaN / ties, refit split over elites, a 256-row rollout (CTA-pair GEMMs from level 1 on)
goal = inp["I_g"][0].to(dev)
for kw in (dict(decode_kept_only=True), dict(decode_kept_only=True, want_images=False), dict()):
    o2 = eng.rollout(inp["I_0"].to(dev), inp["I_g"].to(dev), inp["z"].to(dev), end_ind=inp["end_ind"].to(dev), l2_goal=goal,
                     l2_dense=True, l2_final_step_weight=1.0, **kw)
    assert torch.isfinite(o2["l2_cost"]).all()
c = torch.randn(5000, device=dev)
c[::7] = float("nan")
c[1::11] = c[3]
idx, val = eng.topk(c, 500)
big = synthetic_rollout_inputs(200, seed=12, shared_images=True)
eng2 = Engine(dev, max_candidates=256, attach_cost_mdl=True)
eng2.load_weights(synthetic_state_dict(hp, 1))
o3 = eng2.rollout(big["I_0"][:1].to(dev), big["I_g"][:1].to(dev), big["z"].to(dev), images_shared=True, decode_kept_only=True,
                  want_images=False, l2_goal=goal)
o4 = eng2.rollout(big["I_0"][:1].to(dev), big["I_g"][:1].to(dev), big["z"].to(dev), images_shared=True, decode_kept_only=True,
                  want_images=False, want_existence=False, want_aux=False, l2_goal=goal, tree_kept_only=True, sort_sampled_lengths=True)
assert torch.isfinite(o4["l2_cost"]).all()
i3, _ = eng2.topk(o3["l2_cost"], 20)
eng2.refit(o3["z"], i3)
torch.cuda.synchronize()
eng2.close()
batch = synthetic_train_batch(2, seed=5, end_ind=[61, 198])
ei = batch["end_ind"].numpy()
d = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
res = eng.forward_loss(d["traj_seq"], d["pad_mask"], d["end_ind"], d["states"], d["actions"], d["eps"], np.zeros(2, np.int64),
                       np.ones(2, np.int64), np.zeros(2, np.int64), ei, I_0=d["I_0"], I_g=d["I_g"],
                       want=("nll_per_frame", "kl_per_seq", "images_df"))
torch.cuda.synchronize()
print("ok", cost.tolist(), res["losses"].tolist())
eng.close()
