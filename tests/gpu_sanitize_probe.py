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
batch = synthetic_train_batch(2, seed=5, end_ind=[61, 198])
ei = batch["end_ind"].numpy()
d = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
res = eng.forward_loss(d["traj_seq"], d["pad_mask"], d["end_ind"], d["states"], d["actions"], d["eps"], np.zeros(2, np.int64),
                       np.ones(2, np.int64), np.zeros(2, np.int64), ei, I_0=d["I_0"], I_g=d["I_g"],
                       want=("nll_per_frame", "kl_per_seq", "images_df"))
torch.cuda.synchronize()
print("ok", cost.tolist(), res["losses"].tolist())
eng.close()
