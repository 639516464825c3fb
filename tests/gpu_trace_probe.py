import os, sys
import torch
sys.path.insert(0, "/root/repo")
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict
B = 1024
dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
eng = Engine(dev, max_candidates=B); eng.load_weights(synthetic_state_dict(hp, 1))
inp = synthetic_rollout_inputs(B, seed=5, shared_images=True)
I0, Ig, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["end_ind"].to(dev)
zh = inp["z"].pin_memory()
for _ in range(5):
    eng.rollout(I0, Ig, zh, end_ind=ei, images_shared=True)
torch.cuda.synchronize()
