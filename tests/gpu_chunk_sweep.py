"""Decoder slot-chunk sweep (measurement aid): rollout time vs decoder_slot_chunk at B=1024."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_gcp_b200 import hparams
from video_gcp_b200.engine import Engine
from video_gcp_b200.synthetic import synthetic_rollout_inputs, synthetic_state_dict

B = 1024
dev = torch.device("cuda:0")
hp = hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1))
sd = synthetic_state_dict(hp, 1)
inp = synthetic_rollout_inputs(B, seed=5, shared_images=True)
I0, Ig, z, ei = inp["I_0"][:1].to(dev), inp["I_g"][:1].to(dev), inp["z"].to(dev), inp["end_ind"].to(dev)
for chunk in [int(a) for a in sys.argv[1:]] or [64, 32, 16, 8, 4]:
    eng = Engine(dev, max_candidates=B, decoder_slot_chunk=chunk)
    eng.load_weights(sd)
    for _ in range(2):
        eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
    e1.record(); torch.cuda.synchronize()
    eng.profile_enable(True)
    for _ in range(2):
        eng.rollout(I0, Ig, z, end_ind=ei, images_shared=True)
    torch.cuda.synchronize()
    pr = eng.profile_read()
    print("chunk %3d: rollout %.2f ms | %s" % (chunk, e0.elapsed_time(e1) / 5,
          {k: round(v / 2, 2) for k, v in pr.items() if isinstance(v, float)}), flush=True)
    eng.close(); del eng
    torch.cuda.empty_cache()
